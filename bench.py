#!/usr/bin/env python
"""bench.py — headline benchmark: STFT frames/s on BASELINE.json configs[1]
(4096-pt Blackman-Harris, hop 1024, time-frequency reassignment, 48 kHz mono lanes).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one JSON line on stdout)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference, host threads
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU, NCCL only for the barrier/max

A "step" is one pass of the hot path over one batch: `--lanes` mono lanes x `--samples` samples per GPU
(default 64 x 2^20 = 256 MiB of f32 PCM in, 1.6 GB of points out per step: both far larger than the 126 MB L2,
so nothing is L2-warm between steps).  Unit of work = one analysis frame of one lane.

  value      frames/s, inputs resident in HBM, outputs left in HBM; CUDA events on the launching stream,
             barrier + synchronize on both sides, max over ranks.
  e2e        same metric through the C-ABI host entry point (omb_stft_execute_host) with pinned HOST buffers:
             H2D of the PCM and D2H of points+counts inside the timed region.
  roofline   algorithmic bytes per frame (hop*4 + bins*12 + 4 = 28 688 B) x frames / kernel time,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json; plus the FP32 view (the path is
             FP32-issue bound).
  cpu_baseline  the oracle (CPU restatement of the reference algorithm, own FFT) on a bounded sample, all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from openmeters_b200 import _capi as capi  # noqa: E402
from openmeters_b200 import synth  # noqa: E402
from openmeters_b200.processors import SpectrogramConfig  # noqa: E402

WINDOW, HOP, SR = 4096, 1024, 48000.0
HILBERT = 2 * WINDOW
BINS = WINDOW // 2 + 1
POINT_STRIDE = BINS
ALGO_BYTES_PER_FRAME = HOP * 4 + BINS * 12 + 4  # 28 688 B
# FP32 instructions-independent flop model of the implemented algorithm: 5 complex 4096-pt FFTs
# (5 n log2 n each) + pair step + windows + per-bin reassignment.
FLOPS_PER_FRAME = 5 * 5 * 4096 * 12 + 4096 * 30 + 2049 * 40
METRIC = "STFT frames/s (4096-pt, hop 1024, 48 kHz, Blackman-Harris, time-frequency reassigned)"


def cfg2() -> SpectrogramConfig:
    return SpectrogramConfig(sample_rate=SR, fft_size=WINDOW, hop_size=HOP, window=capi.WINDOW_BLACKMAN_HARRIS,
                             use_reassignment=True, zero_padding_factor=1)


def frames_per_lane(samples: int) -> int:
    return (samples - HILBERT) // HOP + 1 if samples >= HILBERT else 0


def make_lanes(n_lanes: int, samples: int, first_lane: int) -> np.ndarray:
    """SURVEY §8(d) cfg2 signal per lane (chirp + uniform noise, seeded by global lane index)."""
    out = np.empty((n_lanes, samples), np.float32)
    for i in range(min(n_lanes, 8)):
        out[i] = synth.lane_signal(samples, SR, (i % 8 + 1) * 2500.0, 1000 + first_lane + i)
    # lanes beyond the first 8 reuse those signals with a lane-dependent circular shift and gain (cheap to build,
    # still distinct data; values stay in [-0.5, 0.5])
    for i in range(8, n_lanes):
        out[i] = np.roll(out[i % 8], 4099 * (i // 8)) * np.float32(1.0 - 0.01 * (i // 8))
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def recorded_traffic(kernel: str):
    """DRAM bytes per frame from the committed ncu --set full capture (profiles/traffic.json), or None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get(kernel)
        except Exception:
            return None
    return None


class ClockSampler:
    """nvidia-smi clock/throttle sampling during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- CPU arms
def oracle_rate(lanes: np.ndarray, target_seconds: float, threads: int = 0):
    """Times the oracle on a bounded sample of `lanes` sized for ~target_seconds. Returns (frames/s, frames, seconds, threads)."""
    from oracle import oracle_py

    api = oracle_py.api()
    threads = threads or int(api.hw_threads())
    cfg = cfg2()
    L = min(lanes.shape[0], max(threads, 1) * 2)
    probe_frames = 32
    probe = np.ascontiguousarray(lanes[:L, : HILBERT + (probe_frames - 1) * HOP])
    oracle_py.stft_batch(cfg, probe, threads=threads)  # warm caches / page in
    t0 = time.perf_counter()
    oracle_py.stft_batch(cfg, probe, threads=threads)
    rate = L * probe_frames / max(time.perf_counter() - t0, 1e-6)
    want = int(min(max(rate * 2.0 / L, probe_frames), frames_per_lane(lanes.shape[1])))  # ~2 s per pass
    sample = np.ascontiguousarray(lanes[:L, : HILBERT + (want - 1) * HOP])
    frames = 0
    out = oracle_py.stft_batch(cfg, sample, threads=threads)  # allocates + touches the output once (untimed)
    t0 = time.perf_counter()
    while True:  # repeat the bounded sample until ~target_seconds of CPU work have been timed
        _, cnt = oracle_py.stft_batch(cfg, sample, threads=threads, out=out)
        frames += int(cnt.size)
        dt = time.perf_counter() - t0
        if dt >= target_seconds:
            break
    return frames / dt, frames, dt, threads


def run_reference(args):
    """`--impl reference`: the reference's CPU algorithm (oracle restatement; the Rust crate cannot be built here:
    no cargo/rustc, FFT arithmetic in un-vendored rustfft) on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    lanes = make_lanes(min(args.lanes, 16), min(args.samples, HILBERT + 2047 * HOP), 0)
    budget = 150.0 / max(args.steps + args.warmup, 1)  # whole run within a few minutes
    per_step = min(max(budget, 0.5), 20.0)
    rates, frames, secs, threads = [], 0, 0.0, 0
    for i in range(args.warmup + args.steps):
        r, f, s, threads = oracle_rate(lanes, per_step)
        if i >= args.warmup:
            rates.append(r)
            frames, secs = f, s
    value = float(np.mean(rates))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 4096-pt BH reassigned STFT hop 1024, 48 kHz mono lanes (BASELINE configs[1])",
                   "sample_frames_per_step": frames},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{frames} frames/step of the cfg2 workload; CPU restatement of processor.rs with its own radix-2 FFT, not rustfft/AVX"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch

    from openmeters_b200 import batch
    from openmeters_b200._lib import api as lib_api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    api = lib_api()
    assert api.set_device(local) == 0, api.last_error()
    dev = torch.device("cuda", local)

    L, S = args.lanes, args.samples
    F = frames_per_lane(S)
    frames_per_step = L * F
    host_lanes = make_lanes(L, S, first_lane=rank * L)
    pin_in = torch.from_numpy(host_lanes).pin_memory()
    d_lanes = pin_in.to(dev, non_blocking=True)
    d_points = torch.empty((frames_per_step, POINT_STRIDE, 3), dtype=torch.float32, device=dev)
    d_counts = torch.empty((frames_per_step,), dtype=torch.int32, device=dev)
    plan = batch.StftPlan(cfg2(), kernel={"auto": capi.KERNEL_AUTO, "generic": capi.KERNEL_GENERIC, "fast": capi.KERNEL_FAST}[args.kernel],
                          api=api)
    stream = torch.cuda.current_stream(dev)

    def step():
        plan.execute_device(d_lanes.data_ptr(), L, S, S, d_points.data_ptr(), POINT_STRIDE, d_counts.data_ptr(), 0, stream.cuda_stream)

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = api.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    launches = int(api.kernel_launch_count() - launches0)
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    checksum = d_counts.to(torch.int64).sum().reshape(1)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(checksum, op=dist.ReduceOp.SUM)  # proof that every rank produced its columns (outside the timed region)
    ms_step = float(t.item()) / args.steps
    value = world * frames_per_step / (ms_step / 1000.0)

    # ---- e2e through the C-ABI host entry point: pinned host PCM in, points + counts back on the host
    h_points = torch.empty((frames_per_step, POINT_STRIDE, 3), dtype=torch.float32).pin_memory()
    h_counts = torch.empty((frames_per_step,), dtype=torch.int32).pin_memory()

    def step_e2e():
        rc = api.stft_execute_host(plan._h, pin_in.data_ptr(), L, S, S, h_points.data_ptr(), POINT_STRIDE, h_counts.data_ptr(), None)
        assert rc == 0, api.last_error()

    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * frames_per_step / float(te.item())
    assert int(h_counts.to(torch.int64).sum().item()) == int(d_counts.to(torch.int64).sum().item())

    if rank != 0:
        return 0
    peak_gbs, peak_src = measured_peaks()
    # second roofline: the FP32 FMA peak of this device, measured by the library's probe kernel after the timed region
    import ctypes as _C
    fp32_peak = _C.c_double(0.0)
    if api.probe_fp32_tflops(_C.byref(fp32_peak)) != 0:
        fp32_peak.value = 0.0
    kernel_name = {0: "k_reassigned_generic", 1: "k_reassigned_fast", 2: "k_reassigned_fast2"}[plan.kernel_generation]
    achieved_gbs = frames_per_step * ALGO_BYTES_PER_FRAME / (ms_step / 1000.0) / 1e9
    traffic = recorded_traffic(kernel_name)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg2: 4096-pt BH reassigned STFT hop 1024, 48 kHz mono lanes (BASELINE configs[1])",
                   "lanes_per_gpu": L, "samples_per_lane": S, "frames_per_step_per_gpu": frames_per_step,
                   "sharding": f"lanes x{world} (no data-path collective)", "kernel": kernel_name,
                   "l2": "inputs (%d MiB) and outputs (%d MiB) per step exceed the 126 MB L2" % (L * S * 4 >> 20, frames_per_step * POINT_STRIDE * 12 >> 20),
                   "points_checksum": int(checksum.item())},
        "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
                     "traffic": (traffic * frames_per_step if traffic else None), "peak_source": peak_src, "kernel": kernel_name,
                     "algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME,
                     "fp32": {"flops_per_frame": FLOPS_PER_FRAME, "achieved_tflops": frames_per_step * FLOPS_PER_FRAME / (ms_step / 1000.0) / 1e12,
                              "peak_tflops": fp32_peak.value or None, "peak_source": "measured (omb_probe_fp32_tflops: independent FFMA chains)",
                              "frac": (frames_per_step * FLOPS_PER_FRAME / (ms_step / 1000.0) / 1e12 / fp32_peak.value) if fp32_peak.value else None,
                              "note": "path is FP32-issue / shared-memory bound, not HBM bound (arithmetic intensity ~45 flop/B); see DESIGN.md"}},
        "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": int(L * S * 4),
                "d2h_bytes_per_step": int(frames_per_step * (POINT_STRIDE * 12 + 4)), "steps": e2e_steps, "api": "omb_stft_execute_host"},
        "gpu_launches": launches, "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        r, f, s, th = oracle_rate(host_lanes, args.cpu_seconds)
        line["cpu_baseline"] = {"value": r, "unit": "frames/s", "cores": th, "kind": "port",
                                "sample": f"{f} frames of the same cfg2 lanes in {s:.1f} s; CPU restatement of processor.rs (own radix-2 FFT, not rustfft/AVX)"}
    print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lanes", type=int, default=64, help="mono lanes per GPU")
    ap.add_argument("--samples", type=int, default=1 << 20, help="samples per lane")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fast"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
