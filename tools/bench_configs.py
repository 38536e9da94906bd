#!/usr/bin/env python
"""Secondary throughput measurements for the other BASELINE configs (cfg1, cfg3, cfg4, cfg5), device-resident inputs,
CUDA events on the launching stream.  Prints one JSON object; `bench.py` (cfg2) stays the headline."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmeters_b200 import _capi as capi  # noqa: E402
from openmeters_b200 import batch, synth  # noqa: E402
from openmeters_b200._lib import api as lib_api  # noqa: E402
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / 1000.0


def tile_lanes(base: np.ndarray, n: int) -> np.ndarray:
    reps = (n + base.shape[0] - 1) // base.shape[0]
    out = np.concatenate([np.roll(base, 977 * r, axis=1) * np.float32(1.0 - 0.01 * r) for r in range(reps)], 0)[:n]
    return np.ascontiguousarray(out, np.float32)


def main():
    only = None
    if "--only" in sys.argv:  # e.g. --only cfg5 (one config: short enough to run under ncu)
        only = sys.argv[sys.argv.index("--only") + 1].split(",")
    want = lambda name: only is None or name in only
    api = lib_api()
    api.set_device(0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    res = {}

    if want("cfg1"):
        # cfg1: classic 1024/512 Hann
        cfg = SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False)
        L, S = 256, 1 << 18
        lanes = torch.from_numpy(tile_lanes(synth.cfg2_lanes(8, S / 48000.0)[:, :S], L)).to(dev)
        plan = batch.StftPlan(cfg, api=api)
        F = plan.frames_per_lane(S)
        out = torch.empty((L * F, plan.bins), dtype=torch.int16, device=dev)
        t = timed(lambda: plan.execute_device(lanes.data_ptr(), L, S, S, classic_ptr=out.data_ptr(), stream=st))
        b = 512 * 4 + 513 * 2
        res["cfg1_classic_1024_512"] = dict(frames_per_s=L * F / t, ms=t * 1e3, algorithmic_bytes_per_frame=b, achieved_gbs=L * F * b / t / 1e9,
                                            hbm_frac=L * F * b / t / 1e9 / PEAK)
        del plan, out, lanes

    if want("cfg5"):
        # cfg5: reassigned 8192/2048 BH @96k
        cfg = SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=2048, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
        L, S = 32, 1 << 20
        lanes = torch.from_numpy(tile_lanes(synth.cfg5_lanes(8, S), L)).to(dev)
        plan = batch.StftPlan(cfg, api=api)
        F = plan.frames_per_lane(S)
        pts = torch.empty((L * F, plan.bins, 3), dtype=torch.float32, device=dev)
        cnt = torch.empty((L * F,), dtype=torch.int32, device=dev)
        t = timed(lambda: plan.execute_device(lanes.data_ptr(), L, S, S, pts.data_ptr(), plan.bins, cnt.data_ptr(), stream=st), iters=5)
        b = 2048 * 4 + 4097 * 12 + 4
        res["cfg5_reassigned_8192_2048"] = dict(frames_per_s=L * F / t, ms=t * 1e3, algorithmic_bytes_per_frame=b, achieved_gbs=L * F * b / t / 1e9,
                                                hbm_frac=L * F * b / t / 1e9 / PEAK, kernel_generation=plan.kernel_generation)
        del plan, pts, cnt, lanes

    if want("cfg4"):
        # cfg4: spectrum 16384/1024 Hann, PeakHold 12 dB/s, 128 lanes
        cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, floor_db=-100.0)
        L, S = 128, 480000
        lanes = torch.from_numpy(tile_lanes(synth.cfg4_streams(4, S / 48000.0).reshape(8, -1)[:, :S], L)).to(dev)
        plan = batch.SpectrumPlan(cfg, api=api)
        Hh = plan.hops_per_lane(S)
        w = torch.empty((L * Hh, plan.bins), dtype=torch.float32, device=dev)
        r = torch.empty_like(w)
        pk = torch.empty((L * Hh,), dtype=torch.int32, device=dev)
        t = timed(lambda: plan.execute_device(lanes.data_ptr(), L, S, S, w.data_ptr(), r.data_ptr(), pk.data_ptr(), stream=st), iters=5)
        b = 1024 * 4 + 2 * 8193 * 4
        res["cfg4_spectrum_16384_1024"] = dict(lane_hops_per_s=L * Hh / t, ms=t * 1e3, algorithmic_bytes_per_lane_hop=b, achieved_gbs=L * Hh * b / t / 1e9,
                                               hbm_frac=L * Hh * b / t / 1e9 / PEAK)
        del plan, w, r, pk, lanes

    if want("cfg3"):
        # cfg3: loudness 8 ch 48 kHz, 16 streams x 30 s, snapshot every 1024 frames
        x = synth.cfg3_surround(30.0)
        nS = 16
        streams = torch.from_numpy(np.stack([x * np.float32(1.0 - 0.02 * i) for i in range(nS)])).to(dev)
        plan = batch.LoudnessPlan(LoudnessConfig(), 8, capi.SURROUND, api=api)
        frames = x.size // 8
        nb = (frames + 1023) // 1024
        snaps = torch.empty((nS * nb, 116 // 4), dtype=torch.float32, device=dev)
        assert snaps.element_size() * snaps.shape[1] == 116
        t = timed(lambda: plan.execute_device(streams.data_ptr(), nS, frames, frames * 8, 1024, snaps.data_ptr(), stream=st), iters=5)
        res["cfg3_loudness_8ch"] = dict(sample_channels_per_s=nS * frames * 8 / t, ms=t * 1e3, algorithmic_bytes_per_sample_channel=4,
                                        achieved_gbs=nS * frames * 8 * 4 / t / 1e9, hbm_frac=nS * frames * 8 * 4 / t / 1e9 / PEAK)
    if want("splat"):
        # row f2: splat accumulation of cfg2 columns (64 rings x 1024 columns x 2049 points) into 64 images of 1024 x 512
        import ctypes as C
        from openmeters_b200 import splat
        R, hl, stride = 64, 1000, 2049
        rng = np.random.default_rng(1)
        pts = torch.empty((R, hl, stride, 3), dtype=torch.float32, device=dev)
        pts[..., 0].uniform_(-2.5, 0.5)
        pts[..., 1] = torch.exp(torch.empty((R, hl, stride), device=dev).uniform_(float(np.log(1.0)), float(np.log(24000.0))))
        pts[..., 2] = 10.0 ** torch.empty((R, hl, stride), device=dev).uniform_(-14.0, 0.0)
        cnt = torch.full((R, hl), stride, dtype=torch.int32, device=dev)
        fmin, fmax = splat.display_axis(48000.0)
        p = splat.SplatParams(freq_min=fmin, freq_max=fmax, ring_capacity=hl, newest_col=hl - 1, col_count=hl, ext_w=1024.0, ext_h=512.0)
        c = p.to_c()
        acc = torch.empty((R, 512, 1024), dtype=torch.float32, device=dev)
        db = torch.empty_like(acc)

        def run():
            assert api.splat_accumulate_device(pts.data_ptr(), stride, cnt.data_ptr(), R, C.byref(c), acc.data_ptr(), st) == 0
            assert api.splat_resolve_device(acc.data_ptr(), R, C.byref(c), db.data_ptr(), st) == 0
        t = timed(run, iters=5)
        npts = R * hl * stride
        b = 12 * npts + 2 * 4 * acc.numel() + 4 * acc.numel()   # points in; image cleared + accumulated + resolved
        res["splat_accumulate_resolve"] = dict(points_per_s=npts / t, ms=t * 1e3, algorithmic_bytes=b, achieved_gbs=b / t / 1e9, hbm_frac=b / t / 1e9 / PEAK,
                                               columns_per_s=R * hl / t)
        del pts, cnt, acc, db
    if want("bank"):
        # row f1: live ring-buffer input at many streams.  512 lock-step streams, DspBatcher-sized blocks (1024 stereo frames at
        # 48 kHz), cfg2 analysis: one omb_spectrogram_bank_push per tick vs one omb_spectrogram_process_block per stream per tick.
        import time
        from openmeters_b200.meter import SpectrogramBank
        from openmeters_b200.processors import AudioBlock, SpectrogramProcessor
        cfgb = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True, history_length=64)
        S, nf, ticks = 512, 1024, 24
        rng = np.random.default_rng(0)
        blocks = rng.uniform(-0.5, 0.5, (S, nf * 2)).astype(np.float32)
        bank = SpectrogramBank(cfgb, S, api=api)
        for _ in range(10):
            bank.push(blocks, 2, 48000.0, copy=False)       # fills the first window (8 ticks) and warms up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cols = 0
        for _ in range(ticks):
            up = bank.push(blocks, 2, 48000.0, copy=False)
            cols += S * up.columns[0].shape[1]
        tb = time.perf_counter() - t0
        S1 = 32
        procs = [SpectrogramProcessor(cfgb, api=api) for _ in range(S1)]
        blks = [AudioBlock(blocks[s], 2, 48000.0) for s in range(S1)]
        for _ in range(10):
            for p_, b_ in zip(procs, blks):
                p_.process_block(b_)
        t0 = time.perf_counter()
        cols1 = 0
        for _ in range(ticks):
            for p_, b_ in zip(procs, blks):
                cols1 += len(p_.process_block(b_).new_columns)
        t1 = time.perf_counter() - t0
        res["bank_cfg2_live_streams"] = dict(streams=S, block_frames=nf, columns_per_s=cols / tb, ms_per_tick=tb / ticks * 1e3,
                                             realtime_streams=(cols / tb) / (48000.0 / 1024.0),
                                             per_stream_handles=dict(streams=S1, columns_per_s=cols1 / t1, ms_per_tick=t1 / ticks * 1e3,
                                                                     note="includes the Python mirror's per-column copies"))
    if want("loudbank"):
        # row f1, loudness: 512 lock-step 8-channel streams, 1024-frame blocks: one omb_loudness_bank_push per tick vs one
        # omb_loudness_process_block per stream per tick
        import time
        from openmeters_b200.meter import LoudnessBank
        from openmeters_b200.processors import AudioBlock, LoudnessProcessor
        S, nf, ticks = 512, 1024, 24
        x = synth.cfg3_surround(1.0)[: nf * 8]
        blocks = np.stack([x * np.float32(1.0 - 0.001 * i) for i in range(S)]).astype(np.float32)
        bank = LoudnessBank(LoudnessConfig(), S, api=api)
        for _ in range(4):
            bank.push(blocks, 8, 48000.0, capi.SURROUND)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ticks):
            bank.push(blocks, 8, 48000.0, capi.SURROUND)
        tb = (time.perf_counter() - t0) / ticks
        S1 = 32
        procs = [LoudnessProcessor(LoudnessConfig(), api=api) for _ in range(S1)]
        blks = [AudioBlock(blocks[i], 8, 48000.0, capi.SURROUND) for i in range(S1)]
        for p_, b_ in zip(procs, blks):
            p_.process_block(b_)
        t0 = time.perf_counter()
        for _ in range(ticks):
            for p_, b_ in zip(procs, blks):
                p_.process_block(b_)
        t1 = (time.perf_counter() - t0) / ticks
        res["bank_loudness_live_streams"] = dict(streams=S, block_frames=nf, channels=8, ms_per_tick=tb * 1e3,
                                                 sample_channels_per_s=S * nf * 8 / tb, realtime_streams=S * (nf / 48000.0) / tb,
                                                 per_stream_handles=dict(streams=S1, ms_per_tick=t1 * 1e3, realtime_streams=S1 * (nf / 48000.0) / t1))
    if want("specbank"):
        # row f1, spectrum analyzer: 256 lock-step stereo streams (Left + Right traces = 512 lanes), product defaults (16384 / 1024),
        # 1024-frame blocks: one omb_spectrum_bank_push per tick vs one omb_spectrum_process_block per stream per tick
        import time
        from openmeters_b200.meter import SpectrumBank
        from openmeters_b200.processors import AudioBlock, SpectrumProcessor
        scfg = SpectrumConfig(fft_size=16384, hop_size=1024, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, source=capi.CHANNEL_LEFT,
                              secondary_source=capi.CHANNEL_RIGHT, floor_db=-100.0)
        S, nf, ticks = 256, 1024, 24
        rng = np.random.default_rng(0)
        blocks = rng.uniform(-0.5, 0.5, (S, nf * 2)).astype(np.float32)
        bank = SpectrumBank(scfg, S, api=api)
        for _ in range(20):
            bank.push(blocks, 2, 48000.0, copy=False)      # fills the first window (16 ticks) and warms up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(ticks):
            bank.push(blocks, 2, 48000.0, copy=False)
        tb = (time.perf_counter() - t0) / ticks
        S1 = 16
        procs = [SpectrumProcessor(scfg, api=api) for _ in range(S1)]
        blks = [AudioBlock(blocks[i], 2, 48000.0) for i in range(S1)]
        for _ in range(20):
            for p_, b_ in zip(procs, blks):
                p_.process_block(b_)
        t0 = time.perf_counter()
        for _ in range(ticks):
            for p_, b_ in zip(procs, blks):
                p_.process_block(b_)
        t1 = (time.perf_counter() - t0) / ticks
        res["bank_spectrum_live_streams"] = dict(streams=S, traces=2, block_frames=nf, ms_per_tick=tb * 1e3, lane_hops_per_s=S * 2 / tb,
                                                 realtime_streams=S * (nf / 48000.0) / tb,
                                                 per_stream_handles=dict(streams=S1, ms_per_tick=t1 * 1e3, realtime_streams=S1 * (nf / 48000.0) / t1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
