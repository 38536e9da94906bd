#!/usr/bin/env python
"""Throughput over the product's settings grid: the 18 (size, hop, zero padding, window, mode) combinations
of tests/cases.py::settings_grid plus the product's default spectrogram configuration, device-resident inputs, CUDA events.
Shows which kernel tier serves each point of the config space and how fast.  Prints one JSON object."""
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
from openmeters_b200.processors import SpectrogramConfig  # noqa: E402
from tests.cases import settings_grid  # noqa: E402

TIER = {1: "stft_fast.cu", 2: "stft_fast2.cu", 3: "stft_classic_fast.cu", 4: "stft_fast8k.cu", 5: "stft_fast2k.cu", 6: "stft_fast1k.cu", 7: "stft_r64.cu", 8: "stft_r64x.cu"}


def timed(fn, iters=5, warm=2):
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


def main():
    api = lib_api()
    api.set_device(0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    # (spectrogram/processor.rs:47-59) product defaults first, then the grid
    grid = [(2048, 64, 1, capi.WINDOW_HANN, True), (2048, 64, 1, capi.WINDOW_HANN, False),
            # N = 4096 reassigned at every power-of-two hop of the UI (N/4 ... N/128)
            (4096, 1024, 1, capi.WINDOW_BLACKMAN_HARRIS, True), (4096, 512, 1, capi.WINDOW_BLACKMAN_HARRIS, True),
            (4096, 256, 1, capi.WINDOW_BLACKMAN_HARRIS, True), (4096, 128, 1, capi.WINDOW_BLACKMAN_HARRIS, True),
            (4096, 64, 1, capi.WINDOW_BLACKMAN_HARRIS, True), (4096, 32, 1, capi.WINDOW_BLACKMAN_HARRIS, True),
            (1024, 256, 1, capi.WINDOW_HANN, True), (1024, 32, 1, capi.WINDOW_HANN, True),
            (8192, 2048, 1, capi.WINDOW_BLACKMAN_HARRIS, True), (8192, 256, 1, capi.WINDOW_BLACKMAN_HARRIS, True),
            (16384, 4096, 1, capi.WINDOW_BLACKMAN_HARRIS, True), (16384, 256, 1, capi.WINDOW_HANN, True)] + settings_grid()
    if "--first" in sys.argv:  # e.g. --first 1: only the product default (short enough to run under ncu)
        grid = grid[:int(sys.argv[sys.argv.index("--first") + 1])]
    base = synth.cfg2_lanes(8, 4.0)
    res = []
    for n, hop, zp, window, reassign in grid:
        cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=window, use_reassignment=reassign, zero_padding_factor=zp)
        F = n * zp
        bins = F // 2 + 1
        out_bytes = bins * (12 if reassign else 2)
        L = 64
        frames = int(max(8, min(1024, (1 << 30) // (out_bytes * L))))  # <= 1 GiB of output per pass
        S = (2 * n if reassign else n) + (frames - 1) * hop
        S = (S + 3) // 4 * 4
        reps = (S + base.shape[1] - 1) // base.shape[1]
        lanes_np = np.tile(base, (L // 8, reps))[:, :S]
        lanes = torch.from_numpy(np.ascontiguousarray(lanes_np, np.float32)).to(dev)
        plan = batch.StftPlan(cfg, api=api)
        Fr = plan.frames_per_lane(S)
        if reassign:
            pts = torch.empty((L * Fr, bins, 3), dtype=torch.float32, device=dev)
            cnt = torch.empty((L * Fr,), dtype=torch.int32, device=dev)
            t = timed(lambda: plan.execute_device(lanes.data_ptr(), L, S, S, pts.data_ptr(), bins, cnt.data_ptr(), stream=st))
            del pts, cnt
        else:
            out = torch.empty((L * Fr, bins), dtype=torch.int16, device=dev)
            t = timed(lambda: plan.execute_device(lanes.data_ptr(), L, S, S, classic_ptr=out.data_ptr(), stream=st))
            del out
        if plan.kernel_generation:
            tier = TIER[plan.kernel_generation]
        elif (not reassign and F <= 16384) or (reassign and F <= 8192):
            tier = "stft_smem.cu"
        elif n <= 8192:
            tier = "stft_smem.cu (%d residue transforms)" % (F // 8192)
        else:
            tier = "stft_generic.cu"
        alg = hop * 4 + out_bytes + (4 if reassign else 0)
        res.append(dict(fft_size=n, hop=hop, zero_pad=zp, window=window, reassigned=reassign, tier=tier, frames=L * Fr,
                        frames_per_s=L * Fr / t, ms=t * 1e3, algorithmic_gbs=L * Fr * alg / t / 1e9,
                        realtime_48k_streams=(L * Fr / t) / (48000.0 / hop)))
        del plan, lanes
        torch.cuda.empty_cache()
    print(json.dumps({"settings_grid": res}))


if __name__ == "__main__":
    main()
