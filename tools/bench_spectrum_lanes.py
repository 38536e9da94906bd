#!/usr/bin/env python
"""cfg4's spectrum analyzer (16384 / hop 1024 / peak hold) at a few lane counts, device-resident: which path serves them (fused whole-lane
kernel from SMs / 2 lanes up, the two-kernel path below) and how fast.   python tools/bench_spectrum_lanes.py [lanes ...]"""
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
from openmeters_b200.processors import SpectrumConfig  # noqa: E402


def main():
    lanes_list = [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64, 128]
    api = lib_api()
    api.set_device(0)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream(dev).cuda_stream
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, floor_db=-100.0)
    S = 20 * 48000
    base = synth.cfg4_streams(4, S / 48000.0).reshape(8, -1)[:, :S]
    out = {}
    for L in lanes_list:
        x = torch.from_numpy(np.ascontiguousarray(np.tile(base, ((L + 7) // 8, 1))[:L])).to(dev)
        plan = batch.SpectrumPlan(cfg, api=api)
        hops = (S - 16384) // 1024 + 1
        w = torch.empty((L, hops, 8193), dtype=torch.float32, device=dev)
        r = torch.empty_like(w)
        pk = torch.empty((L, hops), dtype=torch.int32, device=dev)
        fn = lambda: plan.execute_device(x.data_ptr(), L, S, S, w.data_ptr(), r.data_ptr(), pk.data_ptr(), stream=st)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[str(L)] = {"lane_hops_per_s": L * hops / (ms * 1e-3), "ms": ms}
        del w, r, pk, x, plan
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
