#!/usr/bin/env python
"""Sharded runs of BASELINE configs[3] (cfg4: 64 streams x 2 ch spectrum analyzer, 4 GPUs) and configs[4] (cfg5: 8192-pt
reassigned spectrogram at 96 kHz, 8 GPUs) under torchrun, one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_sharded.py --config cfg4|cfg5

Rank 0 holds all the PCM and scatters every rank its lanes with NCCL point-to-point sends over NVLink
(openmeters_b200.sharding.scatter_lanes: rank r owns lanes l % R == r); every rank then runs the single-GPU kernels on its
own lanes, outputs stay sharded and GPU-resident, and the ranks all-gather [lane-hops or frames, checksum] so that any rank
can prove the whole job ran.  Two rates are reported: "resident" (kernels only, max over ranks) and
"scatter_inclusive" (scatter + kernels).  Rank 0 prints one JSON object."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmeters_b200 import _capi as capi  # noqa: E402
from openmeters_b200 import batch, sharding, synth  # noqa: E402
from openmeters_b200._lib import api as lib_api  # noqa: E402
from openmeters_b200.processors import SpectrogramConfig, SpectrumConfig  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"])
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    api = lib_api()
    assert api.set_device(local) == 0
    st = torch.cuda.current_stream(dev)

    if args.config == "cfg4":
        n_lanes, S = 128, 480000  # 64 streams x (Left, Right), 10 s at 48 kHz
        cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, floor_db=-100.0)
        plan = batch.SpectrumPlan(cfg, api=api)
        units_per_lane = plan.hops_per_lane(S)
        unit = "lane-hops/s"
        make = lambda: np.concatenate([np.roll(synth.cfg4_streams(4, S / 48000.0).reshape(8, -1)[:, :S], 977 * r, axis=1) for r in range(16)], 0)
    else:
        n_lanes, S = 256, 1 << 19    # 8 lanes per GPU-share x 32 streams' worth, 5.5 s at 96 kHz
        cfg = SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=2048, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
        plan = batch.StftPlan(cfg, api=api)
        units_per_lane = plan.frames_per_lane(S)
        unit = "frames/s"
        make = lambda: np.concatenate([np.roll(synth.cfg5_lanes(8, S), 4099 * r, axis=1) for r in range(32)], 0)
    all_lanes = torch.from_numpy(np.ascontiguousarray(make(), np.float32)).to(dev) if rank == 0 else None

    def scatter():
        return sharding.scatter_lanes(all_lanes, n_lanes, S, src=0, device=dev)

    mine_lanes, mine = scatter()
    k = len(mine)
    if args.config == "cfg4":
        w = torch.empty((k * units_per_lane, plan.bins), dtype=torch.float32, device=dev)
        r = torch.empty_like(w)
        pk = torch.empty((k * units_per_lane,), dtype=torch.int32, device=dev)

        def compute(x):
            plan.execute_device(x.data_ptr(), k, S, S, w.data_ptr(), r.data_ptr(), pk.data_ptr(), stream=st.cuda_stream)

        checksum = lambda: int(pk.to(torch.int64).sum().item())
    else:
        pts = torch.empty((k * units_per_lane, plan.bins, 3), dtype=torch.float32, device=dev)
        cnt = torch.empty((k * units_per_lane,), dtype=torch.int32, device=dev)

        def compute(x):
            plan.execute_device(x.data_ptr(), k, S, S, pts.data_ptr(), plan.bins, cnt.data_ptr(), stream=st.cuda_stream)

        checksum = lambda: int(cnt.to(torch.int64).sum().item())

    def timed(fn):
        for _ in range(3):
            fn()
        dist.barrier()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(args.iters):
            fn()
        e1.record(st)
        dist.barrier()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / args.iters], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / 1000.0

    t_res = timed(lambda: compute(mine_lanes))

    def both():
        x, _ = scatter()
        compute(x)

    t_inc = timed(both)
    summary = sharding.allgather_summary([k * units_per_lane, checksum(), rank])
    if rank == 0:
        total = int(summary[:, 0].sum())
        assert total == n_lanes * units_per_lane and sorted(summary[:, 2].tolist()) == list(range(world))
        print(json.dumps({"config": args.config, "n_gpus": world, "lanes": n_lanes, "lanes_per_rank": [int(v) for v in summary[:, 0] // units_per_lane],
                          "unit": unit, "resident": total / t_res, "scatter_inclusive": total / t_inc, "ms_resident": t_res * 1e3, "ms_scatter_inclusive": t_inc * 1e3,
                          "checksum": int(summary[:, 1].sum()), "scaling": "strong (fixed job, lanes sharded)",
                          "collectives": "NCCL send/recv scatter from rank 0 + all_gather of per-rank summaries; no data-path collective"}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
