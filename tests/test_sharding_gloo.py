"""world_size-2 gloo test of the multi-GPU host logic: lane partition, point-to-point scatter from the
ingest rank, per-rank summaries, parity-only gather.  Compute on each rank is the oracle (no GPU here)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_lanes, q):
    import torch
    import torch.distributed as dist

    from openmeters_b200 import _capi as capi
    from openmeters_b200 import sharding, synth
    from openmeters_b200.processors import SpectrogramConfig
    from oracle import oracle_py

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = SpectrogramConfig(fft_size=256, hop_size=64, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    samples = 2048
    all_lanes = torch.from_numpy(synth.cfg2_lanes(n_lanes, samples / 48000.0)[:, :samples]) if rank == 0 else None

    def compute(lanes):
        return oracle_py.stft_batch(cfg, lanes, threads=1)[1]

    counts, mine, summary = sharding.run_sharded(compute, all_lanes, n_lanes, samples)
    full = sharding.gather_columns(counts, mine, n_lanes, dst=0)
    q.put((rank, mine, summary.tolist(), None if full is None else full.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_lanes", [5, 8])
def test_two_rank_sharding(n_lanes):
    import torch.multiprocessing as mp

    from openmeters_b200 import _capi as capi
    from openmeters_b200 import sharding, synth
    from openmeters_b200.processors import SpectrogramConfig
    from oracle import oracle_py

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_lanes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # partition: disjoint, complete, round-robin
    assert res[0][1] == sharding.lanes_for_rank(n_lanes, 0, 2) and res[1][1] == sharding.lanes_for_rank(n_lanes, 1, 2)
    assert sorted(res[0][1] + res[1][1]) == list(range(n_lanes))
    # both ranks see the same summary; totals equal the single-process run
    assert res[0][2] == res[1][2]
    cfg = SpectrogramConfig(fft_size=256, hop_size=64, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    lanes = synth.cfg2_lanes(n_lanes, 2048 / 48000.0)[:, :2048]
    ref = oracle_py.stft_batch(cfg, lanes, threads=1)[1]
    summ = np.array(res[0][2])
    assert summ[:, 0].sum() == ref.size and summ[:, 1].sum() == int(ref.astype(np.int64).sum())
    assert np.array_equal(np.array(res[0][3]), ref)  # parity gather on rank 0 in global lane order
