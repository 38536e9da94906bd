"""CPU checks of bench.py's host logic (the timed paths need a GPU): lanes are a function of their GLOBAL index, so the
resident run of rank r and the scatter from rank 0 see identical bytes; workload bookkeeping matches SURVEY.md §8d."""
import argparse
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(**kw):
    base = dict(lanes=4, samples=8192 + 7 * 1024, cfg5_lanes=16, cfg5_frames=3, cfg4_seconds=1)
    base.update(kw)
    return argparse.Namespace(**base)


def test_rank_blocks_tile_the_global_lane_set():
    b = _bench()
    base = np.arange(3 * 50, dtype=np.float32).reshape(3, 50)
    whole = b.tile(base, 12, 0)
    for world in (2, 3, 4):
        k = 12 // world
        parts = [b.tile(base, k, r * k) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), whole)
    wl = b.make_workload("cfg2", _args(), 2)
    assert np.array_equal(np.concatenate([wl.host_lanes(0, 4), wl.host_lanes(4, 4)]), wl.host_lanes(0, 8))


def test_workload_bookkeeping_matches_survey_8d():
    b = _bench()
    a = _args()
    cfg2 = b.make_workload("cfg2", a, 1)
    assert cfg2.algo_bytes == 28688 and cfg2.units_per_lane() == 8 and cfg2.scaling == "weak" and cfg2.lanes_per_rank(8) == 4
    cfg5 = b.make_workload("cfg5", a, 8)
    assert cfg5.algo_bytes == 57360 and cfg5.units_per_lane() == 3 and cfg5.scaling == "strong" and cfg5.lanes_per_rank(8) == 2
    cfg1 = b.make_workload("cfg1", a, 1)
    assert cfg1.algo_bytes == 3074
    cfg4 = b.make_workload("cfg4", a, 4)
    assert cfg4.algo_bytes == 69640 and cfg4.lanes_per_rank(4) == 32 and cfg4.units_per_lane() == (48000 - 16384) // 1024 + 1
    cfg3 = b.make_workload("cfg3", a, 1)
    assert cfg3.algo_bytes == 4 and cfg3.units_per_lane() == 30 * 48000 * 8


def test_traffic_is_null_unless_captured_from_this_build(tmp_path, monkeypatch):
    b = _bench()
    t, src = b.recorded_traffic("k_no_such_kernel")
    assert t is None and src is None
