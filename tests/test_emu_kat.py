"""Runs the reference's known-answer tests against the product's *kernel sources* executed by the
fiber-based CUDA emulator (tests/emu/cuda_emu.h).  This is a development aid for a GPU-less container:
it proves kernel logic (indexing, barriers, compaction, host state machines) before GPU minutes are
spent.  It says nothing about the CUDA build itself — tests/test_gpu_*.py (-m gpu) do that."""
import pytest

from tests import kat

# KATs whose signals are long (seconds of audio through a sequential emulated thread) are kept for the GPU suite.
SLOW = {"kat_short_term_matches_analytic_k_weighted_sine", "kat_lfe_and_surround_channel_weights"}
KATS = [getattr(kat, n) for n in sorted(dir(kat)) if n.startswith("kat_") and n not in SLOW]


@pytest.mark.parametrize("fn", KATS, ids=[f.__name__ for f in KATS])
def test_kat_emulated(emu, fn):
    fn(emu)
