"""Parity cases shared by the emulated (CPU) and the real (GPU) suites: implementation vs oracle on the
same seeded input bytes."""
from __future__ import annotations

import numpy as np

from openmeters_b200 import _capi as capi
from openmeters_b200 import batch, synth
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig
from oracle import oracle_py
from tests import parity


def stft_parity(api, cfg: SpectrogramConfig, lanes: np.ndarray, kernel=capi.KERNEL_AUTO, expect_fast=None):
    plan = batch.StftPlan(cfg, kernel=kernel, api=api)
    if expect_fast is not None:
        assert plan.is_fast == expect_fast
    F = cfg.fft_size * max(cfg.zero_padding_factor, 1)
    if cfg.use_reassignment:
        pa, ca = plan.execute_host(lanes)
        pb, cb = oracle_py.stft_batch(cfg, lanes)
        assert ca.shape == cb.shape and ca.shape[1] == plan.frames_per_lane(lanes.shape[1])
        return parity.compare_reassigned(pa, ca, pb, cb, sr=cfg.sample_rate, fft_len=F, window=cfg.fft_size, hop=cfg.hop_size)
    a = plan.execute_host(lanes)
    b = oracle_py.stft_batch(cfg, lanes)
    return parity.compare_classic(a, b)


def spectrum_parity(api, cfg: SpectrumConfig, lanes: np.ndarray):
    plan = batch.SpectrumPlan(cfg, api=api)
    wa, ra, pka = plan.execute_host(lanes)
    wb, rb, pkb = oracle_py.spectrum_batch(cfg, lanes)
    assert wa.shape == wb.shape
    st = parity.compare_db(ra, rb, cfg.floor_db)
    parity.compare_db(wa, wb, cfg.floor_db)
    # floor membership must agree except within 1e-3 dB of the floor edge
    edge = np.abs(rb - cfg.floor_db) < 1e-3
    assert np.array_equal((ra == cfg.floor_db) | edge, (rb == cfg.floor_db) | edge) or np.mean((ra == cfg.floor_db) != (rb == cfg.floor_db)) < 1e-3
    # peak bin: exact, except where the two top raw values are within 1e-4 dB of each other (FFT rounding decides)
    diff = pka != pkb
    if diff.any():
        for l, h in zip(*np.nonzero(diff)):
            assert abs(rb[l, h, pka[l, h]] - rb[l, h, pkb[l, h]]) < 1e-4, (l, h, pka[l, h], pkb[l, h])
    st["peak_mismatch"] = int(diff.sum())
    return st


def loudness_parity(api, cfg: LoudnessConfig, channels: int, positions, streams: np.ndarray, block_frames: int):
    plan = batch.LoudnessPlan(cfg, channels, positions, api=api)
    sa, nb = plan.execute_host(streams, block_frames)
    sb, nb2 = oracle_py.loudness_batch(cfg, channels, positions, streams, block_frames)
    assert nb == nb2
    n = nb * streams.shape[0]
    A = batch.snapshots_to_arrays(sa, n)
    B = batch.snapshots_to_arrays(sb, n)
    out = {}
    for k in A:
        # 1e-5 relative on mean squares == 4.3e-5 dB; true peak is a max of bit-identical f32 FIR sums
        tol = 5e-5 if k != "true_peak" else 1e-5
        err = np.abs(A[k] - B[k])
        assert err.max() <= tol, (k, float(err.max()), int(np.argmax(err)))
        out[k] = float(err.max())
    for i in range(n):
        assert sa[i].channel_count == sb[i].channel_count and tuple(sa[i].positions) == tuple(sb[i].positions)
    return out
