"""Parity cases shared by the emulated (CPU) and the real (GPU) suites: implementation vs oracle on the
same seeded input bytes."""
from __future__ import annotations

import numpy as np

from openmeters_b200 import _capi as capi
from openmeters_b200 import batch, synth
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig
from oracle import oracle_py
from tests import parity


def stft_parity(api, cfg: SpectrogramConfig, lanes: np.ndarray, kernel=capi.KERNEL_AUTO, expect_fast=None, rel=parity.REL):
    plan = batch.StftPlan(cfg, kernel=kernel, api=api)
    if expect_fast is not None:
        assert plan.is_fast == expect_fast
    F = cfg.fft_size * max(cfg.zero_padding_factor, 1)
    if cfg.use_reassignment:
        pa, ca = plan.execute_host(lanes)
        pb, cb = oracle_py.stft_batch(cfg, lanes)
        assert ca.shape == cb.shape and ca.shape[1] == plan.frames_per_lane(lanes.shape[1])
        return parity.compare_reassigned(pa, ca, pb, cb, sr=cfg.sample_rate, fft_len=F, window=cfg.fft_size, hop=cfg.hop_size, rel=rel)
    a = plan.execute_host(lanes)
    b = oracle_py.stft_batch(cfg, lanes)
    return parity.compare_classic(a, b, rel=rel)


def _peak_rows_agree(pka, pkb, db_ref, what):
    """Peak bins: exact, except where the two candidates are within 1e-4 dB of each other in the reference trace
    (FFT rounding decides between near-ties)."""
    diff = pka != pkb
    for idx in zip(*np.nonzero(diff)):
        a, b = int(pka[idx]), int(pkb[idx])
        assert a >= 0 and b >= 0, (what, idx, a, b)
        assert abs(db_ref[idx + (a,)] - db_ref[idx + (b,)]) < 1e-4, (what, idx, a, b)
    return int(diff.sum())


def spectrum_parity(api, cfg: SpectrumConfig, lanes: np.ndarray):
    plan = batch.SpectrumPlan(cfg, api=api)
    wa, ra, pka, fa, la = plan.execute_host_peaks(lanes)
    wb, rb, pkb = oracle_py.spectrum_batch(cfg, lanes)
    assert wa.shape == wb.shape
    st = parity.compare_db(ra, rb, cfg.floor_db)
    parity.compare_db(wa, wb, cfg.floor_db)
    # floor membership must agree except within 1e-3 dB of the floor edge
    edge = np.abs(rb - cfg.floor_db) < 1e-3
    assert np.array_equal((ra == cfg.floor_db) | edge, (rb == cfg.floor_db) | edge) or np.mean((ra == cfg.floor_db) != (rb == cfg.floor_db)) < 1e-3
    # row f3, default peak spec (spectrum/state.rs:106-107,134-136): A-weighted trace, 20 Hz .. last bin
    freqs = oracle_py.spectrum_frequency_bins(cfg.sample_rate, cfg.fft_size)
    pkb2, fb, lb = oracle_py.spectrum_peaks(freqs, wb)
    assert np.array_equal(pkb, pkb2)
    st["peak_mismatch"] = _peak_rows_agree(pka, pkb, wb, "default")
    # the fused arg-max must be the reference's peak_bin of the implementation's OWN trace, exactly
    own_pk, own_f, own_l = oracle_py.spectrum_peaks(freqs, wa)
    assert np.array_equal(pka, own_pk)
    # interpolated_peak (state.rs:327-356) is plain f32 arithmetic on three dB values: bit-exact on identical inputs
    assert np.array_equal(fa.view(np.uint32), own_f.view(np.uint32)) and np.array_equal(la.view(np.uint32), own_l.view(np.uint32))
    # ... and close to the oracle's end-to-end result wherever both picked the same bin
    same = (pka == pkb) & (pka >= 0)
    if same.any():
        bin_hz = float(freqs[1] - freqs[0])
        # the parabola vertex amplifies dB differences by 1/curvature: 0.02 bin / 0.01 dB are far below a display step
        assert np.max(np.abs(fa[same] - fb[same])) <= 0.02 * bin_hz and np.max(np.abs(la[same] - lb[same])) <= 1e-2
    # a non-default spec: raw trace, candidates restricted to [300 Hz, 5 kHz]
    plan.set_peak_spec(trace=1, min_hz=300.0, max_hz=5000.0)
    assert plan.peak_spec() == (1, 300.0, 5000.0)
    _, ra2, pk2, f2, l2 = plan.execute_host_peaks(lanes)
    assert np.array_equal(ra2, ra)
    e_pk, e_f, e_l = oracle_py.spectrum_peaks(freqs, ra2, 300.0, 5000.0)
    assert np.array_equal(pk2, e_pk)
    assert np.array_equal(f2.view(np.uint32), e_f.view(np.uint32)) and np.array_equal(l2.view(np.uint32), e_l.view(np.uint32))
    inb = pk2 >= 0
    assert np.all((freqs[pk2[inb]] >= 300.0) & (freqs[pk2[inb]] <= 5000.0))
    o_pk, _, _ = oracle_py.spectrum_peaks(freqs, rb, 300.0, 5000.0)
    st["peak_mismatch_raw_range"] = _peak_rows_agree(pk2, o_pk, rb, "raw 300-5000")
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


# ---------------------------------------------------------------- the product's config space
def settings_grid():
    """FFT sizes x hop divisors x zero-padding x windows x mode as the settings UI offers them (ui/settings.rs:146-147,
    193-200; ui/settings/spectrogram.rs:13), thinned to a grid that touches every value of every axis at least twice."""
    sizes = [1024, 2048, 4096, 8192, 16384]
    divs = [4, 6, 8, 16, 32, 64, 128]
    zps = [1, 2, 4, 8, 16, 32]
    out = []
    k = 0
    for si, n in enumerate(sizes):
        for di, d in enumerate(divs):
            if (si + di) % 2:      # checkerboard over (size, divisor)
                continue
            zp = zps[k % len(zps)]
            if n * zp > (1 << 17):  # keep the oracle's CPU time bounded: F <= 131072
                zp = max(1, (1 << 17) // n)
            out.append((n, n // d, zp, k % 5, bool(k & 1)))
            k += 1
    return out


def settings_grid_case(api, n, hop, zp, window, reassign):
    cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=window, use_reassignment=reassign, zero_padding_factor=zp)
    need = (2 * n if reassign else n) + 5 * hop  # six columns per lane
    lanes = synth.cfg2_lanes(2, (need + 16) / 48000.0)[:, :need]
    # The 1e-5 rule of SURVEY §8c is stated for the BASELINE sizes (transforms up to 2^14).  Rounding noise of an f32
    # transform grows with its length (more stages, and for dense spectra more energy spread over more bins), so beyond
    # 2^14 the tolerance scales with sqrt(F / 2^14): 2.8e-5 at the 2^17 points that 16384 x 8 zero padding reaches
    # (measured on a B200: 1.3e-5 at -31 dB re the column peak for a rectangular-window column of that size).
    F = n * zp
    rel = parity.REL * max(1.0, (F / 16384.0) ** 0.5)
    st = stft_parity(api, cfg, lanes, rel=rel)
    if reassign:
        assert st["cols"] == 12 and st["checked"] > 50, st
    else:
        assert st["exact"] >= 0.98, st
    return st
