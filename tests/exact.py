"""Error of an f32 implementation (oracle, emulator or CUDA) against the float64 restatements of tests/ref_f64.py,
broken down by level class, and the acceptance rule built on it:

    flat SURVEY §8c tolerance                         wherever the f32 CPU oracle itself meets it against exact math
    err(impl vs f64) <= K * err(oracle vs f64)        per level class, everywhere else

The second line is what "parity with an f32 reference whose FFT bits are unobtainable" can mean at most: the
implementation is no further from the mathematics the reference states than a straightforward f32 CPU implementation
of it is.  K_MAX / K_RMS absorb the sampling noise of comparing two maxima / two RMS values over a finite class.
"""
from __future__ import annotations

import numpy as np

from tests import ref_f64

REL = 1e-5
K_MAX, K_RMS = 3.0, 1.5
CLASS_DB = 10  # level classes: [0,-10), [-10,-20), ... dB re the column peak (float64 power); SURVEY evaluates down to -60
N_CLASS = 6


class Table:
    """max / sum-of-squares / count per (class, field)."""
    FIELDS = ("power", "freq", "time")

    def __init__(self):
        self.mx = np.zeros((N_CLASS, 3))
        self.ss = np.zeros((N_CLASS, 3))
        self.n = np.zeros(N_CLASS, np.int64)
        self.unaligned = 0
        self.columns = 0
        self.membership_diff = 0

    def add(self, cls, errs):
        for c in range(N_CLASS):
            m = cls == c
            k = int(m.sum())
            if not k:
                continue
            self.n[c] += k
            for i, e in enumerate(errs):
                self.mx[c, i] = max(self.mx[c, i], float(e[m].max()))
                self.ss[c, i] += float(np.sum(e[m] ** 2))

    def rms(self):
        return np.sqrt(self.ss / np.maximum(self.n, 1)[:, None])

    def to_json(self):
        r = self.rms()
        return {"columns": self.columns, "unaligned_columns": self.unaligned, "membership_differences": self.membership_diff,
                "classes": [{"level_db": [-CLASS_DB * c, -CLASS_DB * (c + 1)], "points": int(self.n[c]),
                             **{f + "_max": float(self.mx[c, i]) for i, f in enumerate(self.FIELDS)},
                             **{f + "_rms": float(r[c, i]) for i, f in enumerate(self.FIELDS)}} for c in range(N_CLASS)]}


def reassigned_table(points, counts, lanes, *, n, hop, kind, sr, zp=1, chunk=128, table=None) -> Table:
    """points (L,F,stride,3), counts (L,F) of an implementation; lanes (L,S) the f32 input.  Errors per class:
    power relative to max(p, peak*1e-3) (the SURVEY power rule's denominator), freq in Hz, time in hops."""
    t = table or Table()
    L, F = counts.shape
    for l in range(L):
        for f0 in range(0, F, chunk):
            d = ref_f64.reassigned_dense(lanes[l], n, hop, kind, sr, zp, first=f0, count=min(chunk, F - f0))
            for f in range(d["power"].shape[0]):
                p = points[l, f0 + f, : counts[l, f0 + f]].astype(np.float64)
                t.columns += 1
                idx = ref_f64.align_points(p, d, f)
                if idx is None:
                    t.unaligned += 1
                    continue
                t.membership_diff += int(np.sum(d["keep"][f])) != p.shape[0]
                if not idx.size:
                    continue
                rp = d["power"][f][idx]
                peak = float(d["power"][f].max())
                cls = np.floor(-10.0 * np.log10(np.maximum(rp / peak, 1e-30)) / CLASS_DB).astype(int)
                cls = np.where(cls < 0, 0, cls)
                ep = np.abs(p[:, 2] - rp) / np.maximum(rp, peak * 1e-3)
                ef = np.abs(p[:, 1] - d["freq"][f][idx])
                et = np.abs(p[:, 0] - d["time"][f][idx])
                t.add(cls, (ep, ef, et))
    return t


def flat_tolerances(*, n, hop, sr):
    """power 1e-5 (of max(p, peak*1e-3)), freq 1e-5*sr/2, time 1e-5*N/hop."""
    return np.array([REL, REL * sr / 2.0, REL * (n / hop)])


def assert_reassigned(impl: Table, oracle: Table, flat, what=""):
    """flat tolerance where the oracle meets it; K * oracle error elsewhere.  The assertion message carries both tables."""
    ri, ro = impl.rms(), oracle.rms()
    msg = lambda: f"{what}\nimpl   {impl.to_json()}\noracle {oracle.to_json()}\nflat {flat.tolist()}"
    assert impl.unaligned <= max(2, impl.columns // 200), msg()
    for c in range(N_CLASS):
        if impl.n[c] == 0:
            continue
        for i in range(3):
            allow = max(flat[i], K_MAX * oracle.mx[c, i])
            assert impl.mx[c, i] <= allow, (c, Table.FIELDS[i], impl.mx[c, i], allow, msg())
            if impl.n[c] >= 200 and oracle.n[c] >= 200:
                allow_rms = max(flat[i] / 3.0, K_RMS * ro[c, i])
                assert ri[c, i] <= allow_rms, (c, Table.FIELDS[i], "rms", ri[c, i], allow_rms, msg())


# ------------------------------------------------------------------------------------------------ classic
CODES_PER_LN = 4.342944819 * 65535.0 / 156.0  # d(code)/d(ln power)


def classic_stats(codes, lanes, *, n, hop, kind, zp=1, chunk=512):
    """codes (L,F,bins) u16.  A code is round(code_float); with a relative power error e the code moves by CODES_PER_LN*e,
    so excess = (|code - code_float| - 0.5) / (CODES_PER_LN * REL * max(1, peak*1e-3/p)) <= 1 is the SURVEY power rule
    seen through the quantiser.  Returns worst excess, exact fraction vs the float64 rounding on strong bins, max |diff|."""
    L, F, _ = codes.shape
    worst, exact, strong_n, maxdiff = 0.0, 0, 0, 0
    for l in range(L):
        for f0 in range(0, F, chunk):
            p, cf = ref_f64.classic_db(lanes[l], n, hop, kind, zp, first=f0, count=min(chunk, F - f0))
            c = codes[l, f0:f0 + p.shape[0]].astype(np.float64)
            peak = p.max(axis=1, keepdims=True)
            widen = np.maximum(1.0, peak * 1e-3 / np.maximum(p, 1e-300))
            cf = np.clip(cf, 0.0, 65535.0)
            ex = (np.abs(c - cf) - 0.5) / (CODES_PER_LN * REL * widen)
            # bins at the -140 dB floor in exact math: the code must be the floor code (or within the rule of it)
            worst = max(worst, float(ex.max()))
            strong = p >= peak * 1e-3
            exact += int(np.sum((c == np.round(cf))[strong]))
            strong_n += int(strong.sum())
            maxdiff = max(maxdiff, int(np.max(np.abs(c - np.round(cf))[strong], initial=0)))
    return dict(worst_excess=worst, exact_strong=exact / max(strong_n, 1), max_diff_strong=maxdiff)


# ------------------------------------------------------------------------------------------------ spectrum
def spectrum_stats(weighted, raw, lanes, *, n, hop, kind, sr, mode, param, floor_db):
    """weighted/raw (L,H,bins) f32 dB.  SURVEY power rule |pa - pb| <= REL * max(pb, hop_peak*1e-3) on the quantity the
    transform produces, i.e. the (smoothed) linear power: the weighted trace is that power times a per-bin constant
    (one f32 add in dB), so it is judged with the same per-bin relative allowance REL * max(1, peak*1e-3 / p) — judging
    it against the WEIGHTED trace's own peak would tighten the rule by up to the A-weighting of the dominant component
    (-19 dB at 100 Hz) for noise that is additive in the unweighted spectrum.  Bins on the floor in either trace are
    excluded (membership is counted separately).  Returns the worst ratios and the floor-membership mismatch rate."""
    worst = {"raw": 0.0, "weighted": 0.0}
    mism, tot = 0, 0
    for l in range(weighted.shape[0]):
        ref = ref_f64.spectrum_traces(lanes[l], n, hop, kind, sr, mode, param, floor_db)
        p = ref["power"]
        allow = REL * np.maximum(1.0, p.max(axis=-1, keepdims=True) * 1e-3 / np.maximum(p, 1e-300))
        for name, a in (("raw", raw[l]), ("weighted", weighted[l])):
            b = ref[name]
            a64 = a.astype(np.float64)
            sel = (b > floor_db + 1e-3) & (a64 > floor_db + 1e-3)
            err = np.abs(10.0 ** ((a64 - b) / 10.0) - 1.0) / allow
            if sel.any():
                worst[name] = max(worst[name], float(err[sel].max()))
            edge = np.abs(b - floor_db) < 2e-3
            mism += int(np.sum(((a64 == floor_db) != (b == floor_db)) & ~edge))
            tot += a.size
    return dict(worst_raw=worst["raw"], worst_weighted=worst["weighted"], floor_mismatch=mism / max(tot, 1))


# ------------------------------------------------------------------------------------------------ loudness
def loudness_stats(arrays: dict, x, channels, positions, sr, block_frames, floor_db=-99.9):
    """arrays: batch.snapshots_to_arrays output for ONE stream.  Max |dB| error per field against float64."""
    ref = ref_f64.loudness_snapshots(x, channels, positions, sr, block_frames, floor_db)
    out = {}
    for k in ("short_term", "momentary", "rms_fast", "rms_slow", "true_peak"):
        a = arrays[k].astype(np.float64)
        b = ref[k]
        if a.ndim == 2:
            a = a[:, :channels]
        out[k] = float(np.max(np.abs(a - b)))
    return out
