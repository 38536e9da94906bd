"""numpy restatement of the spectrogram view's splat accumulation + resolve (row f2 of SURVEY.md §8).

TEST INFRASTRUCTURE ONLY.  Follows render/shaders/spectrogram.wgsl:126-147 (vs_accum_splat), :215-225 (fs_accum),
:227-237 (fs_resolve) and spectrogram/render.rs:205-252 (Uniforms::from_params), in float64 so that it is the
yard-stick, not a second implementation of the same rounding.  Besides the image it returns, per point, how far the
quad's edges are from the nearest pixel centre: a point closer than `edge_eps` pixels to flipping pixels is a
rasteriser tie (the GPU itself snaps vertices to 1/256 px) and the parity test treats it as "may land either side".
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LOG_KNEE_HZ = 20.0
LN_TO_DB = 4.342944819
DB_TO_LOG2 = 0.3321928095
ANALYSIS_POWER_EPS = 1.0023052e-14
DB_ANALYSIS_FLOOR = -140.0


def freq_scaled(scale: int, hz):
    hz = np.asarray(hz, np.float64)
    if scale == 1:
        return np.arcsinh(hz / LOG_KNEE_HZ)
    if scale == 2:
        return 21.4 * np.log10(1.0 + hz / 228.8)
    return hz


@dataclass
class SplatRef:
    accum: np.ndarray      # (n_rings, H, W) float64: points that are not rasteriser ties
    ties: np.ndarray       # (n_rings, H, W) float64: power of tie points spread over every pixel they may touch
    db: np.ndarray         # resolve of accum (+ties ignored), -inf where nothing accumulated
    n_drawn: int


def render(rings: np.ndarray, counts: np.ndarray, p, edge_eps: float = 1.0 / 128.0) -> SplatRef:
    """rings: (n_rings, ring_capacity, stride, 3) float32 [time_offset, freq_hz, power]; counts: (n_rings, ring_capacity);
    p: any object with the omb_splat_params fields."""
    n_rings, hl, stride, _ = rings.shape
    W = int(np.ceil(max(float(p.ext_w), 1.0)))
    H = int(np.ceil(max(float(p.ext_h), 1.0)))
    sf = max(float(p.scale_factor), 1.0)
    lo, hi = float(freq_scaled(p.freq_scale, np.float32(p.freq_min))), float(freq_scaled(p.freq_scale, np.float32(p.freq_max)))
    axis_inv = 1.0 / max(hi - lo, 1e-12)
    inv_uv = 1.0 / max(float(p.uv_y_range[1]) - float(p.uv_y_range[0]), 1e-12)
    slots = min(int(p.col_count), hl)
    newest = int(p.newest_col) % hl
    accum = np.zeros((n_rings, H, W))
    ties = np.zeros((n_rings, H, W))
    drawn = 0
    for r in range(n_rings):
        for s in range(slots):
            c = min(int(counts[r, s]), stride)
            if c == 0:
                continue
            pts = rings[r, s, :c].astype(np.float64)
            t, f, pw = pts[:, 0], pts[:, 1], pts[:, 2].copy()
            zoomed = ((freq_scaled(p.freq_scale, f) - lo) * axis_inv - float(p.uv_y_range[0])) * inv_uv
            keep = (pw > 0) & ~(zoomed < -0.01) & ~(zoomed > 1.01)
            if p.tilt_db != 0.0:
                keep &= pw > ANALYSIS_POWER_EPS
                pos_f = f > 0
                pw[pos_f] *= np.exp2(float(p.tilt_db) * np.log2(np.where(pos_f, f, 1000.0)[pos_f] / 1000.0) * DB_TO_LOG2)
            age = (newest + hl - s) % hl
            px = float(p.ext_w) - (age - t) * sf
            py = (1.0 - zoomed) * float(p.ext_h)
            h = 0.5 * sf
            # pixel i is covered iff px - h <= i + 0.5 < px + h
            ex0, ex1, ey0, ey1 = px - h - 0.5, px + h - 0.5, py - h - 0.5, py + h - 0.5
            x0, x1, y0, y1 = np.ceil(ex0), np.ceil(ex1), np.ceil(ey0), np.ceil(ey1)
            near = lambda e: np.abs(e - np.round(e)) < edge_eps
            tie = near(ex0) | near(ex1) | near(ey0) | near(ey1)
            for i in np.nonzero(keep)[0]:
                if tie[i]:
                    xa, xb = int(np.floor(ex0[i] - edge_eps)) , int(np.ceil(ex1[i] + edge_eps)) + 1
                    ya, yb = int(np.floor(ey0[i] - edge_eps)), int(np.ceil(ey1[i] + edge_eps)) + 1
                    ties[r, max(ya, 0):min(yb, H), max(xa, 0):min(xb, W)] += pw[i]
                    continue
                xa, xb, ya, yb = max(int(x0[i]), 0), min(int(x1[i]), W), max(int(y0[i]), 0), min(int(y1[i]), H)
                if xa < xb and ya < yb:
                    accum[r, ya:yb, xa:xb] += pw[i]
                    drawn += 1
    power = accum * float(p.reassigned_power_scale)
    with np.errstate(divide="ignore"):
        db = np.where(power > 0, np.maximum(np.log(np.maximum(power, 1e-20)) * LN_TO_DB, DB_ANALYSIS_FLOOR), -np.inf)
    return SplatRef(accum, ties, db, drawn)
