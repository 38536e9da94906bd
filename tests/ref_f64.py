"""float64 restatements of the four hot paths, written from the reference's Rust (not from the C++ oracle).

TEST INFRASTRUCTURE.  The reference's FFT arithmetic lives in rustfft / realfft, which cannot run here, so no f32
implementation can be compared with the reference bit for bit.  What CAN be pinned is the distance of every f32
implementation (the C++ oracle, the CUDA kernels, and by the same argument rustfft) from the exact mathematics the Rust
source states.  These functions evaluate that mathematics in float64 (numpy pocketfft, scipy for the IIR) on the same
f32 input bytes; `tests/test_exact_math.py` (CPU) measures the oracle against them and `tests/test_gpu_exact.py`
(-m gpu) requires the CUDA path to be as close to them as the f32 oracle is, per level class.

Citations are to /root/reference/src (commit 8f09203c).
"""
from __future__ import annotations

import numpy as np

# util/audio/window.rs:25-31
WINDOW_COEFFS = {0: [1.0], 1: [0.5, -0.5], 2: [25.0 / 46.0, -21.0 / 46.0], 3: [0.42, -0.5, 0.08],
                 4: [0.35875, -0.48829, 0.14128, -0.01168]}
LN_TO_DB = float(np.float32(4.342_944_8))  # util/audio/level.rs:5 (the f32 constant, as the reference multiplies by it)
DB_FLOOR = -140.0                          # level.rs:4
ANALYSIS_FLOOR_POWER = float(np.float32(1e-14))  # spectrogram/processor.rs:69
FLT_MIN = float(np.finfo(np.float32).tiny)


def window(kind: int, n: int) -> np.ndarray:
    """util/audio/window.rs:20-43: periodic cosine-sum window, phi = n * TAU / len."""
    if n <= 1 or kind == 0:
        return np.ones(n)
    phi = np.arange(n) * (2.0 * np.pi / n)
    return sum(c * np.cos(phi * k) for k, c in enumerate(WINDOW_COEFFS[kind]))


def bin_norm(w: np.ndarray, fft_size: int) -> np.ndarray:
    """util/audio/window.rs:90-109."""
    s = float(np.sum(w))
    inv = 1.0 / s if abs(s) > float(np.finfo(np.float32).eps) else 1.0 / fft_size
    norm = np.full(fft_size // 2 + 1, 4.0 * inv * inv)
    norm[0] = inv * inv
    if fft_size % 2 == 0 and norm.size > 1:
        norm[-1] = inv * inv
    return norm


def frames_view(x: np.ndarray, length: int, hop: int, first: int = 0, count: int | None = None) -> np.ndarray:
    """(frames, length) float64 copies of x[j*hop : j*hop + length]."""
    total = (x.size - length) // hop + 1 if x.size >= length else 0
    count = total - first if count is None else min(count, total - first)
    if count <= 0:
        return np.zeros((0, length))
    v = np.lib.stride_tricks.as_strided(x[first * hop:], shape=(count, length), strides=(hop * x.strides[0], x.strides[0]))
    return v.astype(np.float64)


def power_to_db(p: np.ndarray, floor: float) -> np.ndarray:
    """level.rs:28-34."""
    with np.errstate(divide="ignore"):
        return np.where(p > 0, np.maximum(np.log(np.maximum(p, 1e-300)) * LN_TO_DB, floor), floor)


# ---------------------------------------------------------------------------------------------------- classic
def classic_db(lane: np.ndarray, n: int, hop: int, kind: int, zp: int = 1, first: int = 0, count: int | None = None):
    """spectrogram/processor.rs:349-380 up to (not including) the u16 rounding: returns (power, code_float) per
    (frame, bin); the reference's code is round(code_float) clamped to [0, 65535] (processor.rs:103-108)."""
    w = window(kind, n)
    fr = frames_view(lane, n, hop, first, count)
    mean = fr.sum(axis=1, keepdims=True) / n          # window.rs:80-84
    r = (fr - mean) * w
    X = np.fft.rfft(r, n * zp, axis=1)
    p = (X.real ** 2 + X.imag ** 2) * bin_norm(w, n * zp)
    db = power_to_db(p, DB_FLOOR)
    return p, (db + 144.0) * (65535.0 / 156.0)


# ---------------------------------------------------------------------------------------------------- reassigned
def reassignment_windows(kind: int, n: int):
    """spectrogram/processor.rs:569-608: dh = Re IFFT(j w FFT(h)) / n with bins 0 and n/2 zeroed; th = (i - (n-1)/2) h."""
    h = window(kind, n)
    k = np.arange(n)
    omega = (2.0 * np.pi / n) * (k - np.where(k > n // 2, n, 0))
    W = np.fft.fft(h)
    W[0] = 0
    if n % 2 == 0:
        W[n // 2] = 0
    dh = np.real(np.fft.ifft(1j * omega * W))
    th = (np.arange(n) - (n - 1) * 0.5) * h
    return h, dh, th


def reassigned_dense(lane: np.ndarray, n: int, hop: int, kind: int, sr: float, zp: int = 1, first: int = 0,
                     count: int | None = None):
    """spectrogram/processor.rs:318-348,439-488,546-557 for frames [first, first+count): per (frame, bin) arrays
    power (= pow * bin_norm / H^2), freq, time and the keep mask, before compaction."""
    H = max(2, 1 << (2 * n - 1).bit_length())
    F = n * zp
    off = (H - n) // 2
    h, dh, th = reassignment_windows(kind, n)
    fr = frames_view(lane, H, hop, first, count)
    A = np.fft.fft(fr, axis=1)
    A[:, 0] = 0
    A[:, H // 2 + 1:] = 0
    a = np.fft.ifft(A, axis=1) * H                    # unnormalised inverse (processor.rs:556)
    c = a[:, off:off + n]
    S = np.fft.fft(c * h, F, axis=1)[:, : F // 2 + 1]
    D = np.fft.fft(c * dh, F, axis=1)[:, : F // 2 + 1]
    T = np.fft.fft(c * th, F, axis=1)[:, : F // 2 + 1]
    norm = bin_norm(h, F) / float(H) ** 2             # processor.rs:263-266
    pw = S.real ** 2 + S.imag ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        d_omega = -(D.imag * S.real - D.real * S.imag) / pw
        freq = np.arange(F // 2 + 1) * (sr / F) + d_omega * (sr / (2.0 * np.pi))
        time = (T.real * S.real + T.imag * S.imag) / pw / hop - off / hop
    power = pw * norm
    keep = (power >= ANALYSIS_FLOOR_POWER) & (freq > 0) & (sr * 0.5 - freq > 0)
    return dict(power=power, freq=freq, time=time, keep=keep)


def align_points(points: np.ndarray, dense: dict, f: int):
    """Bin index of every compacted point of frame f (ascending-bin order) given the float64 dense column, or None if
    the column cannot be aligned.  Membership may differ from the float64 keep mask only on bins that sit on a
    decision threshold (power at 1e-14, frequency at 0 or sr/2); those are resolved by best power match."""
    keep, power = dense["keep"][f], dense["power"][f]
    idx = np.nonzero(keep)[0]
    n = points.shape[0]
    if n == idx.size:
        return idx
    # candidates: bins whose decision is marginal in exact arithmetic
    fr = dense["freq"][f]
    nyq = np.nanmax(fr[np.isfinite(fr)]) if np.isfinite(fr).any() else 0.0
    marginal = np.zeros_like(keep)
    with np.errstate(invalid="ignore"):
        marginal |= np.abs(power - ANALYSIS_FLOOR_POWER) <= 1e-2 * ANALYSIS_FLOOR_POWER
        marginal |= np.abs(fr) <= 1.0
        marginal |= np.abs(fr - nyq) <= 1.0
    marginal[:3] = True
    marginal[-3:] = True
    cand = np.nonzero(marginal)[0]
    if cand.size > 14:
        return None
    base = keep & ~marginal
    best, best_err = None, np.inf
    need = n - int(base.sum())
    if need < 0 or need > cand.size:
        return None
    from itertools import combinations
    for sub in combinations(cand.tolist(), need):
        m = base.copy()
        m[list(sub)] = True
        ii = np.nonzero(m)[0]
        err = float(np.max(np.abs(points[:, 2] - power[ii]) / np.maximum(power[ii], 1e-300))) if n else 0.0
        if err < best_err:
            best, best_err = ii, err
    return best if best_err < 1e-2 else None


# ---------------------------------------------------------------------------------------------------- spectrum
def a_weight_db(freq_hz: np.ndarray) -> np.ndarray:
    """spectrum/processor.rs:410-425 (f64 arithmetic as in the reference; -inf at f <= 0)."""
    c1, c2, c3, c4 = 20.598_997 ** 2, 107.652_65 ** 2, 737.862_23 ** 2, 12_194.217 ** 2
    f2 = np.asarray(freq_hz, np.float64) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        ra = c4 * f2 * f2 / ((f2 + c1) * np.sqrt((f2 + c2) * (f2 + c3)) * (f2 + c4))
        out = 20.0 * np.log10(ra) + 2.0
    return np.where(np.asarray(freq_hz) <= 0, -np.inf, out)


def spectrum_traces(lane: np.ndarray, n: int, hop: int, kind: int, sr: float, mode: int, param: float, floor_db: float):
    """spectrum/processor.rs:179-253,332-402: per hop [weighted, raw] dB traces (hops, bins) and the smoothed power.
    mode: 0 None, 1 Exponential{factor}, 2 PeakHold{decay_per_second} (the C ABI's numbering)."""
    w = window(kind, n)
    fr = frames_view(lane, n, hop)
    mean = fr.sum(axis=1, keepdims=True) / n
    X = np.fft.rfft((fr - mean) * w, n, axis=1)
    p = (X.real ** 2 + X.imag ** 2) * bin_norm(w, n)
    bins = n // 2 + 1
    freqs = np.arange(bins) * float(np.float32(sr) / np.float32(n))   # processor.rs:138-146 (f32 bin_hz)
    aw = a_weight_db(freqs.astype(np.float32)).astype(np.float32).astype(np.float64)  # stored as f32 (processor.rs:424)
    headroom = max(0.0, float(np.max(aw)))
    state_floor = max(10.0 ** ((floor_db - headroom) * 0.1), FLT_MIN)  # processor.rs:332-336
    powers = np.empty_like(p)
    if mode == 0:
        powers = p
    else:
        st = np.zeros(bins)
        if mode == 1:
            alpha = min(max(float(np.float32(param)), 0.0), float(np.float32(0.9999)))
        else:
            decay = 10.0 ** (-max(param, 0.0) * (hop / sr) * 0.1)
        for j in range(p.shape[0]):
            if mode == 1:
                st = np.where(st <= 0.0, p[j], st * alpha + p[j] * (1.0 - alpha))
            else:
                st = np.maximum(st * decay, p[j])
            st = np.where(st < state_floor, 0.0, st)
            powers[j] = st
    below = powers < state_floor
    with np.errstate(divide="ignore"):
        db = np.log(np.maximum(powers, 1e-300)) * LN_TO_DB
    raw = np.where(below, floor_db, np.maximum(db, floor_db))
    weighted = np.where(below, floor_db, np.maximum(db + aw, floor_db))
    return dict(weighted=weighted, raw=raw, power=powers, state_floor=state_floor, a_weight=aw)


# ---------------------------------------------------------------------------------------------------- loudness
def k_weighting(fs: float):
    """loudness/processor.rs:22-55."""
    f0, g, q = 1_681.974_450_955_533, 3.999_843_853_973_347, 0.707_175_236_955_419_6
    k = np.tan(np.pi * f0 / fs)
    vh = 10.0 ** (g / 20.0)
    vb = vh ** 0.499_666_774_154_541_6
    a0 = 1.0 + k / q + k * k
    pb = [(vh + vb * k / q + k * k) / a0, 2.0 * (k * k - vh) / a0, (vh - vb * k / q + k * k) / a0]
    pa = [1.0, 2.0 * (k * k - 1.0) / a0, (1.0 - k / q + k * k) / a0]
    f0, q = 38.135_470_876_024_44, 0.500_327_037_323_877_3
    k = np.tan(np.pi * f0 / fs)
    a0 = 1.0 + k / q + k * k
    rb = [1.0, -2.0, 1.0]
    ra = [1.0, 2.0 * (k * k - 1.0) / a0, (1.0 - k / q + k * k) / a0]
    return np.convolve(pb, rb), np.convolve(pa, ra)


def true_peak_firs():
    """loudness/processor.rs:80-97 (coefficients are stored as f32)."""
    def coef(j, factor):
        offset = j - 24.0
        win = 0.5 * (1.0 - np.cos(2.0 * np.pi * j / 48.0))
        x = offset * np.pi / factor
        return float(np.float32(win * np.sin(x) / x))
    fir4 = np.array([[coef(tap * 4 + ph + 1, 4) for ph in range(3)] for tap in range(12)])
    fir2 = np.array([coef(tap * 2 + 1, 2) for tap in range(24)])
    return fir4, fir2


def channel_weight(position: int) -> float:
    """loudness/processor.rs:174-183 with the C ABI's position codes (openmeters_b200/_capi.py: LFE 3, rear 4/5, side 6/7)."""
    from openmeters_b200 import _capi as capi
    if position == capi.POS_LOW_FREQUENCY:
        return 0.0
    if position in (capi.POS_REAR_LEFT, capi.POS_REAR_RIGHT, capi.POS_SIDE_LEFT, capi.POS_SIDE_RIGHT):
        return 1.41
    return 1.0


def loudness_snapshots(x: np.ndarray, channels: int, positions, sr: float, block_frames: int, floor_db: float = -99.9):
    """loudness/processor.rs:253-311 per block of `block_frames` frames: exact sliding-window means (float64 / long
    double prefix sums instead of the compensated running sums, which approximate exactly these means), the K-weighting
    filter in float64 (scipy.signal.lfilter is the same transposed direct form II), y rounded to f32 (processor.rs:161)."""
    from scipy.signal import lfilter

    sr32 = float(np.float32(sr))
    b, a = k_weighting(sr32)
    x = np.asarray(x, np.float32).reshape(-1, channels)
    frames = x.shape[0]
    n_blocks = (frames + block_frames - 1) // block_frames
    caps = [max(1, int(np.float32(sr32) * np.float32(s))) for s in (3.0, 0.4, 0.3, 1.0)]  # processor.rs:68-71
    ends = np.minimum((np.arange(n_blocks) + 1) * block_frames, frames)
    ms = np.zeros((4, n_blocks, channels))
    peak = np.zeros((n_blocks, channels))
    fir4, fir2 = true_peak_firs()
    for c in range(channels):
        xc = x[:, c].astype(np.float64)
        nz = np.nonzero(x[:, c].view(np.uint32))[0]
        if nz.size == 0:
            continue
        y = lfilter(b, a, xc).astype(np.float32).astype(np.float64)
        csum = np.concatenate([[0.0], np.cumsum((y * y).astype(np.longdouble))])
        for wi, cap in enumerate(caps):
            lo = np.maximum(ends - cap, 0)
            cnt = np.maximum(np.minimum(ends, cap), 1)
            ms[wi, :, c] = ((csum[ends] - csum[lo]) / cnt).astype(np.float64)
        # blocks that end before the channel's first non-zero sample: the channel is not active yet (processor.rs:264-274)
        inactive = ends <= nz[0]
        ms[:, inactive, c] = np.nan
        # true peak (processor.rs:123-150): max |x| and |polyphase outputs| over the block
        ax = np.abs(xc)
        cand = [ax]
        if sr32 < 96000.0:
            for ph in range(3):
                cand.append(np.abs(lfilter(fir4[:, ph], [1.0], xc)))
        elif sr32 < 192000.0:
            cand.append(np.abs(lfilter(fir2, [1.0], xc)))
        m = np.max(np.stack(cand), axis=0)
        m[: nz[0]] = 0.0
        starts = np.arange(n_blocks) * block_frames
        peak[:, c] = np.maximum.reduceat(m, starts)
        peak[inactive, c] = np.nan
    if positions is None:  # dsp.rs:36-47 fallback layout
        pos = list(range(channels))
        if channels == 1:
            pos = [8]
        elif channels == 4:
            pos[2:4] = [4, 5]
        elif channels == 5:
            pos[3:5] = [4, 5]
    else:
        pos = list(positions)
    weights = np.array([channel_weight(pos[c]) for c in range(channels)])

    def lufs(msq):
        tot = np.nansum(msq * weights, axis=1)
        with np.errstate(divide="ignore"):
            return np.where(tot > 0, np.maximum(10.0 * np.log10(np.maximum(tot, 1e-300)) - 0.691, floor_db), floor_db)

    def db(v):
        out = power_to_db(np.nan_to_num(v, nan=0.0), floor_db)
        return np.where(np.isnan(v), floor_db, out)

    return dict(short_term=lufs(ms[0]), momentary=lufs(ms[1]), rms_fast=db(ms[2]), rms_slow=db(ms[3]),
                true_peak=db(peak * peak), mean_squares=ms)
