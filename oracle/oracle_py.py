"""ctypes binding of the CPU oracle (oracle/libomb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg; never by openmeters_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

from openmeters_b200 import _capi as capi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libomb_oracle.so")

_vp, _u32, _u64, _sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_size_t
_f32p, _f64p, _u8p = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint8)

# oracle-only symbols (test hooks + batch entry points with a thread count)
EXTRA = {
    "hw_threads": (C.c_int, []),
    "power_to_db": (C.c_float, [C.c_float, C.c_float]),
    "db_to_power": (C.c_float, [C.c_float]),
    "sanitize_sample_rate": (C.c_float, [C.c_float]),
    "history_columns": (_sz, [C.c_int, _u32, _sz]),
    "fft": (C.c_int, [_f32p, _sz, C.c_int]),
    "windowed_means": (C.c_int, [C.POINTER(_sz), _sz, _sz, _f64p, _sz, _f64p]),
    "stereo_frames": (C.c_int, [_f32p, _sz, _u32, _u8p, _f32p]),
    "spectrogram_pending": (_sz, [_vp, _f32p, _sz]),
    "spectrum_peek": (C.c_int, [_vp, C.POINTER(capi.SpectrumSnapshot)]),
    "spectrum_pending": (_sz, [_vp, C.c_int, _f32p, _sz]),
    "spectrum_update_outputs": (C.c_int, [C.c_int, C.c_float, C.c_float, _f32p, _sz, C.c_float, C.c_float,
                                          _f32p, _f32p, _f32p, _f32p]),
    "loudness_force_eager": (C.c_int, [_vp, _u32, C.c_float]),
    "stft_batch": (C.c_int, [C.POINTER(capi.SpectrogramConfig), _vp, _u32, _u64, _u64, _vp, _u64, _vp, _vp,
                             C.c_int, _u64, _u64]),
    "spectrum_batch": (C.c_int, [C.POINTER(capi.SpectrumConfig), _vp, _u32, _u64, _u64, _vp, _vp, _vp, C.c_int]),
    "spectrum_peaks": (C.c_int, [_vp, _vp, _u32, _u64, C.c_float, C.c_float, _vp, _vp, _vp]),
    "spectrum_interpolate_peaks": (C.c_int, [_vp, _vp, _u32, _u64, _vp, _vp, _vp]),
    "loudness_batch": (C.c_int, [C.POINTER(capi.LoudnessConfig), _u32, _u8p, _vp, _u32, _u64, _u64, _u64, _vp,
                                 C.c_int]),
}

# header symbols the oracle also implements (same signatures, ombo_ prefix)
_SHARED = [k for k in capi.HEADER_SYMBOLS if not (
    k in ("last_error", "device_count", "set_device", "kernel_launch_count", "probe_fp32_tflops", "copy_async") or k.startswith("peer_")
    or k.startswith("stft_plan")
    or k.startswith("stft_execute") or k.startswith("stft_render") or k.startswith("spectrum_plan") or k.startswith("spectrum_execute")
    or k.startswith("loudness_plan") or k.startswith("loudness_execute")
    or k in ("spectrum_default_peak_spec", "spectrum_interpolate_peaks_device")
    # host-side timeline logic: restated in Python (oracle/meter_py.py), not in the C++ oracle
    or k.startswith("timeline_") or k.startswith("meter_")
    # splat accumulation: restated in numpy (oracle/splat_py.py)
    or k.startswith("splat_")
    # the multi-stream bank is checked against S independent oracle processors
    or k.startswith("spectrogram_bank_") or k.startswith("loudness_bank_") or k.startswith("spectrum_bank_"))]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "omb_oracle.cpp")
    hdr = os.path.join(HERE, "..", "include", "omb200.h")
    stale = (not os.path.exists(LIB)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(LIB) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", HERE, "-B", "libomb_oracle.so"], check=True, capture_output=True)
    return LIB


@lru_cache(maxsize=1)
def api():
    build()
    lib = C.CDLL(LIB, mode=C.RTLD_LOCAL)
    syms = {k: capi.HEADER_SYMBOLS[k] for k in _SHARED}
    syms.update(EXTRA)
    ns = capi.bind(lib, "ombo_", syms)
    ns.last_error = lambda: b""
    return ns


# --- numpy conveniences -----------------------------------------------------
def _cfg_ptr(cfg):
    c = cfg.to_c()
    return c, C.byref(c)


def stft_batch(cfg, lanes: np.ndarray, threads: int = 0, point_stride: int | None = None,
               frame_begin: int = 0, frame_end: int | None = None, out=None):
    """lanes: (n_lanes, samples) float32. Returns (points[(L,F,stride,3)], counts[(L,F)]) or codes[(L,F,bins)]."""
    a = api()
    lanes = np.ascontiguousarray(lanes, np.float32)
    L, S = lanes.shape
    c, cp = _cfg_ptr(cfg)
    frames = int(a.stft_frames_per_lane(cp, S))
    cn = cfg.__class__(**{**cfg.__dict__})
    bins = cfg.fft_size * max(cfg.zero_padding_factor, 1) // 2 + 1
    fe = frames if frame_end is None else frame_end
    if cfg.use_reassignment:
        stride = point_stride or bins
        if out is not None:  # caller-provided (pts, cnt): timing loops reuse them so page faults are not measured
            pts, cnt = out
            assert pts.shape == (L, frames, stride, 3) and cnt.shape == (L, frames)
        else:
            pts = np.zeros((L, frames, stride, 3), np.float32)
            cnt = np.zeros((L, frames), np.uint32)
        rc = a.stft_batch(cp, lanes.ctypes.data, L, S, S, pts.ctypes.data, stride, cnt.ctypes.data, None, threads,
                          frame_begin, fe)
        assert rc == 0
        return pts, cnt
    codes = np.zeros((L, frames, bins), np.uint16)
    rc = a.stft_batch(cp, lanes.ctypes.data, L, S, S, None, 0, None, codes.ctypes.data, threads, frame_begin, fe)
    assert rc == 0
    return codes


def spectrum_batch(cfg, lanes: np.ndarray, threads: int = 0, want_peak: bool = True):
    a = api()
    lanes = np.ascontiguousarray(lanes, np.float32)
    L, S = lanes.shape
    c, cp = _cfg_ptr(cfg)
    hops = int(a.spectrum_hops_per_lane(cp, S))
    bins = cfg.fft_size // 2 + 1
    w = np.zeros((L, hops, bins), np.float32)
    r = np.zeros((L, hops, bins), np.float32)
    pk = np.zeros((L, hops), np.int32)
    rc = a.spectrum_batch(cp, lanes.ctypes.data, L, S, S, w.ctypes.data, r.ctypes.data,
                          pk.ctypes.data if want_peak else None, threads)
    assert rc == 0
    return w, r, pk


def spectrum_frequency_bins(sample_rate: float, fft_size: int) -> np.ndarray:
    """frequency_bins of a SpectrumSnapshot: bin as f32 * (sample_rate / fft_size as f32) (spectrum/processor.rs:138-146)."""
    bin_hz = np.float32(sample_rate) / np.float32(fft_size)
    return (np.arange(fft_size // 2 + 1, dtype=np.float32) * bin_hz).astype(np.float32)


def spectrum_peaks(bins_hz: np.ndarray, db: np.ndarray, min_hz: float = 20.0, max_hz: float | None = None):
    """peak_bin + interpolated_peak (spectrum/state.rs:321-356) per row of db[..., bins] -> (bin, freq_hz, level_db)."""
    a = api()
    bins_hz = np.ascontiguousarray(bins_hz, np.float32)
    db = np.ascontiguousarray(db, np.float32)
    n = db.shape[-1]
    rows = db.size // n if n else 0
    if max_hz is None or max_hz <= 0:
        max_hz = float(max(bins_hz[-1], np.float32(min_hz) * np.float32(1.02)))  # state.rs:107
    b = np.zeros(db.shape[:-1], np.int32)
    f = np.zeros(db.shape[:-1], np.float32)
    m = np.zeros(db.shape[:-1], np.float32)
    assert a.spectrum_peaks(bins_hz.ctypes.data, db.ctypes.data, n, rows, min_hz, max_hz, b.ctypes.data, f.ctypes.data, m.ctypes.data) == 0
    return b, f, m


def spectrum_interpolate_peaks(bins_hz: np.ndarray, db: np.ndarray, peak_bin: np.ndarray):
    a = api()
    bins_hz = np.ascontiguousarray(bins_hz, np.float32)
    db = np.ascontiguousarray(db, np.float32)
    pk = np.ascontiguousarray(peak_bin, np.int32)
    n = db.shape[-1]
    f = np.zeros(pk.shape, np.float32)
    m = np.zeros(pk.shape, np.float32)
    assert a.spectrum_interpolate_peaks(bins_hz.ctypes.data, db.ctypes.data, n, pk.size, pk.ctypes.data, f.ctypes.data, m.ctypes.data) == 0
    return f, m


def loudness_batch(cfg, channels: int, positions, streams: np.ndarray, block_frames: int, threads: int = 0):
    """streams: (n_streams, frames*channels) interleaved float32 -> array of capi.LoudnessSnapshot (S, n_blocks)."""
    a = api()
    streams = np.ascontiguousarray(streams, np.float32)
    S, n = streams.shape
    frames = n // channels
    n_blocks = (frames + block_frames - 1) // block_frames
    out = (capi.LoudnessSnapshot * (S * n_blocks))()
    c = capi.LoudnessConfig(cfg.sample_rate, cfg.floor_db)
    rc = a.loudness_batch(C.byref(c), channels, capi.positions_array(positions), streams.ctypes.data, S, frames, n,
                          block_frames, C.addressof(out), threads)
    assert rc == 0
    return out, n_blocks
