"""Batched (offline) plans over the C ABI: the entry points the streaming processors wrap.

`*_host` methods take/return numpy arrays (H2D/D2H inside the call); `*_device` methods take raw
device pointers (e.g. ``tensor.data_ptr()``) and a CUDA stream handle and leave results in HBM.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from .processors import LoudnessConfig, OmbError, SpectrogramConfig, SpectrumConfig, _check, _default_api


class StftPlan:
    def __init__(self, config: SpectrogramConfig, kernel: int = capi.KERNEL_AUTO, api=None):
        self._api = api or _default_api()
        self.config = config
        self._c = config.to_c()
        self._h = C.c_void_p()
        _check(self._api, self._api.stft_plan_create(C.byref(self._c), kernel, C.byref(self._h)), "stft_plan_create")
        self.bins = int(self._api.stft_plan_bins(self._h))
        self.kernel_generation = int(self._api.stft_plan_is_fast(self._h))  # 0 generic / shared-memory tier, 1..8 specialised (stft.h: fast_kind)
        self.is_fast = self.kernel_generation > 0
        self.power_scale = float(self._api.stft_plan_power_scale(self._h))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.stft_plan_destroy(h)

    def frames_per_lane(self, samples: int) -> int:
        return int(self._api.stft_frames_per_lane(C.byref(self._c), samples))

    def execute_host(self, lanes: np.ndarray, point_stride: int | None = None):
        """lanes (L, S) float32 -> reassigned: (points (L,F,stride,3) f32, counts (L,F) u32); classic: codes (L,F,bins) u16."""
        lanes = np.ascontiguousarray(lanes, np.float32)
        L, S = lanes.shape
        F = self.frames_per_lane(S)
        if self.config.use_reassignment:
            stride = point_stride or self.bins
            pts = np.zeros((L, F, stride, 3), np.float32)
            cnt = np.zeros((L, F), np.uint32)
            _check(self._api, self._api.stft_execute_host(self._h, lanes.ctypes.data, L, S, S, pts.ctypes.data, stride,
                                                          cnt.ctypes.data, None), "stft_execute_host")
            return pts, cnt
        codes = np.zeros((L, F, self.bins), np.uint16)
        _check(self._api, self._api.stft_execute_host(self._h, lanes.ctypes.data, L, S, S, None, 0, None, codes.ctypes.data),
               "stft_execute_host")
        return codes

    def render_host(self, lanes: np.ndarray, view, want_counts: bool = True):
        """STFT -> splat accumulate -> resolve on the device (omb_stft_render_host): lanes (L, S) float32 and a
        `splat.SplatParams` view -> dB images (L, H, W) float32 (and counts (L, F)); the points never leave the GPU."""
        lanes = np.ascontiguousarray(lanes, np.float32)
        L, S = lanes.shape
        F = self.frames_per_lane(S)
        c = view.to_c()
        c.ring_capacity = max(F, 1)
        w, h = C.c_uint32(), C.c_uint32()
        self._api.splat_image_size(C.byref(c), C.byref(w), C.byref(h))
        db = np.full((L, h.value, w.value), -np.inf, np.float32)
        cnt = np.zeros((L, F), np.uint32)
        _check(self._api, self._api.stft_render_host(self._h, lanes.ctypes.data, L, S, S, C.byref(c), db.ctypes.data,
                                                     cnt.ctypes.data if want_counts else None), "stft_render_host")
        return (db, cnt) if want_counts else db

    def execute_device(self, lanes_ptr: int, n_lanes: int, samples: int, lane_stride: int, points_ptr: int = 0,
                       point_stride: int = 0, counts_ptr: int = 0, classic_ptr: int = 0, stream: int = 0) -> None:
        _check(self._api, self._api.stft_execute_device(self._h, lanes_ptr, n_lanes, samples, lane_stride, points_ptr or None,
                                                        point_stride, counts_ptr or None, classic_ptr or None, stream or None),
               "stft_execute_device")


class SpectrumPlan:
    def __init__(self, config: SpectrumConfig, api=None):
        self._api = api or _default_api()
        self.config = config
        self._c = config.to_c()
        self._h = C.c_void_p()
        _check(self._api, self._api.spectrum_plan_create(C.byref(self._c), C.byref(self._h)), "spectrum_plan_create")
        self.bins = config.fft_size // 2 + 1

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.spectrum_plan_destroy(h)

    def hops_per_lane(self, samples: int) -> int:
        return int(self._api.spectrum_hops_per_lane(C.byref(self._c), samples))

    def execute_host(self, lanes: np.ndarray, want_peak: bool = True):
        lanes = np.ascontiguousarray(lanes, np.float32)
        L, S = lanes.shape
        H = self.hops_per_lane(S)
        w = np.zeros((L, H, self.bins), np.float32)
        r = np.zeros((L, H, self.bins), np.float32)
        pk = np.full((L, H), -1, np.int32)
        _check(self._api, self._api.spectrum_execute_host(self._h, lanes.ctypes.data, L, S, S, w.ctypes.data, r.ctypes.data,
                                                          pk.ctypes.data if want_peak else None), "spectrum_execute_host")
        return w, r, pk

    def set_peak_spec(self, trace: int = 0, min_hz: float = 20.0, max_hz: float = 0.0) -> None:
        """Peak label spec (spectrum/state.rs:106-107,134-136): trace 0 = A-weighted / 1 = raw, candidate range in Hz
        (max_hz <= 0: up to the last bin)."""
        spec = capi.SpectrumPeakSpec(trace, min_hz, max_hz)
        _check(self._api, self._api.spectrum_plan_set_peak_spec(self._h, C.byref(spec)), "spectrum_plan_set_peak_spec")

    def peak_spec(self):
        spec = capi.SpectrumPeakSpec()
        _check(self._api, self._api.spectrum_plan_get_peak_spec(self._h, C.byref(spec)), "spectrum_plan_get_peak_spec")
        return int(spec.trace), float(spec.min_hz), float(spec.max_hz)

    def execute_host_peaks(self, lanes: np.ndarray):
        """execute_host + interpolated_peak (state.rs:327-356) of every hop -> (weighted, raw, peak_bin, freq_hz, level_db)."""
        lanes = np.ascontiguousarray(lanes, np.float32)
        L, S = lanes.shape
        H = self.hops_per_lane(S)
        w = np.zeros((L, H, self.bins), np.float32)
        r = np.zeros((L, H, self.bins), np.float32)
        pk = np.full((L, H), -1, np.int32)
        f = np.full((L, H), np.nan, np.float32)
        m = np.full((L, H), np.nan, np.float32)
        _check(self._api, self._api.spectrum_execute_host_peaks(self._h, lanes.ctypes.data, L, S, S, w.ctypes.data, r.ctypes.data,
                                                                pk.ctypes.data, f.ctypes.data, m.ctypes.data),
               "spectrum_execute_host_peaks")
        return w, r, pk, f, m

    def interpolate_peaks_device(self, db_ptr: int, peak_ptr: int, rows: int, freq_ptr: int, level_ptr: int, stream: int = 0) -> None:
        _check(self._api, self._api.spectrum_interpolate_peaks_device(self._h, db_ptr, peak_ptr, rows, freq_ptr, level_ptr,
                                                                      stream or None), "spectrum_interpolate_peaks_device")

    def execute_device(self, lanes_ptr: int, n_lanes: int, samples: int, lane_stride: int, weighted_ptr: int, raw_ptr: int,
                       peak_ptr: int = 0, stream: int = 0) -> None:
        _check(self._api, self._api.spectrum_execute_device(self._h, lanes_ptr, n_lanes, samples, lane_stride, weighted_ptr,
                                                            raw_ptr, peak_ptr or None, stream or None), "spectrum_execute_device")


class LoudnessPlan:
    def __init__(self, config: LoudnessConfig, channels: int, positions=None, api=None):
        self._api = api or _default_api()
        self.config = config
        self.channels = channels
        self._c = capi.LoudnessConfig(config.sample_rate, config.floor_db)
        self._h = C.c_void_p()
        _check(self._api, self._api.loudness_plan_create(C.byref(self._c), channels, capi.positions_array(positions),
                                                         C.byref(self._h)), "loudness_plan_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.loudness_plan_destroy(h)

    def execute_host(self, streams: np.ndarray, block_frames: int):
        """streams (n_streams, frames*channels) interleaved f32 -> ctypes array of LoudnessSnapshot [n_streams*n_blocks]."""
        streams = np.ascontiguousarray(streams, np.float32)
        S, n = streams.shape
        frames = n // self.channels
        n_blocks = (frames + block_frames - 1) // block_frames
        out = (capi.LoudnessSnapshot * (S * n_blocks))()
        _check(self._api, self._api.loudness_execute_host(self._h, streams.ctypes.data, S, frames, n, block_frames,
                                                          C.addressof(out)), "loudness_execute_host")
        return out, n_blocks

    def execute_device(self, in_ptr: int, n_streams: int, frames: int, stream_stride: int, block_frames: int, out_ptr: int,
                       stream: int = 0) -> None:
        _check(self._api, self._api.loudness_execute_device(self._h, in_ptr, n_streams, frames, stream_stride, block_frames,
                                                            out_ptr, stream or None), "loudness_execute_device")


def snapshots_to_arrays(snaps, n: int):
    """ctypes LoudnessSnapshot array -> dict of numpy arrays (for comparisons)."""
    st = np.array([snaps[i].short_term_loudness for i in range(n)], np.float32)
    mo = np.array([snaps[i].momentary_loudness for i in range(n)], np.float32)
    fast = np.array([snaps[i].rms_fast_db[:] for i in range(n)], np.float32)
    slow = np.array([snaps[i].rms_slow_db[:] for i in range(n)], np.float32)
    tp = np.array([snaps[i].true_peak_db[:] for i in range(n)], np.float32)
    return dict(short_term=st, momentary=mo, rms_fast=fast, rms_slow=slow, true_peak=tp)
