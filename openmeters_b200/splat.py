"""Host-side mirror of the spectrogram view's accumulation + resolve passes over the C ABI (row f2 of SURVEY.md §8):
reassigned points in, power / dB images out (spectrogram/render.rs:104-165, render/shaders/spectrogram.wgsl)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Tuple

import numpy as np

from . import _capi as capi
from .processors import _check, _default_api


@dataclass
class SplatParams:
    """The subset of render.rs `Uniforms` the two passes read (same meaning as omb_splat_params)."""

    freq_scale: int = capi.FREQ_LOG
    freq_min: float = 1.0
    freq_max: float = 24000.0
    uv_y_range: Tuple[float, float] = (0.0, 1.0)
    ext_w: float = 1024.0
    ext_h: float = 512.0
    scale_factor: float = 1.0
    tilt_db: float = 0.0
    ring_capacity: int = 1024
    newest_col: int = 0
    col_count: int = 0
    reassigned_power_scale: float = 1.0

    def to_c(self) -> capi.SplatParams:
        c = capi.SplatParams()
        c.freq_scale = self.freq_scale
        c.freq_min, c.freq_max = self.freq_min, self.freq_max
        c.uv_y_range[0], c.uv_y_range[1] = self.uv_y_range
        c.ext_w, c.ext_h = self.ext_w, self.ext_h
        c.scale_factor = self.scale_factor
        c.tilt_db = self.tilt_db
        c.ring_capacity = self.ring_capacity
        c.newest_col = self.newest_col
        c.col_count = self.col_count
        c.reassigned_power_scale = self.reassigned_power_scale
        return c


def display_axis(sample_rate: float) -> Tuple[float, float]:
    """spectrogram/state.rs:49-52."""
    nyq = max(sample_rate / 2.0, 1.0)
    return min(1.0, nyq * 0.5), nyq


def image_size(params: SplatParams, api=None) -> Tuple[int, int]:
    api = api or _default_api()
    w, h = C.c_uint32(), C.c_uint32()
    c = params.to_c()
    api.splat_image_size(C.byref(c), C.byref(w), C.byref(h))
    return w.value, h.value


def render_host(rings: np.ndarray, counts: np.ndarray, params: SplatParams, api=None):
    """rings: (n_rings, ring_capacity, stride, 3) float32; counts: (n_rings, ring_capacity) uint32.
    Returns (accum, db), each (n_rings, H, W) float32."""
    api = api or _default_api()
    rings = np.ascontiguousarray(rings, np.float32)
    counts = np.ascontiguousarray(counts, np.uint32)
    n_rings, hl, stride, three = rings.shape
    assert three == 3 and hl == params.ring_capacity and counts.shape == (n_rings, hl)
    w, h = image_size(params, api)
    accum = np.empty((n_rings, h, w), np.float32)
    db = np.empty((n_rings, h, w), np.float32)
    c = params.to_c()
    _check(api, api.splat_render_host(rings.ctypes.data, stride, counts.ctypes.data, n_rings, C.byref(c),
                                      accum.ctypes.data, db.ctypes.data), "splat_render_host")
    return accum, db
