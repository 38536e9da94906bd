"""Two interchangeable backends for the shared known-answer tests: the CPU oracle
and the CUDA product, both driven through the same ctypes surface."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

import numpy as np

from openmeters_b200 import _capi as capi
from openmeters_b200 import processors as P


def _utils(api):
    def window(kind, n):
        out = np.zeros(n, np.float32)
        api.window_coefficients(kind, n, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def bin_norm(window_arr, fft_size):
        w = np.ascontiguousarray(window_arr, np.float32)
        out = np.zeros(fft_size // 2 + 1, np.float32)
        api.fft_bin_normalization(w.ctypes.data_as(C.POINTER(C.c_float)), w.size, fft_size,
                                  out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def reassignment_windows(window_arr):
        w = np.ascontiguousarray(window_arr, np.float32)
        d = np.zeros_like(w)
        t = np.zeros_like(w)
        api.reassignment_windows(w.ctypes.data_as(C.POINTER(C.c_float)), w.size,
                                 d.ctypes.data_as(C.POINTER(C.c_float)), t.ctypes.data_as(C.POINTER(C.c_float)))
        return d, t

    def power_scale(window_arr, fft_size):
        w = np.ascontiguousarray(window_arr, np.float32)
        return float(api.reassigned_power_scale(w.ctypes.data_as(C.POINTER(C.c_float)), w.size, fft_size))

    def k_weighting(fs):
        b = (C.c_double * 5)()
        a = (C.c_double * 5)()
        api.k_weighting_coefficients(fs, b, a)
        return np.array(b[:]), np.array(a[:])

    def true_peak_fir(factor):
        n = 36 if factor == 4 else 24
        out = np.zeros(n, np.float32)
        api.true_peak_fir(factor, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out.reshape(12, 3) if factor == 4 else out

    def fallback_positions(ch):
        out = (C.c_uint8 * 8)()
        api.fallback_positions(ch, out)
        return tuple(out[:])

    def stereo_matrix(ch, positions):
        out = np.zeros((8, 2), np.float32)
        api.stereo_matrix(ch, capi.positions_array(positions), out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def downmix(samples, channels, positions, channel):
        s = np.ascontiguousarray(samples, np.float32).reshape(-1)
        frames = s.size // channels
        out = np.zeros(frames, np.float32)
        rc = api.downmix_project(s.ctypes.data_as(C.POINTER(C.c_float)), frames, channels,
                                 capi.positions_array(positions), channel, out.ctypes.data_as(C.POINTER(C.c_float)))
        assert rc == 0, rc
        return out

    return SimpleNamespace(window=window, bin_norm=bin_norm, reassignment_windows=reassignment_windows,
                           power_scale=power_scale, k_weighting=k_weighting, true_peak_fir=true_peak_fir,
                           fallback_positions=fallback_positions, stereo_matrix=stereo_matrix, downmix=downmix,
                           pack_classic_db=lambda db: int(api.pack_classic_db(float(db))),
                           a_weight=lambda f: float(api.a_weight(float(f))))


def _backend(api, name):
    return SimpleNamespace(
        name=name, api=api, u=_utils(api),
        Spectrogram=lambda cfg=None: P.SpectrogramProcessor(cfg, api=api),
        Spectrum=lambda cfg=None: P.SpectrumProcessor(cfg, api=api),
        Loudness=lambda cfg=None: P.LoudnessProcessor(cfg, api=api),
    )


def oracle_backend():
    from oracle import oracle_py

    return _backend(oracle_py.api(), "oracle")


def product_backend():
    from openmeters_b200 import _lib

    return _backend(_lib.api(), "product")


def emu_backend():
    """The product's kernel sources executed thread-for-thread by the CPU emulator (tests/emu).
    Development aid: validates kernel logic without a GPU. Not a product path."""
    import ctypes

    from tests.emu import build_emu

    lib = ctypes.CDLL(build_emu.build(), mode=ctypes.RTLD_LOCAL)
    return _backend(capi.bind(lib, "omb_"), "emu")
