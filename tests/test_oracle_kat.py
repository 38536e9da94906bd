"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
 + the oracle-only hooks for tests that poke reference internals."""
import ctypes as C

import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200.processors import AudioBlock, LoudnessConfig, SpectrogramConfig, SpectrumConfig
from tests import kat

KATS = [getattr(kat, n) for n in sorted(dir(kat)) if n.startswith("kat_")]


@pytest.mark.parametrize("fn", KATS, ids=[f.__name__ for f in KATS])
def test_kat(oracle, fn):
    fn(oracle)


# ---- tests that inspect reference internals: oracle hooks only -------------------------------
def test_power_conversion_preserves_deep_levels(oracle):
    """util/audio/level.rs:45-48"""
    assert abs(oracle.api.power_to_db(1.0e-21, -300.0) + 210.0) < 1e-4
    assert oracle.api.power_to_db(0.0, -140.0) == -140.0
    assert abs(oracle.api.db_to_power(-20.0) - 0.01) < 1e-8


def test_sanitize_sample_rate(oracle):
    """util/audio/rate.rs:6-13"""
    f = oracle.api.sanitize_sample_rate
    assert f(float("nan")) == 48000.0 and f(-1.0) == 48000.0 and f(0.5) == 1.0 and f(1e9) == 768000.0 and f(44100.0) == 44100.0


def test_classic_retention_budget_uses_packed_column_width(oracle):
    """spectrogram/processor.rs:773-792"""
    bins = 16384 * 32 // 2 + 1
    packed = (bins + 1) // 2 * 4
    assert oracle.api.history_columns(0, bins, 8192) == 128 * 1024 * 1024 // packed
    assert oracle.api.history_columns(1, 2049, 0) == 1
    assert oracle.api.history_columns(1, 2049, 100000) == 8192


def test_fft_rebuild_keeps_newest_pending_audio(oracle):
    """spectrogram/processor.rs:794-805 (push via process_block with a window too long to fire)."""
    p = oracle.Spectrogram(SpectrogramConfig(fft_size=256, hop_size=16, history_length=4, use_reassignment=False))
    samples = np.arange(200, dtype=np.float32)
    assert p.process_block(AudioBlock(samples, 1, 48000.0)) is None
    c = p.config()
    c.fft_size = 16
    p.update_config(c)
    buf = np.zeros(512, np.float32)
    n = oracle.api.spectrogram_pending(p._h, buf.ctypes.data_as(C.POINTER(C.c_float)), 512)
    assert n == 32 and np.array_equal(buf[:32], samples[168:])


def test_configured_sources_are_projected_before_fft(oracle):
    """spectrum/processor.rs:480-493"""
    p = oracle.Spectrum(SpectrumConfig(fft_size=8, source=capi.CHANNEL_LEFT, secondary_source=capi.CHANNEL_SIDE))
    p.process_block(AudioBlock(np.array([1.0, 0.0, 0.0, 1.0], np.float32), 2, 48000.0))
    buf = np.zeros(8, np.float32)
    bp = buf.ctypes.data_as(C.POINTER(C.c_float))
    assert oracle.api.spectrum_pending(p._h, 0, bp, 8) == 2 and list(buf[:2]) == [1.0, 0.0]
    assert oracle.api.spectrum_pending(p._h, 1, bp, 8) == 2 and list(buf[:2]) == [0.5, -0.5]


def test_floor_change_reseeds_state_buffers_without_clearing_pending_audio(oracle):
    """spectrum/processor.rs:459-478"""
    p = oracle.Spectrum(SpectrumConfig())
    p.prepare()
    p.process_block(AudioBlock(np.array([0.25, -0.25], np.float32), 1, 48000.0))
    c = p.config()
    c.floor_db = -96.0
    p.update_config(c)
    buf = np.zeros(4, np.float32)
    assert oracle.api.spectrum_pending(p._h, 0, buf.ctypes.data_as(C.POINTER(C.c_float)), 4) == 2
    snap = capi.SpectrumSnapshot()
    oracle.api.spectrum_peek(p._h, C.byref(snap))
    s = p._snapshot(snap)
    assert all(np.all(v == np.float32(-96.0)) for v in s.traces[0])


def _update_outputs(oracle, mode, param, weighting, floor, dt, smoothed, scratch):
    w = np.array(weighting, np.float32)
    sm = np.array(smoothed, np.float32)
    sc = np.array(scratch, np.float32)
    wo = np.zeros_like(w)
    ro = np.zeros_like(w)
    fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
    oracle.api.spectrum_update_outputs(mode, param, 0.0, fp(w), w.size, floor, dt, fp(sm), fp(sc), fp(wo), fp(ro))
    return sm, wo, ro


def test_averaged_power_is_zeroed_below_the_visible_floor(oracle):
    """spectrum/processor.rs:613-627"""
    sm, _, _ = _update_outputs(oracle, capi.AVG_EXPONENTIAL, 0.95, [0.0], -100.0, 1.0,
                               [oracle.api.db_to_power(-101.0)], [0.0])
    assert sm[0] == 0.0


def test_smoothing_retains_power_visible_after_weighting(oracle):
    """spectrum/processor.rs:629-651"""
    for mode, param in [(capi.AVG_EXPONENTIAL, 0.95), (capi.AVG_PEAK_HOLD, 12.0)]:
        _, wo, ro = _update_outputs(oracle, mode, param, [1.2], -100.0, 1.0, [0.0], [oracle.api.db_to_power(-100.5)])
        assert ro[0] == -100.0
        assert -99.4 < wo[0] < -99.2, wo[0]


def _means(oracle, caps, values, leading=0):
    caps_a = (C.c_size_t * len(caps))(*caps)
    v = np.array(values, np.float64)
    out = np.zeros(len(caps), np.float64)
    oracle.api.windowed_means(caps_a, len(caps), leading, v.ctypes.data_as(C.POINTER(C.c_double)), v.size,
                              out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def test_rolling_mean_square_tracks_average(oracle):
    """loudness/processor.rs:323-336"""
    eps = np.finfo(np.float64).eps
    assert abs(_means(oracle, [4, 2, 1, 4], [1.0, 9.0])[0] - 5.0) < eps
    m = _means(oracle, [4, 2, 1, 4], [1.0, 9.0, 16.0, 25.0, 36.0])
    assert abs(m[0] - 21.5) < eps and abs(m[1] - 30.5) < eps and abs(m[2] - 36.0) < eps


def test_running_means_sanitize_non_finite_values(oracle):
    """dsp.rs:626-635"""
    for v in (np.nan, np.inf, -np.inf):
        assert _means(oracle, [1], [v])[0] == 0.0
    assert _means(oracle, [1], [np.nan, np.inf, 1.0])[0] == 1.0


def test_running_means_preserve_small_values_after_a_large_value_expires(oracle):
    """dsp.rs:637-656"""
    assert _means(oracle, [4], [1.0, 1.0e100, 1.0, -1.0e100])[0] == 0.5
    assert _means(oracle, [2], [2.0 ** 53, 1.0, 1.0])[0] == 1.0
    assert _means(oracle, [2], [1.0e100, 2.0, 1.0e-100, 1.0e-100])[0] == 1.0e-100


def test_leading_silence_matches_eager_channel_state(oracle):
    """loudness/processor.rs:400-417"""
    x = kat.sine_wave(1000.0, 48000.0, int(np.float32(48000.0) * np.float32(0.1)), 0.5)
    samples = np.concatenate([np.zeros(48001 * 2, np.float32), np.repeat(x, 2)])
    lazy = oracle.Loudness(LoudnessConfig())
    eager = oracle.Loudness(LoudnessConfig())
    oracle.api.loudness_force_eager(eager._h, 2, 48000.0)
    assert lazy.process_block(AudioBlock(samples, 2, 48000.0)) == eager.process_block(AudioBlock(samples, 2, 48000.0))


def test_oracle_fft_matches_numpy(oracle):
    """The stand-in for rustfft: unnormalised forward/inverse, checked against numpy.fft in f64."""
    rng = np.random.default_rng(0)
    for n in (2, 8, 64, 1024, 8192, 12):
        z = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        buf = z.view(np.float32).copy()
        oracle.api.fft(buf.ctypes.data_as(C.POINTER(C.c_float)), n, 0)
        ref = np.fft.fft(z.astype(np.complex128))
        assert np.max(np.abs(buf.view(np.complex64) - ref)) < 3e-6 * np.sqrt(n) * np.max(np.abs(ref)) / np.sqrt(n) + 1e-4
        oracle.api.fft(buf.ctypes.data_as(C.POINTER(C.c_float)), n, 1)
        assert np.max(np.abs(buf.view(np.complex64) / n - z)) < 1e-5
