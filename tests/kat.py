"""Known-answer tests restated from the reference's in-file `#[cfg(test)]` modules.

Each function takes a backend (tests/backends.py) so the same KAT pins the CPU
oracle (CPU suite) and the CUDA product (`-m gpu` suite).  The docstring of each
names the reference test it restates (paths relative to /root/reference/).
"""
from __future__ import annotations

import numpy as np

from openmeters_b200 import _capi as capi
from openmeters_b200.processors import (AudioBlock, LoudnessConfig, SpectrogramConfig, SpectrumConfig)

TAU32 = np.float32(6.28318530717958647692)
DB_FLOOR = -140.0
ANALYSIS_FLOOR_POWER = 1e-14


def sine_wave(freq, sample_rate, count, amplitude=1.0):
    """src/util/audio.rs:28-33 — all f32."""
    i = np.arange(count, dtype=np.float32)
    ph = (TAU32 * np.float32(freq) * i) / np.float32(sample_rate)
    return (np.sin(ph, dtype=np.float32) * np.float32(amplitude)).astype(np.float32)


def _cfg(fft_size, hop_size, use_reassignment, **kw):
    """spectrogram/processor.rs:626-635 helper `cfg`."""
    base = dict(fft_size=fft_size, hop_size=hop_size, history_length=4, use_reassignment=use_reassignment,
                zero_padding_factor=1)
    base.update(kw)
    return SpectrogramConfig(**base)


def _process_samples(b, cfg, samples):
    p = b.Spectrogram(cfg)
    up = p.process_block(AudioBlock(samples, 1, cfg.sample_rate))
    assert up is not None, "expected snapshot"
    return up


def _peak_point(points):
    pts = points[points[:, 2] > ANALYSIS_FLOOR_POWER]
    assert len(pts), "expected non-sentinel point"
    return pts[np.argmax(pts[:, 2])]


# ------------------------------------------------------------------ util/audio
def kat_fft_windows_are_periodic(b):
    """util/audio/window.rs:115-122"""
    hann = b.u.window(capi.WINDOW_HANN, 8)
    assert hann[0] == 0.0
    assert abs(hann[4] - 1.0) < 1e-6
    assert abs(hann[7] - 0.1464465) < 1e-6


def kat_window_families(b):
    """window.rs:20-43 — closed forms in f64 for all five kinds (periodic)."""
    n = 64
    ph = 2 * np.pi * np.arange(n) / n
    exp = {
        capi.WINDOW_RECTANGULAR: np.ones(n),
        capi.WINDOW_HANN: 0.5 - 0.5 * np.cos(ph),
        capi.WINDOW_HAMMING: 25 / 46 - 21 / 46 * np.cos(ph),
        capi.WINDOW_BLACKMAN: 0.42 - 0.5 * np.cos(ph) + 0.08 * np.cos(2 * ph),
        capi.WINDOW_BLACKMAN_HARRIS: 0.35875 - 0.48829 * np.cos(ph) + 0.14128 * np.cos(2 * ph) - 0.01168 * np.cos(3 * ph),
    }
    for kind, e in exp.items():
        np.testing.assert_allclose(b.u.window(kind, n), e, atol=2e-6)
    assert list(b.u.window(capi.WINDOW_HANN, 1)) == [1.0]


def kat_bin_normalization(b):
    """window.rs:90-109"""
    w = b.u.window(capi.WINDOW_HANN, 16)
    norm = b.u.bin_norm(w, 16)
    inv = 1.0 / float(np.sum(w.astype(np.float64)))
    assert norm.shape == (9,)
    np.testing.assert_allclose(norm[1:-1], 4 * inv * inv, rtol=1e-6)
    np.testing.assert_allclose(norm[[0, -1]], inv * inv, rtol=1e-6)
    # zero-sum window falls back to 1/fft_size
    z = b.u.bin_norm(np.zeros(8, np.float32), 8)
    np.testing.assert_allclose(z[1], 4.0 / 64.0, rtol=1e-6)


# ------------------------------------------------------------------ spectrogram
def kat_classic_db_packing_rounds_to_nearest_code(b):
    """spectrogram/processor.rs:663-668"""
    step = np.float32(156.0) / np.float32(65535.0)
    lo = np.float32(-144.0)
    assert b.u.pack_classic_db(lo + step * np.float32(1234.49)) == 1234
    assert b.u.pack_classic_db(lo + step * np.float32(1234.50)) == 1235
    assert b.u.pack_classic_db(-1000.0) == 0 and b.u.pack_classic_db(1000.0) == 65535


def kat_invalid_config_values_are_normalized(b):
    """spectrogram/processor.rs:670-684"""
    p = b.Spectrogram(SpectrogramConfig(sample_rate=float("nan"), fft_size=0, hop_size=0, zero_padding_factor=0))
    c = p.config()
    assert c.sample_rate == 48000.0 and c.fft_size == 2048 and c.hop_size == 64 and c.zero_padding_factor == 1


def kat_switching_analysis_modes_rebuilds_the_active_buffers(b):
    """spectrogram/processor.rs:686-707"""
    p = b.Spectrogram(_cfg(64, 16, True))
    c = p.config()
    c.use_reassignment = False
    p.update_config(c)
    classic = p.process_block(AudioBlock(np.full(64, 0.25, np.float32), 1, c.sample_rate))
    assert classic is not None and classic.kind == capi.COLUMN_CLASSIC
    c.use_reassignment = True
    p.update_config(c)
    p.reset_audio()
    re = p.process_block(AudioBlock(np.full(128, 0.25, np.float32), 1, c.sample_rate))
    assert re is not None and re.kind == capi.COLUMN_REASSIGNED


def kat_detects_sine_frequency_peak(b):
    """spectrogram/processor.rs:709-724"""
    cfg = _cfg(1024, 512, False, history_length=8, window=capi.WINDOW_HANN)
    freq = np.float32(200.0) * np.float32(cfg.sample_rate) / np.float32(cfg.fft_size)
    up = _process_samples(b, cfg, sine_wave(freq, cfg.sample_rate, 2048))
    mags = up.new_columns[-1]
    assert mags.shape == (cfg.fft_size // 2 + 1,)
    # Rust max_by_key returns the LAST maximum
    idx = len(mags) - 1 - int(np.argmax(mags[::-1]))
    assert idx == 200
    assert mags[idx] >= b.u.pack_classic_db(-0.01)


def kat_retained_history_matches_full_suffix(b):
    """spectrogram/processor.rs:726-743"""
    full_cfg = _cfg(64, 16, False, history_length=32)
    capped_cfg = _cfg(64, 16, False, history_length=3)
    i = np.arange(192)
    samples = np.sin(((i * i + 3 * i).astype(np.float32) * np.float32(0.017)), dtype=np.float32)
    full = _process_samples(b, full_cfg, samples)
    capped = _process_samples(b, capped_cfg, samples)
    expected = full.new_columns[len(full.new_columns) - len(capped.new_columns):]
    assert len(capped.new_columns) == 3
    assert not np.array_equal(full.new_columns[0], expected[0])
    for e, a in zip(expected, capped.new_columns):
        assert np.array_equal(e, a)


def kat_hops_larger_than_the_window_are_block_partition_independent(b):
    """spectrogram/processor.rs:745-771"""
    cfg = SpectrogramConfig(sample_rate=32.0, fft_size=8, hop_size=16, window=capi.WINDOW_RECTANGULAR,
                            history_length=32, use_reassignment=False)
    samples = np.sin(np.arange(29, dtype=np.float32) * np.float32(0.73), dtype=np.float32)
    whole = _process_samples(b, cfg, samples).new_columns
    p = b.Spectrogram(cfg)
    parts = []
    for s in range(0, 29, 8):
        up = p.process_block(AudioBlock(samples[s:s + 8], 1, 32.0))
        if up is not None:
            parts.extend(up.new_columns)
    assert len(whole) == len(parts) and len(whole) == 2
    for e, a in zip(whole, parts):
        assert np.array_equal(e, a)


def kat_silent_input_advances_transparent_columns(b):
    """spectrogram/processor.rs:807-825"""
    samples = np.zeros(192, np.float32)
    floor = b.u.pack_classic_db(DB_FLOOR)
    classic = _process_samples(b, _cfg(64, 16, False), samples)
    assert len(classic.new_columns) == 4
    assert all(np.all(c == floor) for c in classic.new_columns)
    re = _process_samples(b, _cfg(64, 16, True), samples)
    assert len(re.new_columns) == 4
    assert all(len(c) == 0 for c in re.new_columns)


def kat_reassignment_places_peak_frequency_time_and_power(b):
    """spectrogram/processor.rs:827-860"""
    cfg = _cfg(2048, 512, True, zero_padding_factor=4)
    latency = (4096 - cfg.fft_size) // 2
    expected_time = -latency / cfg.hop_size
    for bin_ in [3.4, 10.25, 50.25, 200.75, 800.4]:
        freq = np.float32(bin_) * np.float32(cfg.sample_rate) / np.float32(cfg.fft_size)
        up = _process_samples(b, cfg, sine_wave(freq, cfg.sample_rate, 4096))
        pts = up.new_columns[-1]
        peak = _peak_point(pts)
        assert abs(peak[1] - freq) < 2.0, (peak, freq)
        assert abs(peak[0] - expected_time) < 0.05, (peak, expected_time)
        power = float(np.sum(pts[:, 2], dtype=np.float32)) * up.reassigned_power_scale
        assert abs(power - 1.0) < 0.01, power
        assert len(pts) < up.fft_size // 2 + 1
        assert up.fft_size == 8192


def kat_reassignment_resolves_a_low_fractional_fft_bin(b):
    """spectrogram/processor.rs:862-874"""
    cfg = _cfg(2048, 512, True, zero_padding_factor=4)
    freq = np.float32(1.37) * np.float32(cfg.sample_rate) / np.float32(cfg.fft_size)
    up = _process_samples(b, cfg, sine_wave(freq, cfg.sample_rate, 4096))
    peak = _peak_point(up.new_columns[-1])
    assert freq < cfg.sample_rate / cfg.fft_size * 2.0
    assert abs(peak[1] - freq) < 2.0


def kat_reassignment_removes_constant_dc(b):
    """spectrogram/processor.rs:876-888"""
    up = _process_samples(b, _cfg(64, 16, True), np.full(128, 0.25, np.float32))
    assert len(up.new_columns) >= 1
    for col in up.new_columns:
        assert len(col) == 0


def kat_reassignment_localizes_a_centered_impulse_in_time(b):
    """spectrogram/processor.rs:890-908"""
    cfg = _cfg(256, 32, True)
    read_len = 512
    center_offset = (read_len - cfg.fft_size) // 2
    position = cfg.fft_size // 2
    samples = np.zeros(read_len, np.float32)
    samples[center_offset + position] = 1.0
    up = _process_samples(b, cfg, samples)
    pts = up.new_columns[-1]
    expected = (position - (cfg.fft_size - 1) * 0.5 - center_offset) / cfg.hop_size
    assert len(pts) > 0
    assert np.all(np.abs(pts[:, 0] - expected) < 1e-4), np.max(np.abs(pts[:, 0] - expected))


def kat_history_length_zero_keeps_one_column(b):
    """spectrogram/processor.rs:153-158,301-304: default history_length=0 => 1 column per call."""
    cfg = _cfg(64, 16, False, history_length=0)
    up = _process_samples(b, cfg, sine_wave(1000.0, 48000.0, 192))
    assert len(up.new_columns) == 1
    ref = _process_samples(b, _cfg(64, 16, False, history_length=32), sine_wave(1000.0, 48000.0, 192))
    assert np.array_equal(up.new_columns[0], ref.new_columns[-1])


def kat_update_reports_reset_once(b):
    """spectrogram/processor.rs:506-515 — `reset` is taken on the first update only."""
    p = b.Spectrogram(_cfg(64, 16, False))
    x = sine_wave(2000.0, 48000.0, 64)
    assert p.process_block(AudioBlock(x, 1, 48000.0)).reset is True
    assert p.process_block(AudioBlock(x, 1, 48000.0)).reset is False
    p.reset_audio()
    assert p.process_block(AudioBlock(x, 1, 48000.0)).reset is True
    # an empty block yields None (processor.rs:491)
    assert p.process_block(AudioBlock(np.zeros(0, np.float32), 1, 48000.0)) is None


def kat_hop_change_mid_stream_uses_the_live_hop(b):
    """spectrogram/processor.rs:283,446-450,518-543 — update_config with ONLY the hop changed does not rebuild the transforms;
    process_ready_windows and reassigned_points read self.config.hop_size live, so the very next column steps by the new hop
    and its time_offset is (T.S*/|S|^2 - center_offset) / NEW hop; `reset` is reported once; pending audio is kept."""
    n = 1024
    cfg = _cfg(n, 256, True, history_length=64, window=capi.WINDOW_HANN)
    freq = np.float32(100.0) * np.float32(cfg.sample_rate) / np.float32(n)
    x = sine_wave(freq, cfg.sample_rate, 2 * n + 8 * 256, 0.5)
    p = b.Spectrogram(cfg)
    first = p.process_block(AudioBlock(x[: 2 * n + 256], 1, cfg.sample_rate))
    assert first is not None and len(first.new_columns) == 2 and first.reset is True and first.hop_size == 256
    assert abs(float(_peak_point(first.new_columns[-1])[0]) + 2.0) < 0.05   # -(2048-1024)/2/256
    c = p.config()
    c.hop_size = 128
    p.update_config(c)
    assert p.config().hop_size == 128 and p.config().fft_size == n
    # pending after two hops of 256: 2n+256-512 = 2n-256 samples; 384 more complete one frame plus one new hop of 128
    second = p.process_block(AudioBlock(x[2 * n + 256: 2 * n + 256 + 384], 1, cfg.sample_rate))
    assert second is not None and second.reset is True and second.hop_size == 128
    assert len(second.new_columns) == 2
    for col in second.new_columns:
        pk = _peak_point(col)
        assert abs(float(pk[0]) + 4.0) < 0.05, pk                              # -(2048-1024)/2/128
        assert abs(float(pk[1]) - float(freq)) < 2.0
    third = p.process_block(AudioBlock(x[2 * n + 256 + 384: 2 * n + 256 + 384 + 128], 1, cfg.sample_rate))
    assert third is not None and third.reset is False and len(third.new_columns) == 1


def kat_window_and_mode_change_mid_stream(b):
    """spectrogram/processor.rs:518-543,229-279 — a window change rebuilds (reset reported, newest pending audio kept:
    fft_rebuild_keeps_newest_pending_audio :794-805), toggling reassignment switches the column kind, and the columns
    after each change equal those of a fresh processor with the new config fed the retained audio."""
    n, hop = 256, 64
    i = np.arange(6 * n)
    x = np.sin(((i * i + 5 * i).astype(np.float32) * np.float32(0.0009)), dtype=np.float32)
    cfg = _cfg(n, hop, False, history_length=64, window=capi.WINDOW_HANN)
    p = b.Spectrogram(cfg)
    a = p.process_block(AudioBlock(x[: 2 * n], 1, cfg.sample_rate))
    assert a is not None and a.kind == capi.COLUMN_CLASSIC and len(a.new_columns) == (2 * n - n) // hop + 1
    consumed = len(a.new_columns) * hop                       # samples drained so far; pending = 2n - consumed
    c = p.config()
    c.window = capi.WINDOW_BLACKMAN_HARRIS
    p.update_config(c)
    u = p.process_block(AudioBlock(x[2 * n: 3 * n], 1, cfg.sample_rate))
    assert u is not None and u.reset is True and u.kind == capi.COLUMN_CLASSIC
    fresh = b.Spectrogram(_cfg(n, hop, False, history_length=64, window=capi.WINDOW_BLACKMAN_HARRIS))
    v = fresh.process_block(AudioBlock(x[consumed: 3 * n], 1, cfg.sample_rate))
    assert len(u.new_columns) == len(v.new_columns) > 0
    for e, g in zip(v.new_columns, u.new_columns):
        assert np.array_equal(e, g)
    consumed += len(u.new_columns) * hop
    c.use_reassignment = True
    p.update_config(c)
    w = p.process_block(AudioBlock(x[3 * n: 5 * n], 1, cfg.sample_rate))
    assert w is not None and w.reset is True and w.kind == capi.COLUMN_REASSIGNED
    fresh = b.Spectrogram(_cfg(n, hop, True, history_length=64, window=capi.WINDOW_BLACKMAN_HARRIS))
    # rebuild_fft drains the pending audio to at most 2 * active_len = 4n samples (:275-277): nothing is dropped here
    z = fresh.process_block(AudioBlock(x[consumed: 5 * n], 1, cfg.sample_rate))
    assert len(w.new_columns) == len(z.new_columns) > 0
    for e, g in zip(z.new_columns, w.new_columns):
        assert np.array_equal(e, g)


# ------------------------------------------------------------------ spectrum
def kat_changing_averaging_mode_clears_stale_state(b):
    """spectrum/processor.rs:566-581 (and update_config :300-322).  The reference pokes `levels[0].smoothed_power`; the
    behavioural statement of the same fact: after PeakHold has accumulated a loud passage, switching to Exponential
    mid-stream must give — from that hop on — exactly what a FRESH Exponential processor gives on the retained audio
    (only the level buffers are reset: pending PCM is kept), and a change of the decay PARAMETER alone (same enum
    discriminant, same floor) must keep the held state."""
    n, hop, sr = 256, 64, 48000.0
    loud = sine_wave(3000.0, sr, 4 * n, 0.9)
    quiet = sine_wave(3000.0, sr, 4 * n, 0.001)
    base = dict(sample_rate=sr, fft_size=n, hop_size=hop, window=capi.WINDOW_HANN, floor_db=-120.0)
    p = b.Spectrum(SpectrumConfig(averaging=capi.AVG_PEAK_HOLD, averaging_param=0.5, **base))
    s0 = p.process_block(AudioBlock(loud, 1, sr))
    assert s0 is not None
    hops_done = (loud.size - n) // hop + 1
    held_peak = float(np.max(s0.traces[0][1]))
    c = p.config()
    c.averaging, c.averaging_param = capi.AVG_EXPONENTIAL, 0.5
    p.update_config(c)
    s1 = p.process_block(AudioBlock(quiet, 1, sr))
    fresh = b.Spectrum(SpectrumConfig(averaging=capi.AVG_EXPONENTIAL, averaging_param=0.5, **base))
    s2 = fresh.process_block(AudioBlock(np.concatenate([loud[hops_done * hop:], quiet]), 1, sr))
    assert s1 is not None and s2 is not None
    for w in range(2):
        assert np.array_equal(s1.traces[0][w], s2.traces[0][w])
    assert float(np.max(s1.traces[0][1])) < held_peak - 30.0      # stale hold (decaying 0.5 dB/s) would still sit at held_peak
    # parameter-only change: state survives
    q = b.Spectrum(SpectrumConfig(averaging=capi.AVG_PEAK_HOLD, averaging_param=0.5, **base))
    q.process_block(AudioBlock(loud, 1, sr))
    c = q.config()
    c.averaging_param = 1.0
    q.update_config(c)
    s3 = q.process_block(AudioBlock(quiet, 1, sr))
    assert float(np.max(s3.traces[0][1])) > held_peak - 1.0


def kat_spectrum_normalization(b):
    """spectrum/processor.rs:432-457"""
    p = b.Spectrum(SpectrumConfig(sample_rate=float("nan"), fft_size=0, hop_size=0, floor_db=float("inf")))
    c = p.config()
    assert c.fft_size == 1 and c.hop_size == 1 and c.floor_db == -100.0 and c.sample_rate == 48000.0
    assert b.Spectrum(SpectrumConfig(floor_db=1.0)).config().floor_db == -100.0
    assert b.Spectrum(SpectrumConfig(floor_db=-280.0)).config().floor_db == -280.0


def kat_secondary_source_can_drive_processing_without_primary(b):
    """spectrum/processor.rs:495-509"""
    p = b.Spectrum(SpectrumConfig(fft_size=8, hop_size=8, source=capi.CHANNEL_NONE, secondary_source=capi.CHANNEL_LEFT))
    assert p.process_block(AudioBlock(np.zeros(8, np.float32), 1, 48000.0)) is not None


def kat_fft_size_update_resizes_before_processing(b):
    """spectrum/processor.rs:511-536"""
    p = b.Spectrum(SpectrumConfig(fft_size=128, hop_size=128))
    p.prepare()
    c = p.config()
    c.fft_size = 256
    c.hop_size = 256
    p.update_config(c)
    bins = 129
    snap = p.process_block(AudioBlock(np.zeros(256, np.float32), 1, c.sample_rate))
    assert snap is not None
    assert (len(snap.traces[0][0]), len(snap.traces[0][1])) == (bins, bins)


def kat_peak_hold_decays_for_each_audio_hop_in_large_batch(b):
    """spectrum/processor.rs:538-563"""
    p = b.Spectrum(SpectrumConfig(sample_rate=8.0, fft_size=8, hop_size=8, window=capi.WINDOW_RECTANGULAR,
                                  averaging=capi.AVG_PEAK_HOLD, averaging_param=24.0, floor_db=-100.0))
    samples = np.concatenate([sine_wave(1.0, 8.0, 8, 1.0), np.zeros(8, np.float32)])
    snap = p.process_block(AudioBlock(samples, 1, 8.0))
    assert snap is not None
    held = snap.traces[0][1][1]
    assert -24.1 < held < -23.9, held


def kat_spectrum_hops_larger_than_fft_partition_independent(b):
    """spectrum/processor.rs:583-611"""
    cfg = SpectrumConfig(sample_rate=32.0, fft_size=8, hop_size=16, window=capi.WINDOW_RECTANGULAR,
                         source=capi.CHANNEL_LEFT)
    samples = np.sin(np.arange(29, dtype=np.float32) * np.float32(0.73), dtype=np.float32)
    whole = b.Spectrum(cfg).process_block(AudioBlock(samples, 1, 32.0))
    p = b.Spectrum(cfg)
    part = None
    for s in range(0, 29, 8):
        r = p.process_block(AudioBlock(samples[s:s + 8], 1, 32.0))
        part = r if r is not None else part
    assert whole is not None and part is not None
    for w in range(2):
        assert np.array_equal(whole.traces[0][w], part.traces[0][w])


def kat_a_weight_matches_iec_reference_points(b):
    """spectrum/processor.rs:653-678"""
    ref = [(1.0, -148.6), (5.0, -93.1), (31.5, -39.4), (63.0, -26.2), (100.0, -19.1), (200.0, -10.9), (500.0, -3.2),
           (1000.0, 0.0), (2000.0, 1.2), (4000.0, 1.0), (8000.0, -1.1), (16000.0, -6.6)]
    for f, e in ref:
        assert abs(b.u.a_weight(f) - e) <= 0.15, (f, b.u.a_weight(f))
    assert b.u.a_weight(0.0) == -np.inf


def kat_spectrum_floor_and_weighting(b):
    """spectrum/processor.rs:613-651 behaviourally: power below the raw floor but lifted above it by
    positive A-weighting stays visible in the weighted trace; raw trace sits on the floor."""
    sr, n = 48000.0, 1024
    # 3 kHz is where A-weighting is ~ +1.2 dB.  Amplitude for ~ -100.5 dB raw level.
    amp = 10 ** (-100.5 / 20)
    k = 64  # 64*48000/1024 = 3000 Hz, bin-centred
    x = sine_wave(k * sr / n, sr, n, amp)
    for mode, param in [(capi.AVG_NONE, 0.0), (capi.AVG_EXPONENTIAL, 0.95), (capi.AVG_PEAK_HOLD, 12.0)]:
        p = b.Spectrum(SpectrumConfig(sample_rate=sr, fft_size=n, hop_size=n, window=capi.WINDOW_RECTANGULAR,
                                      averaging=mode, averaging_param=param, source=capi.CHANNEL_LEFT, floor_db=-100.0))
        snap = p.process_block(AudioBlock(x, 1, sr))
        raw, wt = snap.traces[0][1][k], snap.traces[0][0][k]
        assert raw == -100.0, (mode, raw)
        assert -99.6 < wt < -99.0, (mode, wt)
        # a quiet bin far from the tone is floored in both
        assert snap.traces[0][0][300] == -100.0 and snap.traces[0][1][300] == -100.0
        assert snap.frequency_bins[k] == np.float32(k) * (np.float32(sr) / np.float32(n))


# ------------------------------------------------------------------ dsp.rs
def kat_fallback_positions(b):
    """dsp.rs:36-47 and the layouts asserted at dsp.rs:511-537"""
    FL, FR, FC, LFE, RL, RR, SL, SR_, MONO, UNK = range(10)
    assert b.u.fallback_positions(1)[:1] == (MONO,)
    assert b.u.fallback_positions(2)[:2] == (FL, FR)
    assert b.u.fallback_positions(4)[:4] == (FL, FR, RL, RR)
    assert b.u.fallback_positions(5)[:5] == (FL, FR, FC, RL, RR)
    assert b.u.fallback_positions(6)[:6] == (FL, FR, FC, LFE, RL, RR)
    assert b.u.fallback_positions(8) == (FL, FR, FC, LFE, RL, RR, SL, SR_)
    assert b.u.fallback_positions(3)[3:] == (UNK,) * 5


def kat_stereo_matrix_folds_semantic_channels_and_ignores_lfe(b):
    """dsp.rs:558-589"""
    samples = np.array([1.0, 2.0, 3.0, 100.0, 4.0, 5.0, 6.0, 7.0], np.float32)
    g = np.float32(0.70710678118654752440)
    left = b.u.downmix(samples, 8, capi.SURROUND, capi.CHANNEL_LEFT)
    right = b.u.downmix(samples, 8, capi.SURROUND, capi.CHANNEL_RIGHT)
    # fold order ch0..ch7 from 0.0, f32 (dsp.rs:240-247)
    l = np.float32(0.0)
    r = np.float32(0.0)
    w = [(1, 0), (0, 1), (g, g), (0, 0), (g, 0), (0, g), (g, 0), (0, g)]
    for s, (wl, wr) in zip(samples, w):
        l = np.float32(l + s * np.float32(wl))
        r = np.float32(r + s * np.float32(wr))
    assert left[0] == l and right[0] == r
    assert abs(left[0] - (1.0 + g * 13.0)) < 1e-5 and abs(right[0] - (2.0 + g * 15.0)) < 1e-5
    m = b.u.stereo_matrix(1, [capi.POS_MONO])
    assert tuple(m[0]) == (1.0, 1.0)
    m = b.u.stereo_matrix(8, [capi.POS_LOW_FREQUENCY, capi.POS_AUX0] + [capi.POS_UNKNOWN] * 6)
    assert [tuple(r_) for r_ in m[:2]] == [(1.0, 0.0), (0.0, 1.0)]


def kat_common_stereo_paths_preserve_general_fold_bits(b):
    """dsp.rs:591-624 — NaN payloads, -0.0 and inf must fold bit-exactly."""
    nan = np.array([0x7FC01234], np.uint32).view(np.float32)[0]
    for samples, ch in [(np.array([0.0, -0.0, nan, np.inf], np.float32), 1),
                        (np.array([0.0, -0.0, 0.25, -0.5, nan, np.inf], np.float32), 2)]:
        pos = b.u.fallback_positions(ch)
        m = b.u.stereo_matrix(ch, pos)
        for side, chan in ((0, capi.CHANNEL_LEFT), (1, capi.CHANNEL_RIGHT)):
            got = b.u.downmix(samples, ch, pos, chan)
            exp = []
            with np.errstate(invalid="ignore"):
                for fr in samples.reshape(-1, ch):
                    acc = np.float32(0.0)
                    for c in range(ch):
                        acc = np.float32(acc + np.float32(fr[c] * m[c][side]))
                    exp.append(acc)
            exp = np.array(exp, np.float32)
            g, e = got.view(np.uint32), exp.view(np.uint32)
            for gi, ei, ev in zip(g, e, exp):
                if np.isnan(ev):  # NaN payload propagation is platform-defined; NaN-ness must hold
                    assert np.isnan(np.array([gi], np.uint32).view(np.float32)[0])
                else:
                    assert gi == ei


def kat_mid_side_projection(b):
    """util/audio/channel.rs:12-21 via spectrum/processor.rs:480-493 (Left / Side of [1,0],[0,1])."""
    s = np.array([1.0, 0.0, 0.0, 1.0], np.float32)
    pos = b.u.fallback_positions(2)
    assert list(b.u.downmix(s, 2, pos, capi.CHANNEL_LEFT)) == [1.0, 0.0]
    assert list(b.u.downmix(s, 2, pos, capi.CHANNEL_SIDE)) == [0.5, -0.5]
    assert list(b.u.downmix(s, 2, pos, capi.CHANNEL_MID)) == [0.5, 0.5]
    assert list(b.u.downmix(s, 2, pos, capi.CHANNEL_NONE)) == [0.0, 0.0]


# ------------------------------------------------------------------ loudness
BS1770_48K_SHELF_B = [1.53512485958697, -2.69169618940638, 1.19839281085285]
BS1770_48K_SHELF_A = [1.0, -1.69065929318241, 0.73248077421585]
BS1770_48K_HP_B = [1.0, -2.0, 1.0]
BS1770_48K_HP_A = [1.0, -1.99004745483398, 0.99007225036621]


def kat_k_weighting_matches_bs1770_table(b):
    """loudness/processor.rs:22-55 vs the ITU-R BS.1770 48 kHz coefficient table."""
    bb, aa = b.u.k_weighting(48000.0)
    eb = np.convolve(BS1770_48K_SHELF_B, BS1770_48K_HP_B)
    ea = np.convolve(BS1770_48K_SHELF_A, BS1770_48K_HP_A)
    np.testing.assert_allclose(bb, eb, atol=1e-10)
    np.testing.assert_allclose(aa, ea, atol=1e-10)


def _k_gain_sq(b, sr, f):
    from scipy.signal import freqz

    bb, aa = b.u.k_weighting(float(sr))
    _, h = freqz(bb, aa, worN=[f], fs=sr)
    return float(np.abs(h[0]) ** 2)


def kat_silence_respects_configured_floor(b):
    """loudness/processor.rs:338-350"""
    snap = b.Loudness(LoudnessConfig(floor_db=-140.0)).process_block(AudioBlock(np.zeros(2048, np.float32), 2, 48000.0))
    assert snap is not None
    assert snap.short_term_loudness == -140.0
    assert list(snap.rms_fast_db[:2]) == [-140.0, -140.0]
    assert snap.channel_count == 2


def kat_rms_tracks_amplitude(b):
    """loudness/processor.rs:352-364"""
    def measure(amp):
        x = sine_wave(1000.0, 48000.0, int(48000 * 3.0), amp)
        return b.Loudness(LoudnessConfig()).process_block(AudioBlock(x, 1, 48000.0)).rms_fast_db[0]
    d = measure(0.5) - measure(0.25)
    assert 5.8 < d < 6.3, d


def kat_short_term_matches_analytic_k_weighted_sine(b):
    """loudness/processor.rs:366-398 (`processor_matches_ebur128_short_term`) — the ebur128 crate is
    absent, so the oracle for this KAT is the stationary-sine closed form
    LUFS = -0.691 + 10 log10( sum_c w_c * A^2/2 * |K(f)|^2 ), with |K| taken in the frequency
    domain (scipy.freqz) from the filter coefficients; same 0.001 LU band as the reference test
    (+ the window not being a whole number of periods: < 2e-4 LU)."""
    from openmeters_b200 import _capi as c_
    for sr in (44100.0, 48000.0, 96000.0):
        for ch in (2, 4, 5, 6):
            mono = sine_wave(1000.0, sr, int(np.float32(sr) * np.float32(4.0)), 0.5)
            inter = np.repeat(mono, ch)
            snap = b.Loudness(LoudnessConfig(sample_rate=sr)).process_block(AudioBlock(inter, ch, sr))
            pos = b.u.fallback_positions(ch)
            wsum = sum(0.0 if p == c_.POS_LOW_FREQUENCY else (1.41 if p in (4, 5, 6, 7) else 1.0) for p in pos[:ch])
            # the sine generator's f32 phase drifts slightly from 1 kHz; measure mean-square of the actual input
            ms = float(np.mean(mono[-int(sr * 3):].astype(np.float64) ** 2))
            exp = -0.691 + 10 * np.log10(wsum * ms * _k_gain_sq(b, sr, 1000.0))
            assert abs(snap.short_term_loudness - exp) < 2e-3, (sr, ch, snap.short_term_loudness, exp)


def kat_short_term_matches_scipy_time_domain(b):
    """Second, time-domain cross-check of the same reference test: scipy.signal.lfilter with the
    BS.1770 48 kHz table coefficients, mean square over the last 3 s / 0.4 s."""
    from scipy.signal import lfilter
    sr = 48000.0
    rng = np.random.default_rng(7)
    x = (0.3 * np.sin(2 * np.pi * 997.0 * np.arange(int(sr * 4)) / sr) + 0.05 * rng.uniform(-1, 1, int(sr * 4))).astype(np.float32)
    snap = b.Loudness(LoudnessConfig()).process_block(AudioBlock(np.repeat(x, 2), 2, sr))
    y = lfilter(np.convolve(BS1770_48K_SHELF_B, BS1770_48K_HP_B), np.convolve(BS1770_48K_SHELF_A, BS1770_48K_HP_A),
                x.astype(np.float64))
    st = -0.691 + 10 * np.log10(2 * np.mean(y[-144000:] ** 2))
    mo = -0.691 + 10 * np.log10(2 * np.mean(y[-19200:] ** 2))
    assert abs(snap.short_term_loudness - st) < 1e-3, (snap.short_term_loudness, st)
    assert abs(snap.momentary_loudness - mo) < 1e-3, (snap.momentary_loudness, mo)
    fast = 10 * np.log10(np.mean(y[-14400:] ** 2))
    slow = 10 * np.log10(np.mean(y[-48000:] ** 2))
    assert abs(snap.rms_fast_db[0] - fast) < 1e-3 and abs(snap.rms_slow_db[1] - slow) < 1e-3


def _true_peak_reference(x, sr):
    """libebur128-style interpolator restated independently in f64: 49-tap Hann-windowed sinc
    (loudness/processor.rs:73-97 describes it), factor 4 below 96 kHz, 2 below 192 kHz."""
    factor = 4 if sr < 96000 else (2 if sr < 192000 else 1)
    peak = float(np.max(np.abs(x)))
    if factor == 1:
        return peak
    taps = 49
    j = np.arange(taps)
    m = j - (taps - 1) / 2
    win = 0.5 * (1 - np.cos(2 * np.pi * j / (taps - 1)))
    with np.errstate(invalid="ignore", divide="ignore"):
        h = np.where(m == 0, 1.0, win * np.sin(m * np.pi / factor) / (m * np.pi / factor))
    up = np.zeros(len(x) * factor)
    up[::factor] = x.astype(np.float64)
    y = np.convolve(up, h)[: len(up)]  # causal, like the streaming delay line
    return max(peak, float(np.max(np.abs(y))))


def kat_true_peak_matches_interpolator_at_standard_rates(b):
    """loudness/processor.rs:426-454 (`true_peak_matches_ebur128_at_standard_rates`)"""
    for sr in (48000.0, 96000.0, 192000.0):
        x = sine_wave(17000.0, sr, int(np.float32(sr) * np.float32(0.01)), 0.9)
        ours = b.Loudness(LoudnessConfig(sample_rate=sr)).process_block(AudioBlock(x, 1, sr)).true_peak_db[0]
        exp = 20 * np.log10(_true_peak_reference(x, sr))
        assert abs(ours - exp) < 1e-3, (sr, ours, exp)


def kat_lfe_and_surround_channel_weights(b):
    """loudness/processor.rs:419-424 (`fallback_channel_weights...`) behaviourally: 6ch layout
    FL FR FC LFE RL RR -> weights 1,1,1,0,1.41,1.41."""
    sr = 48000.0
    mono = sine_wave(1000.0, sr, int(sr * 3.5), 0.25)
    g = _k_gain_sq(b, sr, 1000.0)
    ms = float(np.mean(mono[-144000:].astype(np.float64) ** 2))
    for active, w in [((3,), 0.0), ((4,), 1.41), ((0, 4, 5), 1 + 2.82), ((0, 1, 2, 3, 4, 5), 3 + 2.82)]:
        inter = np.zeros((len(mono), 6), np.float32)
        for c in active:
            inter[:, c] = mono
        snap = b.Loudness(LoudnessConfig()).process_block(AudioBlock(inter.reshape(-1), 6, sr))
        if w == 0.0:
            assert snap.short_term_loudness == np.float32(-99.9)
        else:
            exp = -0.691 + 10 * np.log10(w * ms * g)
            assert abs(snap.short_term_loudness - exp) < 2e-3, (active, snap.short_term_loudness, exp)


def kat_true_peak_resets_each_block(b):
    """loudness/processor.rs:301 — peak is taken per process_block call."""
    p = b.Loudness(LoudnessConfig())
    loud = sine_wave(1000.0, 48000.0, 1024, 0.9)
    quiet = sine_wave(1000.0, 48000.0, 1024, 0.01)
    a = p.process_block(AudioBlock(loud, 1, 48000.0)).true_peak_db[0]
    c = p.process_block(AudioBlock(quiet, 1, 48000.0)).true_peak_db[0]
    d = p.process_block(AudioBlock(quiet, 1, 48000.0)).true_peak_db[0]
    # (block `c` still sees the loud tail in the 12-tap delay line, so only a and d are pinned)
    assert a > -1.5 and d < -39.0 and c > d
