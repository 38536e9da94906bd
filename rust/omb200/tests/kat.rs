//! Three of the reference's own `#[test]`s restated against the drop-in types (they need a B200: no CPU fallback).
use omb200::*;

/// `src/util/audio.rs:28-33`
fn sine_wave(freq: f32, sample_rate: f32, count: usize, amplitude: f32) -> Vec<f32> {
    (0..count).map(|i| (core::f32::consts::TAU * freq * i as f32 / sample_rate).sin() * amplitude).collect()
}

/// `spectrogram/processor.rs:709-724` detects_sine_frequency_peak
#[test]
fn detects_sine_frequency_peak() {
    let cfg = SpectrogramConfig { fft_size: 1024, hop_size: 512, history_length: 8, use_reassignment: false, window: WindowKind::Hann,
                                  ..SpectrogramConfig::default() };
    let freq = 200.0 * cfg.sample_rate / cfg.fft_size as f32;
    let samples = sine_wave(freq, cfg.sample_rate, 2048, 1.0);
    let mut p = SpectrogramProcessor::new(cfg);
    let up = p.process_block(&AudioBlock::new(&samples, 1, cfg.sample_rate)).expect("expected snapshot");
    let SpectrogramColumn::Classic(mags) = up.new_columns.last().unwrap() else { panic!("classic column expected") };
    let (idx, _) = mags.iter().enumerate().max_by_key(|(_, m)| **m).unwrap();
    assert_eq!(idx, 200);
}

/// `spectrum/processor.rs:538-563` peak_hold_decays_for_each_audio_hop_in_large_batch
#[test]
fn peak_hold_decays_for_each_audio_hop_in_large_batch() {
    let mut p = SpectrumProcessor::new(SpectrumConfig {
        sample_rate: 8.0, fft_size: 8, hop_size: 8, window: WindowKind::Rectangular,
        averaging: AveragingMode::PeakHold { decay_per_second: 24.0 }, floor_db: -100.0, ..SpectrumConfig::default()
    });
    let mut samples = sine_wave(1.0, 8.0, 8, 1.0);
    samples.extend([0.0; 8]);
    let snap = p.process_block(&AudioBlock::new(&samples, 1, 8.0)).expect("expected snapshot");
    let held_db = snap.traces[0][1][1];
    assert!((-24.1..-23.9).contains(&held_db), "held peak should decay once per hop, got {held_db} dB");
}

/// `loudness/processor.rs:338-350` silence_respects_configured_floor
#[test]
fn silence_respects_configured_floor() {
    let samples = [0.0f32; 2048];
    let snapshot = LoudnessProcessor::new(LoudnessConfig { floor_db: -140.0, ..Default::default() })
        .process_block(&AudioBlock::new(&samples, 2, DEFAULT_SAMPLE_RATE))
        .expect("expected snapshot");
    assert_eq!(snapshot.short_term_loudness, -140.0);
    assert_eq!(snapshot.rms_fast_db[..2], [-140.0; 2]);
}
