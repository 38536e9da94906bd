//! `LoudnessProcessor` — drop-in for `src/visuals/loudness/processor.rs` (BS.1770 K-weighting, sliding LUFS / RMS windows, true peak).
use crate::{status, sys, AudioBlock, ChannelPosition, DEFAULT_SAMPLE_RATE, MAX_AUDIO_CHANNELS};
use std::ptr::NonNull;

/// `loudness/processor.rs:164` MAX_CHANNELS
pub const MAX_CHANNELS: usize = MAX_AUDIO_CHANNELS;

/// `loudness/processor.rs:210-216`
#[derive(Debug, Clone, Copy)]
pub struct LoudnessConfig {
    pub sample_rate: f32,
    pub floor_db: f32,
}

impl Default for LoudnessConfig {
    fn default() -> Self {
        Self { sample_rate: DEFAULT_SAMPLE_RATE, floor_db: -99.9 }
    }
}

/// `loudness/processor.rs:185-194` — `Copy`, field for field `omb_loudness_snapshot`.
#[derive(Debug, Clone, Copy, Default, PartialEq)]
pub struct LoudnessSnapshot {
    pub short_term_loudness: f32,
    pub momentary_loudness: f32,
    pub rms_fast_db: [f32; MAX_CHANNELS],
    pub rms_slow_db: [f32; MAX_CHANNELS],
    pub true_peak_db: [f32; MAX_CHANNELS],
    pub channel_count: usize,
    pub positions: [ChannelPosition; MAX_CHANNELS],
}

pub struct LoudnessProcessor {
    h: NonNull<sys::omb_loudness>,
}

impl LoudnessProcessor {
    /// `processor.rs:225`
    pub fn new(config: LoudnessConfig) -> Self {
        let c = sys::omb_loudness_config { sample_rate: config.sample_rate, floor_db: config.floor_db };
        let mut h = std::ptr::null_mut();
        status(unsafe { sys::omb_loudness_create(&c, &mut h) }, "omb_loudness_create");
        Self { h: NonNull::new(h).expect("omb_loudness_create returned null") }
    }
    pub fn config(&self) -> LoudnessConfig {
        let mut c = sys::omb_loudness_config { sample_rate: 0.0, floor_db: 0.0 };
        status(unsafe { sys::omb_loudness_get_config(self.h.as_ptr(), &mut c) }, "omb_loudness_get_config");
        LoudnessConfig { sample_rate: c.sample_rate, floor_db: c.floor_db }
    }
    /// `processor.rs:234`
    pub fn reset_audio(&mut self) {
        status(unsafe { sys::omb_loudness_reset_audio(self.h.as_ptr()) }, "omb_loudness_reset_audio");
    }
    /// `processor.rs:253-311` — one snapshot per call; true peak is the maximum over this block only.
    pub fn process_block(&mut self, block: &AudioBlock<'_>) -> Option<LoudnessSnapshot> {
        if block.is_empty() {
            return None;
        }
        let mut s = std::mem::MaybeUninit::<sys::omb_loudness_snapshot>::zeroed();
        let pos = block.position_codes();
        let rc = unsafe {
            sys::omb_loudness_process_block(self.h.as_ptr(), block.samples.as_ptr(), block.samples.len(), block.channels as u32,
                                            block.sample_rate, pos.as_ptr(), s.as_mut_ptr())
        };
        status(rc, "omb_loudness_process_block")?;
        let s = unsafe { s.assume_init() };
        Some(LoudnessSnapshot {
            short_term_loudness: s.short_term_loudness, momentary_loudness: s.momentary_loudness, rms_fast_db: s.rms_fast_db,
            rms_slow_db: s.rms_slow_db, true_peak_db: s.true_peak_db, channel_count: s.channel_count as usize,
            positions: s.positions.map(ChannelPosition::from_code),
        })
    }
}

impl Drop for LoudnessProcessor {
    fn drop(&mut self) {
        unsafe { sys::omb_loudness_destroy(self.h.as_ptr()) }
    }
}
