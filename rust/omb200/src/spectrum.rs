//! `SpectrumProcessor` — drop-in for `src/visuals/spectrum/processor.rs` (real FFT, smoothing, A-weighted + raw dB traces).
use crate::{status, sys, AudioBlock, Channel, WindowKind, DEFAULT_SAMPLE_RATE};
use std::ptr::NonNull;

/// `spectrum/processor.rs:64-70`
#[derive(Debug, Clone, Copy)]
pub enum AveragingMode {
    None,
    Exponential { factor: f32 },
    PeakHold { decay_per_second: f32 },
}

/// `spectrum/processor.rs:39-51`
#[derive(Debug, Clone, Copy)]
pub struct SpectrumConfig {
    pub sample_rate: f32,
    pub fft_size: usize,
    pub hop_size: usize,
    pub window: WindowKind,
    pub averaging: AveragingMode,
    pub source: Channel,
    pub secondary_source: Channel,
    pub floor_db: f32,
}

impl Default for SpectrumConfig {
    fn default() -> Self {
        Self { sample_rate: DEFAULT_SAMPLE_RATE, fft_size: 16_384, hop_size: 16_384 / 16, window: WindowKind::Hann,
               averaging: AveragingMode::None, source: Channel::Mid, secondary_source: Channel::None, floor_db: -100.0 }
    }
}

impl SpectrumConfig {
    fn to_c(self) -> sys::omb_spectrum_config {
        let (averaging, averaging_param) = match self.averaging {
            AveragingMode::None => (sys::OMB_AVG_NONE as u32, 0.0),
            AveragingMode::Exponential { factor } => (sys::OMB_AVG_EXPONENTIAL as u32, factor),
            AveragingMode::PeakHold { decay_per_second } => (sys::OMB_AVG_PEAK_HOLD as u32, decay_per_second),
        };
        sys::omb_spectrum_config {
            sample_rate: self.sample_rate, window: self.window.code(), fft_size: self.fft_size as u64, hop_size: self.hop_size as u64,
            averaging, averaging_param, source: self.source.code(), secondary_source: self.secondary_source.code(),
            floor_db: self.floor_db, _pad: 0,
        }
    }
    fn from_c(c: &sys::omb_spectrum_config) -> Self {
        let averaging = match c.averaging as i32 {
            sys::OMB_AVG_EXPONENTIAL => AveragingMode::Exponential { factor: c.averaging_param },
            sys::OMB_AVG_PEAK_HOLD => AveragingMode::PeakHold { decay_per_second: c.averaging_param },
            _ => AveragingMode::None,
        };
        Self { sample_rate: c.sample_rate, fft_size: c.fft_size as usize, hop_size: c.hop_size as usize, window: WindowKind::from_code(c.window),
               averaging, source: Channel::from_code(c.source), secondary_source: Channel::from_code(c.secondary_source), floor_db: c.floor_db }
    }
}

/// `spectrum/processor.rs:31-37`: `traces[trace][0]` = weighted, `[trace][1]` = raw.
#[derive(Debug, Clone, Default)]
pub struct SpectrumSnapshot {
    pub frequency_bins: Vec<f32>,
    pub traces: [[Vec<f32>; 2]; 2],
}

pub struct SpectrumProcessor {
    h: NonNull<sys::omb_spectrum>,
    snapshot: SpectrumSnapshot, // refilled from the library-owned buffers; `process_block` lends it like the reference does
}

impl SpectrumProcessor {
    /// `processor.rs:89`
    pub fn new(config: SpectrumConfig) -> Self {
        let mut h = std::ptr::null_mut();
        status(unsafe { sys::omb_spectrum_create(&config.to_c(), &mut h) }, "omb_spectrum_create");
        Self { h: NonNull::new(h).expect("omb_spectrum_create returned null"), snapshot: SpectrumSnapshot::default() }
    }
    /// `processor.rs:108`
    pub fn config(&self) -> SpectrumConfig {
        let mut c = SpectrumConfig::default().to_c();
        status(unsafe { sys::omb_spectrum_get_config(self.h.as_ptr(), &mut c) }, "omb_spectrum_get_config");
        SpectrumConfig::from_c(&c)
    }
    /// `processor.rs:300-322`
    pub fn update_config(&mut self, config: SpectrumConfig) {
        status(unsafe { sys::omb_spectrum_update_config(self.h.as_ptr(), &config.to_c()) }, "omb_spectrum_update_config");
    }
    /// `processor.rs:120-124`
    pub fn prepare(&mut self) {
        status(unsafe { sys::omb_spectrum_prepare(self.h.as_ptr()) }, "omb_spectrum_prepare");
    }
    /// `processor.rs:112-118`
    pub fn reset_audio(&mut self) {
        status(unsafe { sys::omb_spectrum_reset_audio(self.h.as_ptr()) }, "omb_spectrum_reset_audio");
    }
    /// `processor.rs:255-269` — lends the snapshot of the LAST hop of the block (the reference overwrites it per hop).
    pub fn process_block(&mut self, block: &AudioBlock<'_>) -> Option<&SpectrumSnapshot> {
        if block.is_empty() {
            return None;
        }
        let mut snap = std::mem::MaybeUninit::<sys::omb_spectrum_snapshot>::zeroed();
        let pos = block.position_codes();
        let rc = unsafe {
            sys::omb_spectrum_process_block(self.h.as_ptr(), block.samples.as_ptr(), block.samples.len(), block.channels as u32,
                                            block.sample_rate, pos.as_ptr(), snap.as_mut_ptr())
        };
        status(rc, "omb_spectrum_process_block")?;
        let snap = unsafe { snap.assume_init() };
        let bins = snap.bins as usize;
        let copy = |p: *const f32, dst: &mut Vec<f32>| {
            dst.clear();
            if !p.is_null() {
                dst.extend_from_slice(unsafe { std::slice::from_raw_parts(p, bins) });
            }
        };
        copy(snap.frequency_bins, &mut self.snapshot.frequency_bins);
        for t in 0..2 {
            for w in 0..2 {
                copy(snap.traces[t][w], &mut self.snapshot.traces[t][w]);
            }
        }
        Some(&self.snapshot)
    }
}

impl Drop for SpectrumProcessor {
    fn drop(&mut self) {
        unsafe { sys::omb_spectrum_destroy(self.h.as_ptr()) }
    }
}
