//! `SpectrogramProcessor` — drop-in for `src/visuals/spectrogram/processor.rs` (classic + reassigned STFT columns).
use crate::{status, sys, AudioBlock, WindowKind, DEFAULT_SAMPLE_RATE};
use std::ptr::NonNull;

/// `spectrogram/processor.rs:37-43` — `#[repr(C)]`, 12 bytes, identical to `omb_spectrogram_point`.
#[repr(C)]
#[derive(Debug, Clone, Copy, PartialEq)]
pub struct SpectrogramPoint {
    pub time_offset: f32,
    pub freq_hz: f32,
    pub power: f32,
}

/// `spectrogram/processor.rs:45-57`
#[derive(Debug, Clone, Copy)]
pub struct SpectrogramConfig {
    pub sample_rate: f32,
    pub fft_size: usize,
    pub hop_size: usize,
    pub window: WindowKind,
    pub history_length: usize,
    pub use_reassignment: bool,
    pub zero_padding_factor: usize,
}

impl Default for SpectrogramConfig {
    fn default() -> Self {
        Self { sample_rate: DEFAULT_SAMPLE_RATE, fft_size: 2048, hop_size: 64, window: WindowKind::Hann, history_length: 0,
               use_reassignment: true, zero_padding_factor: 1 }
    }
}

impl SpectrogramConfig {
    fn to_c(self) -> sys::omb_spectrogram_config {
        sys::omb_spectrogram_config {
            sample_rate: self.sample_rate, window: self.window.code(), fft_size: self.fft_size as u64, hop_size: self.hop_size as u64,
            history_length: self.history_length as u64, zero_padding_factor: self.zero_padding_factor as u64,
            use_reassignment: self.use_reassignment as i32, _pad: 0,
        }
    }
    fn from_c(c: &sys::omb_spectrogram_config) -> Self {
        Self { sample_rate: c.sample_rate, fft_size: c.fft_size as usize, hop_size: c.hop_size as usize, window: WindowKind::from_code(c.window),
               history_length: c.history_length as usize, use_reassignment: c.use_reassignment != 0,
               zero_padding_factor: c.zero_padding_factor as usize }
    }
}

/// `spectrogram/processor.rs:122-127`
#[derive(Debug)]
pub enum SpectrogramColumn {
    Reassigned(Vec<SpectrogramPoint>),
    Classic(Vec<u16>),
}

/// `spectrogram/processor.rs:160-168`
pub struct SpectrogramUpdate {
    pub fft_size: usize,
    pub hop_size: usize,
    pub sample_rate: f32,
    pub history_length: usize,
    pub reset: bool,
    pub reassigned_power_scale: f32,
    pub new_columns: Vec<SpectrogramColumn>,
}

/// `!Send` / `!Sync` like the reference's processor (it lives in `Rc<RefCell<_>>`, `registry.rs:23`): the handle is a raw pointer.
pub struct SpectrogramProcessor {
    h: NonNull<sys::omb_spectrogram>,
}

impl SpectrogramProcessor {
    /// `processor.rs:188` — the config is normalised, never rejected (`:71-82`).
    pub fn new(cfg: SpectrogramConfig) -> Self {
        let mut h = std::ptr::null_mut();
        status(unsafe { sys::omb_spectrogram_create(&cfg.to_c(), &mut h) }, "omb_spectrogram_create");
        Self { h: NonNull::new(h).expect("omb_spectrogram_create returned null") }
    }
    /// `processor.rs:208`
    pub fn config(&self) -> SpectrogramConfig {
        let mut c = SpectrogramConfig::default().to_c();
        status(unsafe { sys::omb_spectrogram_get_config(self.h.as_ptr(), &mut c) }, "omb_spectrogram_get_config");
        SpectrogramConfig::from_c(&c)
    }
    /// `processor.rs:518-543`
    pub fn update_config(&mut self, cfg: SpectrogramConfig) {
        status(unsafe { sys::omb_spectrogram_update_config(self.h.as_ptr(), &cfg.to_c()) }, "omb_spectrogram_update_config");
    }
    /// `processor.rs:219-223`
    pub fn prepare(&mut self) {
        status(unsafe { sys::omb_spectrogram_prepare(self.h.as_ptr()) }, "omb_spectrogram_prepare");
    }
    /// `processor.rs:212-217`
    pub fn reset_audio(&mut self) {
        status(unsafe { sys::omb_spectrogram_reset_audio(self.h.as_ptr()) }, "omb_spectrogram_reset_audio");
    }
    /// `processor.rs:490-516` — `None` when no column became ready (`OMB_NO_DATA`).
    pub fn process_block(&mut self, block: &AudioBlock<'_>) -> Option<SpectrogramUpdate> {
        if block.is_empty() {
            return None;
        }
        let mut up = std::mem::MaybeUninit::<sys::omb_spectrogram_update>::zeroed();
        let pos = block.position_codes();
        let rc = unsafe {
            sys::omb_spectrogram_process_block(self.h.as_ptr(), block.samples.as_ptr(), block.samples.len(), block.channels as u32,
                                               block.sample_rate, pos.as_ptr(), up.as_mut_ptr())
        };
        status(rc, "omb_spectrogram_process_block")?;
        // library-owned buffers, valid until the next call on this handle: copied out here
        let up = unsafe { up.assume_init() };
        let n = up.n_columns as usize;
        let bins = up.bins as usize;
        let new_columns = if up.kind == sys::OMB_COLUMN_REASSIGNED {
            let offs = unsafe { std::slice::from_raw_parts(up.column_offsets, n + 1) };
            let pts = unsafe { std::slice::from_raw_parts(up.points as *const SpectrogramPoint, offs[n] as usize) };
            (0..n).map(|c| SpectrogramColumn::Reassigned(pts[offs[c] as usize..offs[c + 1] as usize].to_vec())).collect()
        } else {
            let codes = unsafe { std::slice::from_raw_parts(up.classic_db, n * bins) };
            (0..n).map(|c| SpectrogramColumn::Classic(codes[c * bins..(c + 1) * bins].to_vec())).collect()
        };
        Some(SpectrogramUpdate {
            fft_size: up.fft_size as usize, hop_size: up.hop_size as usize, sample_rate: up.sample_rate,
            history_length: up.history_length as usize, reset: up.reset != 0, reassigned_power_scale: up.reassigned_power_scale,
            new_columns,
        })
    }
}

impl Drop for SpectrogramProcessor {
    fn drop(&mut self) {
        unsafe { sys::omb_spectrogram_destroy(self.h.as_ptr()) }
    }
}
