//! Safe wrapper over `omb200-sys`: the three processor types of OpenMeters' DSP hot path with the reference's method set
//! (`src/visuals/registry.rs:100-118`), backed by hand-written sm_100a CUDA kernels.  No CPU fallback: without a usable
//! device every call panics with the library's error text (the reference's release profile is `panic = "abort"` and its
//! processors `expect()` their invariants, `spectrogram/processor.rs:319-321`).
//!
//! The shared value types below mirror `src/dsp.rs` and `src/util/audio/*` name for name so that `registry.rs`, `state.rs`
//! and `render.rs` compile unchanged against this crate.
use omb200_sys as sys;
use std::ffi::CStr;

pub mod loudness;
pub mod spectrogram;
pub mod spectrum;

pub use loudness::{LoudnessConfig, LoudnessProcessor, LoudnessSnapshot};
pub use spectrogram::{SpectrogramColumn, SpectrogramConfig, SpectrogramPoint, SpectrogramProcessor, SpectrogramUpdate};
pub use spectrum::{AveragingMode, SpectrumConfig, SpectrumProcessor, SpectrumSnapshot};

/// `src/dsp.rs:6`
pub const MAX_AUDIO_CHANNELS: usize = sys::OMB_MAX_CHANNELS as usize;
/// `src/util/audio.rs` DEFAULT_SAMPLE_RATE
pub const DEFAULT_SAMPLE_RATE: f32 = 48_000.0;

/// `src/dsp.rs:8-22`
#[derive(Debug, Clone, Copy, Default, PartialEq, Eq, Hash)]
pub enum ChannelPosition {
    FrontLeft,
    FrontRight,
    FrontCenter,
    LowFrequency,
    RearLeft,
    RearRight,
    SideLeft,
    SideRight,
    Mono,
    Aux(u8),
    #[default]
    Unknown,
}

impl ChannelPosition {
    /// `src/dsp.rs:25-34`
    pub const SURROUND: [Self; MAX_AUDIO_CHANNELS] = [
        Self::FrontLeft,
        Self::FrontRight,
        Self::FrontCenter,
        Self::LowFrequency,
        Self::RearLeft,
        Self::RearRight,
        Self::SideLeft,
        Self::SideRight,
    ];

    /// `src/dsp.rs:36-47`, computed by the library (`omb_fallback_positions`).
    pub fn fallback(channels: usize) -> [Self; MAX_AUDIO_CHANNELS] {
        let mut codes = [0u8; MAX_AUDIO_CHANNELS];
        unsafe { sys::omb_fallback_positions(channels.min(MAX_AUDIO_CHANNELS) as u32, codes.as_mut_ptr()) };
        codes.map(Self::from_code)
    }

    pub(crate) fn code(self) -> u8 {
        match self {
            Self::FrontLeft => sys::OMB_POS_FRONT_LEFT as u8,
            Self::FrontRight => sys::OMB_POS_FRONT_RIGHT as u8,
            Self::FrontCenter => sys::OMB_POS_FRONT_CENTER as u8,
            Self::LowFrequency => sys::OMB_POS_LOW_FREQUENCY as u8,
            Self::RearLeft => sys::OMB_POS_REAR_LEFT as u8,
            Self::RearRight => sys::OMB_POS_REAR_RIGHT as u8,
            Self::SideLeft => sys::OMB_POS_SIDE_LEFT as u8,
            Self::SideRight => sys::OMB_POS_SIDE_RIGHT as u8,
            Self::Mono => sys::OMB_POS_MONO as u8,
            Self::Unknown => sys::OMB_POS_UNKNOWN as u8,
            Self::Aux(i) => sys::OMB_POS_AUX0 as u8 + i.min(7),
        }
    }

    pub(crate) fn from_code(code: u8) -> Self {
        match code as i32 {
            sys::OMB_POS_FRONT_LEFT => Self::FrontLeft,
            sys::OMB_POS_FRONT_RIGHT => Self::FrontRight,
            sys::OMB_POS_FRONT_CENTER => Self::FrontCenter,
            sys::OMB_POS_LOW_FREQUENCY => Self::LowFrequency,
            sys::OMB_POS_REAR_LEFT => Self::RearLeft,
            sys::OMB_POS_REAR_RIGHT => Self::RearRight,
            sys::OMB_POS_SIDE_LEFT => Self::SideLeft,
            sys::OMB_POS_SIDE_RIGHT => Self::SideRight,
            sys::OMB_POS_MONO => Self::Mono,
            c if c >= sys::OMB_POS_AUX0 => Self::Aux((c - sys::OMB_POS_AUX0) as u8),
            _ => Self::Unknown,
        }
    }
}

/// `src/dsp.rs:108-115`: interleaved f32 view.  The stereo fold-down matrix the reference precomputes here
/// (`with_positions`, `:190-213`) is built on the device side from the same fields.
pub struct AudioBlock<'a> {
    pub samples: &'a [f32],
    pub channels: usize,
    pub sample_rate: f32,
    pub positions: [ChannelPosition; MAX_AUDIO_CHANNELS],
}

impl<'a> AudioBlock<'a> {
    /// `src/dsp.rs:180-188`
    pub fn new(samples: &'a [f32], channels: usize, sample_rate: f32) -> Self {
        let channels = channels.max(1);
        Self::with_positions(samples, channels, sample_rate, ChannelPosition::fallback(channels))
    }
    /// `src/dsp.rs:190-213`
    pub fn with_positions(samples: &'a [f32], channels: usize, sample_rate: f32, positions: [ChannelPosition; MAX_AUDIO_CHANNELS]) -> Self {
        Self { samples, channels: channels.max(1), sample_rate, positions }
    }
    pub fn frame_count(&self) -> usize {
        self.samples.len() / self.channels.max(1)
    }
    pub fn is_empty(&self) -> bool {
        self.frame_count() == 0
    }
    pub(crate) fn position_codes(&self) -> [u8; MAX_AUDIO_CHANNELS] {
        self.positions.map(ChannelPosition::code)
    }
}

/// `src/util/audio/window.rs:9-18`
#[derive(Debug, Clone, Copy, PartialEq, Eq, Hash)]
pub enum WindowKind {
    Rectangular,
    Hann,
    Hamming,
    Blackman,
    BlackmanHarris,
}

impl WindowKind {
    pub(crate) fn code(self) -> u32 {
        self as u32 // same order as omb_window_kind
    }
    pub(crate) fn from_code(c: u32) -> Self {
        [Self::Rectangular, Self::Hann, Self::Hamming, Self::Blackman, Self::BlackmanHarris][(c as usize).min(4)]
    }
}

/// `src/util/audio/channel.rs:4-10`
#[derive(Debug, Clone, Copy, PartialEq, Eq, Hash)]
pub enum Channel {
    Left,
    Right,
    Mid,
    Side,
    None,
}

impl Channel {
    pub(crate) fn code(self) -> u32 {
        self as u32 // same order as omb_channel
    }
    pub(crate) fn from_code(c: u32) -> Self {
        [Self::Left, Self::Right, Self::Mid, Self::Side, Self::None][(c as usize).min(4)]
    }
}

pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::omb_last_error()) }.to_string_lossy().into_owned()
}

/// 0 -> Some(()), OMB_NO_DATA -> None, errors panic (see the crate docs).
pub(crate) fn status(rc: i32, what: &str) -> Option<()> {
    match rc {
        sys::OMB_OK => Some(()),
        sys::OMB_NO_DATA => None,
        _ => panic!("{what} failed ({rc}): {}", last_error()),
    }
}
