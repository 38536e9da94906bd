"""ctypes declarations for the C ABI in include/omb200.h.

`bind(lib, prefix)` attaches argtypes/restypes for every entry point the header
declares and returns a small namespace object whose attributes are the
functions with the prefix stripped (``api.spectrogram_create`` ...).  The
product binds ``libomb200.so`` with prefix ``omb_``; the test-only CPU oracle
exports the same signatures under ``ombo_`` and is bound by ``oracle/oracle_py.py``
— never from inside this package.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace

MAX_CHANNELS = 8

OK, NO_DATA = 0, 1
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM = -1, -2, -3, -4

WINDOW_RECTANGULAR, WINDOW_HANN, WINDOW_HAMMING, WINDOW_BLACKMAN, WINDOW_BLACKMAN_HARRIS = range(5)
CHANNEL_LEFT, CHANNEL_RIGHT, CHANNEL_MID, CHANNEL_SIDE, CHANNEL_NONE = range(5)
AVG_NONE, AVG_EXPONENTIAL, AVG_PEAK_HOLD = range(3)
COLUMN_REASSIGNED, COLUMN_CLASSIC = 0, 1
KERNEL_AUTO, KERNEL_GENERIC, KERNEL_FAST = 0, 1, 2

POS_FRONT_LEFT, POS_FRONT_RIGHT, POS_FRONT_CENTER, POS_LOW_FREQUENCY = 0, 1, 2, 3
POS_REAR_LEFT, POS_REAR_RIGHT, POS_SIDE_LEFT, POS_SIDE_RIGHT = 4, 5, 6, 7
POS_MONO, POS_UNKNOWN, POS_AUX0 = 8, 9, 16
SURROUND = (0, 1, 2, 3, 4, 5, 6, 7)


class SpectrogramConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_float),
        ("window", C.c_uint32),
        ("fft_size", C.c_uint64),
        ("hop_size", C.c_uint64),
        ("history_length", C.c_uint64),
        ("zero_padding_factor", C.c_uint64),
        ("use_reassignment", C.c_int32),
        ("_pad", C.c_int32),
    ]


class SpectrogramPoint(C.Structure):
    _fields_ = [("time_offset", C.c_float), ("freq_hz", C.c_float), ("power", C.c_float)]


class SpectrogramUpdate(C.Structure):
    _fields_ = [
        ("fft_size", C.c_uint64),
        ("hop_size", C.c_uint64),
        ("history_length", C.c_uint64),
        ("sample_rate", C.c_float),
        ("reassigned_power_scale", C.c_float),
        ("reset", C.c_int32),
        ("kind", C.c_int32),
        ("n_columns", C.c_uint32),
        ("bins", C.c_uint32),
        ("column_offsets", C.POINTER(C.c_uint32)),
        ("points", C.POINTER(SpectrogramPoint)),
        ("classic_db", C.POINTER(C.c_uint16)),
    ]


class SpectrumConfig(C.Structure):
    _fields_ = [
        ("sample_rate", C.c_float),
        ("window", C.c_uint32),
        ("fft_size", C.c_uint64),
        ("hop_size", C.c_uint64),
        ("averaging", C.c_uint32),
        ("averaging_param", C.c_float),
        ("source", C.c_uint32),
        ("secondary_source", C.c_uint32),
        ("floor_db", C.c_float),
        ("_pad", C.c_int32),
    ]


class SpectrumSnapshot(C.Structure):
    _fields_ = [
        ("bins", C.c_uint32),
        ("_pad", C.c_int32),
        ("frequency_bins", C.POINTER(C.c_float)),
        ("traces", (C.POINTER(C.c_float) * 2) * 2),
    ]


class SpectrumPeakSpec(C.Structure):
    """omb_spectrum_peak_spec: trace 0 = A-weighted / 1 = raw; min_hz; max_hz <= 0 -> Nyquist (state.rs:106-107)."""
    _fields_ = [("trace", C.c_uint32), ("min_hz", C.c_float), ("max_hz", C.c_float)]


class SpectrumBankSnapshot(C.Structure):
    _fields_ = [("bins", C.c_uint32), ("n_streams", C.c_uint32), ("n_traces", C.c_uint32), ("trace_index", C.c_uint32 * 2),
                ("frequency_bins", C.POINTER(C.c_float)), ("weighted", C.POINTER(C.c_float)), ("raw", C.POINTER(C.c_float))]


class LoudnessConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_float), ("floor_db", C.c_float)]


class LoudnessSnapshot(C.Structure):
    _fields_ = [
        ("short_term_loudness", C.c_float),
        ("momentary_loudness", C.c_float),
        ("rms_fast_db", C.c_float * MAX_CHANNELS),
        ("rms_slow_db", C.c_float * MAX_CHANNELS),
        ("true_peak_db", C.c_float * MAX_CHANNELS),
        ("channel_count", C.c_uint32),
        ("positions", C.c_uint8 * MAX_CHANNELS),
    ]


class AudioFormat(C.Structure):
    _fields_ = [
        ("channels", C.c_uint32),
        ("sample_rate", C.c_float),
        ("generation", C.c_uint64),
        ("positions", C.c_uint8 * MAX_CHANNELS),
    ]


class SplatParams(C.Structure):
    _fields_ = [
        ("freq_scale", C.c_uint32),
        ("freq_min", C.c_float),
        ("freq_max", C.c_float),
        ("uv_y_range", C.c_float * 2),
        ("ext_w", C.c_float),
        ("ext_h", C.c_float),
        ("scale_factor", C.c_float),
        ("tilt_db", C.c_float),
        ("ring_capacity", C.c_uint32),
        ("newest_col", C.c_uint32),
        ("col_count", C.c_uint32),
        ("reassigned_power_scale", C.c_float),
    ]


class SpectrogramBankUpdate(C.Structure):
    _fields_ = [
        ("fft_size", C.c_uint64),
        ("hop_size", C.c_uint64),
        ("history_length", C.c_uint64),
        ("sample_rate", C.c_float),
        ("reassigned_power_scale", C.c_float),
        ("reset", C.c_int32),
        ("kind", C.c_int32),
        ("n_streams", C.c_uint32),
        ("n_columns", C.c_uint32),
        ("bins", C.c_uint32),
        ("_pad", C.c_uint32),
        ("counts", C.POINTER(C.c_uint32)),
        ("points", C.POINTER(SpectrogramPoint)),
        ("classic_db", C.POINTER(C.c_uint16)),
    ]


FREQ_LINEAR, FREQ_LOG, FREQ_ERB = 0, 1, 2
SPAN_PCM, SPAN_SILENCE, SPAN_RESET = 0, 1, 2
SPAN_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.c_size_t, C.c_uint64, C.POINTER(AudioFormat))
INGEST_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_float), C.c_size_t, C.POINTER(AudioFormat),
                        C.POINTER(SpectrogramUpdate), C.POINTER(SpectrumSnapshot), C.POINTER(LoudnessSnapshot))

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u32p = C.POINTER(C.c_uint32)
_i32p = C.POINTER(C.c_int32)
_vp = C.c_void_p
_sz = C.c_size_t
_u64 = C.c_uint64
_u32 = C.c_uint32

# name -> (restype, argtypes).  Exactly the entry points include/omb200.h declares.
HEADER_SYMBOLS = {
    "last_error": (C.c_char_p, []),
    "version": (C.c_char_p, []),
    "device_count": (C.c_int, []),
    "set_device": (C.c_int, [C.c_int]),
    "kernel_launch_count": (_u64, []),
    "probe_fp32_tflops": (C.c_int, [_f64p]),
    "peer_alloc": (C.c_int, [_sz, C.POINTER(_vp), _u8p]),
    "peer_open": (C.c_int, [_u8p, C.POINTER(_vp)]),
    "peer_close": (C.c_int, [_vp]),
    "peer_free": (C.c_int, [_vp]),
    "copy_async": (C.c_int, [_vp, _vp, _sz, _vp]),
    "window_coefficients": (C.c_int, [C.c_int, _sz, _f32p]),
    "fft_bin_normalization": (C.c_int, [_f32p, _sz, _sz, _f32p]),
    "reassignment_windows": (C.c_int, [_f32p, _sz, _f32p, _f32p]),
    "reassigned_power_scale": (C.c_float, [_f32p, _sz, _sz]),
    "pack_classic_db": (C.c_uint16, [C.c_float]),
    "a_weight": (C.c_float, [C.c_float]),
    "k_weighting_coefficients": (C.c_int, [C.c_double, _f64p, _f64p]),
    "true_peak_fir": (C.c_int, [C.c_int, _f32p]),
    "fallback_positions": (None, [_u32, _u8p]),
    "stereo_matrix": (None, [_u32, _u8p, _f32p]),
    "downmix_project": (C.c_int, [_f32p, _sz, _u32, _u8p, C.c_int, _f32p]),
    "spectrogram_default_config": (None, [C.POINTER(SpectrogramConfig)]),
    "spectrogram_create": (C.c_int, [C.POINTER(SpectrogramConfig), C.POINTER(_vp)]),
    "spectrogram_destroy": (None, [_vp]),
    "spectrogram_get_config": (C.c_int, [_vp, C.POINTER(SpectrogramConfig)]),
    "spectrogram_update_config": (C.c_int, [_vp, C.POINTER(SpectrogramConfig)]),
    "spectrogram_prepare": (C.c_int, [_vp]),
    "spectrogram_reset_audio": (C.c_int, [_vp]),
    "spectrogram_process_block": (C.c_int, [_vp, _f32p, _sz, _u32, C.c_float, _u8p, C.POINTER(SpectrogramUpdate)]),
    "spectrum_default_config": (None, [C.POINTER(SpectrumConfig)]),
    "spectrum_create": (C.c_int, [C.POINTER(SpectrumConfig), C.POINTER(_vp)]),
    "spectrum_destroy": (None, [_vp]),
    "spectrum_get_config": (C.c_int, [_vp, C.POINTER(SpectrumConfig)]),
    "spectrum_update_config": (C.c_int, [_vp, C.POINTER(SpectrumConfig)]),
    "spectrum_prepare": (C.c_int, [_vp]),
    "spectrum_reset_audio": (C.c_int, [_vp]),
    "spectrum_process_block": (C.c_int, [_vp, _f32p, _sz, _u32, C.c_float, _u8p, C.POINTER(SpectrumSnapshot)]),
    "loudness_default_config": (None, [C.POINTER(LoudnessConfig)]),
    "loudness_create": (C.c_int, [C.POINTER(LoudnessConfig), C.POINTER(_vp)]),
    "loudness_destroy": (None, [_vp]),
    "loudness_get_config": (C.c_int, [_vp, C.POINTER(LoudnessConfig)]),
    "loudness_reset_audio": (C.c_int, [_vp]),
    "loudness_process_block": (C.c_int, [_vp, _f32p, _sz, _u32, C.c_float, _u8p, C.POINTER(LoudnessSnapshot)]),
    "stft_frames_per_lane": (_u64, [C.POINTER(SpectrogramConfig), _u64]),
    "stft_plan_create": (C.c_int, [C.POINTER(SpectrogramConfig), C.c_int, C.POINTER(_vp)]),
    "stft_plan_destroy": (None, [_vp]),
    "stft_plan_bins": (_u32, [_vp]),
    "stft_plan_is_fast": (C.c_int, [_vp]),
    "stft_plan_power_scale": (C.c_float, [_vp]),
    "stft_execute_device": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _vp, _u64, _vp, _vp, _vp]),
    "stft_execute_host": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _vp, _u64, _vp, _vp]),
    "spectrum_hops_per_lane": (_u64, [C.POINTER(SpectrumConfig), _u64]),
    "spectrum_plan_create": (C.c_int, [C.POINTER(SpectrumConfig), C.POINTER(_vp)]),
    "spectrum_plan_destroy": (None, [_vp]),
    "spectrum_execute_device": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _vp, _vp, _vp, _vp]),
    "spectrum_execute_host": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _vp, _vp, _vp]),
    # row f3: peak label (peak_bin range / trace selection, interpolated_peak)
    "spectrum_default_peak_spec": (None, [C.POINTER(SpectrumPeakSpec)]),
    "spectrum_plan_set_peak_spec": (C.c_int, [_vp, C.POINTER(SpectrumPeakSpec)]),
    "spectrum_plan_get_peak_spec": (C.c_int, [_vp, C.POINTER(SpectrumPeakSpec)]),
    "spectrum_interpolate_peaks_device": (C.c_int, [_vp, _vp, _vp, _u64, _vp, _vp, _vp]),
    "spectrum_execute_host_peaks": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _vp, _vp, _vp, _vp, _vp]),
    "loudness_plan_create": (C.c_int, [C.POINTER(LoudnessConfig), _u32, _u8p, C.POINTER(_vp)]),
    "loudness_plan_destroy": (None, [_vp]),
    "loudness_execute_device": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _u64, _vp, _vp]),
    "loudness_execute_host": (C.c_int, [_vp, _vp, _u32, _u64, _u64, _u64, _vp]),
    # rows f1 / f4: the ordered audio timeline (host-side state machines)
    "timeline_create": (C.c_int, [C.POINTER(AudioFormat), C.POINTER(_vp)]),
    "timeline_destroy": (None, [_vp]),
    "timeline_accept": (C.c_int, [_vp, _f32p, _u64, C.POINTER(AudioFormat), _u64, _u64, SPAN_FN, _vp]),
    "timeline_flush": (C.c_int, [_vp, SPAN_FN, _vp]),
    "timeline_reset": (C.c_int, [_vp, _u64]),
    "timeline_cursor": (_u64, [_vp]),
    "timeline_pending_samples": (_sz, [_vp]),
    "meter_create": (C.c_int, [C.POINTER(_vp)]),
    "meter_destroy": (None, [_vp]),
    "meter_attach": (C.c_int, [_vp, _vp, _vp, _vp]),
    "meter_set_callback": (C.c_int, [_vp, INGEST_FN, _vp]),
    "meter_push": (C.c_int, [_vp, _f32p, _sz, C.POINTER(AudioFormat), _u32p]),
    "meter_push_silence": (C.c_int, [_vp, _u64, C.POINTER(AudioFormat), _u32p]),
    "meter_reset": (C.c_int, [_vp]),
    "meter_clear": (C.c_int, [_vp]),
    "meter_consume_span": (C.c_int, [_vp, C.c_int, _f32p, _sz, _u64, C.POINTER(AudioFormat), _u32p]),
    "meter_pending_samples": (_sz, [_vp]),
    "meter_has_format": (C.c_int, [_vp]),
    # row f1, device side: multi-stream ring
    "spectrogram_bank_create": (C.c_int, [C.POINTER(SpectrogramConfig), _u32, C.POINTER(_vp)]),
    "spectrogram_bank_destroy": (None, [_vp]),
    "spectrogram_bank_reset_audio": (C.c_int, [_vp]),
    "spectrogram_bank_push": (C.c_int, [_vp, _f32p, _u64, _sz, _u32, C.c_float, _u8p, C.POINTER(SpectrogramBankUpdate)]),
    "spectrogram_bank_pending": (_sz, [_vp]),
    "spectrum_bank_create": (C.c_int, [C.POINTER(SpectrumConfig), _u32, C.POINTER(_vp)]),
    "spectrum_bank_destroy": (None, [_vp]),
    "spectrum_bank_reset_audio": (C.c_int, [_vp]),
    "spectrum_bank_push": (C.c_int, [_vp, _f32p, _u64, _sz, _u32, C.c_float, _u8p, C.POINTER(SpectrumBankSnapshot)]),
    "spectrum_bank_pending": (_sz, [_vp]),
    "loudness_bank_create": (C.c_int, [C.POINTER(LoudnessConfig), _u32, C.POINTER(_vp)]),
    "loudness_bank_destroy": (None, [_vp]),
    "loudness_bank_reset_audio": (C.c_int, [_vp]),
    "loudness_bank_push": (C.c_int, [_vp, _f32p, _u64, _sz, _u32, C.c_float, _u8p, C.POINTER(LoudnessSnapshot)]),
    # row f2: splat accumulation + resolve
    "splat_image_size": (None, [C.POINTER(SplatParams), _u32p, _u32p]),
    "splat_accumulate_device": (C.c_int, [_vp, _u64, _vp, _u32, C.POINTER(SplatParams), _vp, _vp]),
    "splat_resolve_device": (C.c_int, [_vp, _u32, C.POINTER(SplatParams), _vp, _vp]),
    "splat_render_host": (C.c_int, [_vp, _u64, _vp, _u32, C.POINTER(SplatParams), _vp, _vp]),
    "stft_render_host": (C.c_int, [_vp, _vp, _u32, _u64, _u64, C.POINTER(SplatParams), _vp, _vp]),
}


def bind(lib: C.CDLL, prefix: str, symbols: dict | None = None, required: bool = True) -> SimpleNamespace:
    """Attach prototypes for `symbols` (default: every header symbol) found in `lib`."""
    ns = SimpleNamespace()
    ns._lib = lib
    ns._prefix = prefix
    for name, (restype, argtypes) in (symbols or HEADER_SYMBOLS).items():
        try:
            fn = getattr(lib, prefix + name)
        except AttributeError:
            if required:
                raise
            continue
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(ns, name, fn)
    return ns


def positions_array(positions) -> "C.Array | None":
    if positions is None:
        return None
    arr = (C.c_uint8 * MAX_CHANNELS)(*([POS_UNKNOWN] * MAX_CHANNELS))
    for i, p in enumerate(positions):
        arr[i] = int(p)
    return arr
