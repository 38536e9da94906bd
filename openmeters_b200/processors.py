"""Host-side mirror of the reference's processor surface over the C ABI.

Same names, argument meaning and error behaviour as the reference's
``SpectrogramProcessor`` / ``SpectrumProcessor`` / ``LoudnessProcessor``
(src/visuals/{spectrogram,spectrum,loudness}/processor.rs) and ``AudioBlock``
(src/dsp.rs:108-262): ``new(config)`` -> constructor, ``config()``,
``update_config()``, ``prepare()``, ``process_block(block) -> snapshot | None``,
``reset_audio()``.  ``None`` means "nothing new", configs are normalised rather
than rejected; a failing CUDA call raises ``OmbError`` (the reference would
``panic=abort``).

All arithmetic happens behind the C ABI (CUDA kernels); this file only marshals.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _capi as capi


class OmbError(RuntimeError):
    pass


def _default_api():
    from ._lib import api

    return api()


def _check(api, rc: int, what: str) -> int:
    if rc < 0:
        msg = api.last_error() if hasattr(api, "last_error") else b""
        raise OmbError(f"{what} failed with status {rc}: {(msg or b'').decode(errors='replace')}")
    return rc


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray, ty=C.c_float):
    return a.ctypes.data_as(C.POINTER(ty))


@dataclass
class AudioBlock:
    """dsp.rs:108-115 — interleaved f32 samples + format."""

    samples: np.ndarray
    channels: int = 1
    sample_rate: float = 48000.0
    positions: Optional[Sequence[int]] = None  # None => ChannelPosition::fallback(channels)

    def __post_init__(self):
        self.samples = _f32(self.samples).reshape(-1)
        self.channels = min(max(int(self.channels), 1), capi.MAX_CHANNELS)

    def frame_count(self) -> int:
        return self.samples.size // max(self.channels, 1)

    def is_empty(self) -> bool:
        return self.samples.size < max(self.channels, 1)


# --------------------------------------------------------------------------- spectrogram
@dataclass
class SpectrogramConfig:
    """spectrogram/processor.rs:45-56 (same defaults)."""

    sample_rate: float = 48000.0
    fft_size: int = 2048
    hop_size: int = 64
    window: int = capi.WINDOW_HANN
    history_length: int = 0
    use_reassignment: bool = True
    zero_padding_factor: int = 1

    def to_c(self) -> capi.SpectrogramConfig:
        c = capi.SpectrogramConfig()
        c.sample_rate = self.sample_rate
        c.window = self.window
        c.fft_size = self.fft_size
        c.hop_size = self.hop_size
        c.history_length = self.history_length
        c.zero_padding_factor = self.zero_padding_factor
        c.use_reassignment = 1 if self.use_reassignment else 0
        return c

    @staticmethod
    def from_c(c: capi.SpectrogramConfig) -> "SpectrogramConfig":
        return SpectrogramConfig(c.sample_rate, c.fft_size, c.hop_size, c.window, c.history_length,
                                 bool(c.use_reassignment), c.zero_padding_factor)


@dataclass
class SpectrogramUpdate:
    """spectrogram/processor.rs:160-168; new_columns is a list of
    (n,3) float32 arrays [time_offset, freq_hz, power] (reassigned) or (bins,) uint16 arrays (classic)."""

    fft_size: int
    hop_size: int
    sample_rate: float
    history_length: int
    reset: bool
    reassigned_power_scale: float
    kind: int
    new_columns: list = field(default_factory=list)


class SpectrogramProcessor:
    def __init__(self, config: SpectrogramConfig | None = None, api=None):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        cfg = (config or SpectrogramConfig()).to_c()
        _check(self._api, self._api.spectrogram_create(C.byref(cfg), C.byref(self._h)), "spectrogram_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.spectrogram_destroy(h)

    def config(self) -> SpectrogramConfig:
        c = capi.SpectrogramConfig()
        _check(self._api, self._api.spectrogram_get_config(self._h, C.byref(c)), "spectrogram_get_config")
        return SpectrogramConfig.from_c(c)

    def update_config(self, config: SpectrogramConfig) -> None:
        c = config.to_c()
        _check(self._api, self._api.spectrogram_update_config(self._h, C.byref(c)), "spectrogram_update_config")

    def prepare(self) -> None:
        _check(self._api, self._api.spectrogram_prepare(self._h), "spectrogram_prepare")

    def reset_audio(self) -> None:
        _check(self._api, self._api.spectrogram_reset_audio(self._h), "spectrogram_reset_audio")

    def process_block(self, block: AudioBlock) -> Optional[SpectrogramUpdate]:
        up = capi.SpectrogramUpdate()
        rc = _check(self._api, self._api.spectrogram_process_block(
            self._h, _ptr(block.samples), block.samples.size, block.channels, block.sample_rate,
            capi.positions_array(block.positions), C.byref(up)), "spectrogram_process_block")
        if rc == capi.NO_DATA:
            return None
        return self._update(up)

    @staticmethod
    def _update(up: capi.SpectrogramUpdate) -> SpectrogramUpdate:
        """Copies a library-owned omb_spectrogram_update into numpy-backed columns."""
        n = up.n_columns
        offs = np.ctypeslib.as_array(up.column_offsets, shape=(n + 1,)).astype(np.int64)
        cols = []
        if up.kind == capi.COLUMN_REASSIGNED:
            total = int(offs[-1])
            pts = (np.ctypeslib.as_array(C.cast(up.points, C.POINTER(C.c_float)), shape=(total, 3)).copy()
                   if total else np.zeros((0, 3), np.float32))
            for c in range(n):
                cols.append(pts[offs[c]:offs[c + 1]])
        else:
            total = n * up.bins
            codes = np.ctypeslib.as_array(up.classic_db, shape=(total,)).copy()
            for c in range(n):
                cols.append(codes[c * up.bins:(c + 1) * up.bins])
        return SpectrogramUpdate(up.fft_size, up.hop_size, up.sample_rate, up.history_length, bool(up.reset),
                                 up.reassigned_power_scale, up.kind, cols)


# --------------------------------------------------------------------------- spectrum
@dataclass
class SpectrumConfig:
    """spectrum/processor.rs:39-51 (same defaults). averaging: (mode, param)."""

    sample_rate: float = 48000.0
    fft_size: int = 16384
    hop_size: int = 1024
    window: int = capi.WINDOW_HANN
    averaging: int = capi.AVG_NONE
    averaging_param: float = 0.0
    source: int = capi.CHANNEL_MID
    secondary_source: int = capi.CHANNEL_NONE
    floor_db: float = -100.0

    def to_c(self) -> capi.SpectrumConfig:
        c = capi.SpectrumConfig()
        c.sample_rate = self.sample_rate
        c.window = self.window
        c.fft_size = self.fft_size
        c.hop_size = self.hop_size
        c.averaging = self.averaging
        c.averaging_param = self.averaging_param
        c.source = self.source
        c.secondary_source = self.secondary_source
        c.floor_db = self.floor_db
        return c

    @staticmethod
    def from_c(c: capi.SpectrumConfig) -> "SpectrumConfig":
        return SpectrumConfig(c.sample_rate, c.fft_size, c.hop_size, c.window, c.averaging, c.averaging_param,
                              c.source, c.secondary_source, c.floor_db)


@dataclass
class SpectrumSnapshot:
    """spectrum/processor.rs:33-37; traces[trace][0]=weighted dB, [trace][1]=raw dB."""

    frequency_bins: np.ndarray
    traces: list


class SpectrumProcessor:
    def __init__(self, config: SpectrumConfig | None = None, api=None):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        cfg = (config or SpectrumConfig()).to_c()
        _check(self._api, self._api.spectrum_create(C.byref(cfg), C.byref(self._h)), "spectrum_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.spectrum_destroy(h)

    def config(self) -> SpectrumConfig:
        c = capi.SpectrumConfig()
        _check(self._api, self._api.spectrum_get_config(self._h, C.byref(c)), "spectrum_get_config")
        return SpectrumConfig.from_c(c)

    def update_config(self, config: SpectrumConfig) -> None:
        c = config.to_c()
        _check(self._api, self._api.spectrum_update_config(self._h, C.byref(c)), "spectrum_update_config")

    def prepare(self) -> None:
        _check(self._api, self._api.spectrum_prepare(self._h), "spectrum_prepare")

    def reset_audio(self) -> None:
        _check(self._api, self._api.spectrum_reset_audio(self._h), "spectrum_reset_audio")

    @staticmethod
    def _snapshot(snap: capi.SpectrumSnapshot) -> SpectrumSnapshot:
        b = snap.bins
        freq = np.ctypeslib.as_array(snap.frequency_bins, shape=(b,)).copy() if b else np.zeros(0, np.float32)
        traces = [[np.ctypeslib.as_array(snap.traces[t][w], shape=(b,)).copy() if b else np.zeros(0, np.float32)
                   for w in range(2)] for t in range(2)]
        return SpectrumSnapshot(freq, traces)

    def process_block(self, block: AudioBlock) -> Optional[SpectrumSnapshot]:
        snap = capi.SpectrumSnapshot()
        rc = _check(self._api, self._api.spectrum_process_block(
            self._h, _ptr(block.samples), block.samples.size, block.channels, block.sample_rate,
            capi.positions_array(block.positions), C.byref(snap)), "spectrum_process_block")
        if rc == capi.NO_DATA:
            return None
        return self._snapshot(snap)


# --------------------------------------------------------------------------- loudness
@dataclass
class LoudnessConfig:
    sample_rate: float = 48000.0
    floor_db: float = -99.9


@dataclass
class LoudnessSnapshot:
    """loudness/processor.rs:185-194."""

    short_term_loudness: float
    momentary_loudness: float
    rms_fast_db: np.ndarray
    rms_slow_db: np.ndarray
    true_peak_db: np.ndarray
    channel_count: int
    positions: tuple

    @staticmethod
    def from_c(s: capi.LoudnessSnapshot) -> "LoudnessSnapshot":
        return LoudnessSnapshot(
            float(s.short_term_loudness), float(s.momentary_loudness),
            np.array(s.rms_fast_db[:], np.float32), np.array(s.rms_slow_db[:], np.float32),
            np.array(s.true_peak_db[:], np.float32), int(s.channel_count), tuple(s.positions[:]))

    def __eq__(self, o):  # PartialEq on the reference struct
        return (np.float32(self.short_term_loudness) == np.float32(o.short_term_loudness)
                and np.float32(self.momentary_loudness) == np.float32(o.momentary_loudness)
                and np.array_equal(self.rms_fast_db, o.rms_fast_db) and np.array_equal(self.rms_slow_db, o.rms_slow_db)
                and np.array_equal(self.true_peak_db, o.true_peak_db) and self.channel_count == o.channel_count
                and self.positions == o.positions)


class LoudnessProcessor:
    def __init__(self, config: LoudnessConfig | None = None, api=None):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        cfg = config or LoudnessConfig()
        c = capi.LoudnessConfig(cfg.sample_rate, cfg.floor_db)
        _check(self._api, self._api.loudness_create(C.byref(c), C.byref(self._h)), "loudness_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.loudness_destroy(h)

    def config(self) -> LoudnessConfig:
        c = capi.LoudnessConfig()
        _check(self._api, self._api.loudness_get_config(self._h, C.byref(c)), "loudness_get_config")
        return LoudnessConfig(c.sample_rate, c.floor_db)

    def reset_audio(self) -> None:
        _check(self._api, self._api.loudness_reset_audio(self._h), "loudness_reset_audio")

    def process_block(self, block: AudioBlock) -> Optional[LoudnessSnapshot]:
        snap = capi.LoudnessSnapshot()
        rc = _check(self._api, self._api.loudness_process_block(
            self._h, _ptr(block.samples), block.samples.size, block.channels, block.sample_rate,
            capi.positions_array(block.positions), C.byref(snap)), "loudness_process_block")
        if rc == capi.NO_DATA:
            return None
        return LoudnessSnapshot.from_c(snap)
