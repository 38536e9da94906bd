"""Host-side mirror of the reference's ordered audio timeline over the C ABI (rows f1 / f4 of SURVEY.md §8).

``PacketTimeline``  — ``AudioReader::accept / flush / reset_timeline`` (infra/pipewire/transport.rs:573-657):
                     packets with capture-clock extents in, ``CapturedSpan``s out.
``Meter``           — ``DspBatcher`` + ``ingest_silence`` (meter.rs:27-84,143-165) feeding
                     ``VisualManager::ingest_samples`` (visuals/registry.rs:396-418) for the three hot-path
                     processors; every ingest is reported as an ``Ingest`` record.

All logic lives behind the C ABI (csrc/meter.cu); this file only marshals.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _capi as capi
from .processors import (LoudnessProcessor, LoudnessSnapshot, SpectrogramProcessor, SpectrogramUpdate, SpectrumProcessor,
                         SpectrumSnapshot, _check, _default_api, _f32, _ptr)


@dataclass(frozen=True)
class AudioFormat:
    """dsp.rs:79-101 (``AudioFormat::new`` clamps channels to 1..=8 and the rate to >= 1)."""

    channels: int = 2
    sample_rate: float = 48000.0
    generation: int = 1
    positions: Optional[Tuple[int, ...]] = None  # None => ChannelPosition::fallback(channels)

    def to_c(self, api) -> capi.AudioFormat:
        f = capi.AudioFormat()
        f.channels = min(max(int(self.channels), 1), capi.MAX_CHANNELS)
        f.sample_rate = float(self.sample_rate)
        f.generation = int(self.generation)
        if self.positions is None:
            arr = (C.c_uint8 * capi.MAX_CHANNELS)()
            api.fallback_positions(f.channels, arr)
        else:
            arr = capi.positions_array(self.positions)
        for i in range(capi.MAX_CHANNELS):
            f.positions[i] = arr[i]
        return f

    def rate(self) -> int:  # dsp.rs:103-105
        return max(int(np.round(np.float32(self.sample_rate))), 1)


@dataclass
class Span:
    """transport.rs:39-54 CapturedSpan. kind: capi.SPAN_*; samples for PCM, frames for silence."""

    kind: int
    samples: Optional[np.ndarray] = None
    frames: int = 0
    generation: int = 0


class PacketTimeline:
    def __init__(self, fmt: AudioFormat, api=None):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        c = fmt.to_c(self._api)
        _check(self._api, self._api.timeline_create(C.byref(c), C.byref(self._h)), "timeline_create")
        self._out: List[Span] = []

        def on_span(_user, kind, samples, n, frames, f):
            if kind == capi.SPAN_PCM:
                self._out.append(Span(kind, np.ctypeslib.as_array(samples, shape=(n,)).copy(), n // max(f.contents.channels, 1),
                                      f.contents.generation))
            else:
                self._out.append(Span(kind, None, int(frames), f.contents.generation if f else 0))

        self._cb = capi.SPAN_FN(on_span)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.timeline_destroy(h)

    def _take(self) -> List[Span]:
        out, self._out = self._out, []
        return out

    def accept(self, samples: Optional[np.ndarray], frames: int, fmt: AudioFormat, start_ns: int, end_ns: int) -> List[Span]:
        c = fmt.to_c(self._api)
        buf = None if samples is None else _f32(samples).reshape(-1)
        _check(self._api, self._api.timeline_accept(self._h, None if buf is None else _ptr(buf), frames, C.byref(c),
                                                    start_ns, end_ns, self._cb, None), "timeline_accept")
        return self._take()

    def flush(self) -> List[Span]:
        _check(self._api, self._api.timeline_flush(self._h, self._cb, None), "timeline_flush")
        return self._take()

    def reset_timeline(self, cursor_ns: int) -> None:
        _check(self._api, self._api.timeline_reset(self._h, cursor_ns), "timeline_reset")

    @property
    def cursor(self) -> int:
        return int(self._api.timeline_cursor(self._h))

    @property
    def pending_samples(self) -> int:
        return int(self._api.timeline_pending_samples(self._h))


@dataclass
class Ingest:
    """One ``VisualManager::ingest_samples`` call: the chunk and what each attached processor returned."""

    n_samples: int
    generation: int
    samples: Optional[np.ndarray] = None
    spectrogram: Optional[SpectrogramUpdate] = None
    spectrum: Optional[SpectrumSnapshot] = None
    loudness: Optional[LoudnessSnapshot] = None


class Meter:
    """``DspBatcher`` bound to (optional) spectrogram / spectrum / loudness processors."""

    def __init__(self, spectrogram: SpectrogramProcessor | None = None, spectrum: SpectrumProcessor | None = None,
                 loudness: LoudnessProcessor | None = None, api=None, keep_samples: bool = False):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        _check(self._api, self._api.meter_create(C.byref(self._h)), "meter_create")
        self._procs = (spectrogram, spectrum, loudness)  # keep the borrowed handles alive
        _check(self._api, self._api.meter_attach(self._h, spectrogram._h if spectrogram else None,
                                                 spectrum._h if spectrum else None, loudness._h if loudness else None), "meter_attach")
        self._out: List[Ingest] = []

        def on_ingest(_user, samples, n, f, up, sn, ls):
            rec = Ingest(int(n), int(f.contents.generation))
            if keep_samples:
                rec.samples = np.ctypeslib.as_array(samples, shape=(n,)).copy()
            if up:
                rec.spectrogram = SpectrogramProcessor._update(up.contents)
            if sn:
                rec.spectrum = SpectrumProcessor._snapshot(sn.contents)
            if ls:
                rec.loudness = LoudnessSnapshot.from_c(ls.contents)
            self._out.append(rec)

        self._cb = capi.INGEST_FN(on_ingest)
        _check(self._api, self._api.meter_set_callback(self._h, self._cb, None), "meter_set_callback")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.meter_destroy(h)

    def _take(self) -> List[Ingest]:
        out, self._out = self._out, []
        return out

    def push(self, samples, fmt: AudioFormat) -> List[Ingest]:
        x = _f32(samples).reshape(-1)
        c = fmt.to_c(self._api)
        n = C.c_uint32(0)
        _check(self._api, self._api.meter_push(self._h, _ptr(x), x.size, C.byref(c), C.byref(n)), "meter_push")
        out = self._take()
        assert len(out) == n.value
        return out

    def push_silence(self, frames: int, fmt: AudioFormat) -> List[Ingest]:
        c = fmt.to_c(self._api)
        n = C.c_uint32(0)
        _check(self._api, self._api.meter_push_silence(self._h, frames, C.byref(c), C.byref(n)), "meter_push_silence")
        return self._take()

    def consume(self, span: Span, fmt: AudioFormat) -> List[Ingest]:
        c = fmt.to_c(self._api)
        n = C.c_uint32(0)
        x = None if span.samples is None else _f32(span.samples).reshape(-1)
        _check(self._api, self._api.meter_consume_span(self._h, span.kind, None if x is None else _ptr(x), 0 if x is None else x.size,
                                                       span.frames, C.byref(c), C.byref(n)), "meter_consume_span")
        return self._take()

    def reset(self) -> None:
        _check(self._api, self._api.meter_reset(self._h), "meter_reset")

    def clear(self) -> None:
        _check(self._api, self._api.meter_clear(self._h), "meter_clear")

    @property
    def pending_samples(self) -> int:
        return int(self._api.meter_pending_samples(self._h))

    @property
    def has_format(self) -> bool:
        return bool(self._api.meter_has_format(self._h))


@dataclass
class BankUpdate:
    """S SpectrogramUpdates in dense form: columns[s] is the list of stream s's new columns ((n,3) float32 point arrays or
    (bins,) uint16 code arrays), exactly what S separate SpectrogramProcessors would have returned."""

    fft_size: int
    hop_size: int
    sample_rate: float
    history_length: int
    reset: bool
    reassigned_power_scale: float
    kind: int
    columns: list


class SpectrogramBank:
    """Device-side multi-stream ring (row f1): S lock-step spectrogram streams, one kernel launch per push."""

    def __init__(self, config, n_streams: int, api=None):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        self.n_streams = int(n_streams)
        c = config.to_c()
        _check(self._api, self._api.spectrogram_bank_create(C.byref(c), self.n_streams, C.byref(self._h)), "spectrogram_bank_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.spectrogram_bank_destroy(h)

    def reset_audio(self) -> None:
        _check(self._api, self._api.spectrogram_bank_reset_audio(self._h), "spectrogram_bank_reset_audio")

    @property
    def pending(self) -> int:
        return int(self._api.spectrogram_bank_pending(self._h))

    def push(self, blocks, channels: int = 1, sample_rate: float = 48000.0, positions=None, copy: bool = True) -> Optional[BankUpdate]:
        """blocks: (S, frames * channels) float32, one interleaved block per stream."""
        x = np.ascontiguousarray(blocks, np.float32)
        assert x.ndim == 2 and x.shape[0] == self.n_streams
        frames = x.shape[1] // max(channels, 1)
        up = capi.SpectrogramBankUpdate()
        rc = _check(self._api, self._api.spectrogram_bank_push(self._h, _ptr(x), x.shape[1], frames, channels, sample_rate,
                                                               capi.positions_array(positions), C.byref(up)), "spectrogram_bank_push")
        if rc == capi.NO_DATA:
            return None
        S, n, bins = up.n_streams, up.n_columns, up.bins
        cols = []
        if up.kind == capi.COLUMN_REASSIGNED:
            cnt = np.ctypeslib.as_array(up.counts, shape=(S, n))
            pts = np.ctypeslib.as_array(C.cast(up.points, C.POINTER(C.c_float)), shape=(S, n, bins, 3))
            if copy:
                cols = [[pts[s, c, :cnt[s, c]].copy() for c in range(n)] for s in range(S)]
            else:
                cols = (cnt, pts)
        else:
            codes = np.ctypeslib.as_array(up.classic_db, shape=(S, n, bins))
            cols = [[codes[s, c].copy() for c in range(n)] for s in range(S)] if copy else codes
        return BankUpdate(up.fft_size, up.hop_size, up.sample_rate, up.history_length, bool(up.reset), up.reassigned_power_scale, up.kind, cols)


class LoudnessBank:
    """Device-side multi-stream loudness (row f1): S lock-step LoudnessProcessors, one kernel launch per push."""

    def __init__(self, config=None, n_streams: int = 1, api=None):
        from .processors import LoudnessConfig

        self._api = api or _default_api()
        self._h = C.c_void_p()
        self.n_streams = int(n_streams)
        cfg = config or LoudnessConfig()
        c = capi.LoudnessConfig(cfg.sample_rate, cfg.floor_db)
        _check(self._api, self._api.loudness_bank_create(C.byref(c), self.n_streams, C.byref(self._h)), "loudness_bank_create")
        self._snaps = (capi.LoudnessSnapshot * self.n_streams)()

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.loudness_bank_destroy(h)

    def reset_audio(self) -> None:
        _check(self._api, self._api.loudness_bank_reset_audio(self._h), "loudness_bank_reset_audio")

    def push(self, blocks, channels: int = 2, sample_rate: float = 48000.0, positions=None):
        """blocks: (S, frames * channels) float32, one interleaved block per stream -> ctypes array of S LoudnessSnapshots
        (valid until the next push), or None when the block is shorter than one frame."""
        x = np.ascontiguousarray(blocks, np.float32)
        assert x.ndim == 2 and x.shape[0] == self.n_streams
        rc = _check(self._api, self._api.loudness_bank_push(self._h, _ptr(x), x.shape[1], x.shape[1], channels, sample_rate,
                                                            capi.positions_array(positions), self._snaps), "loudness_bank_push")
        return None if rc == capi.NO_DATA else self._snaps


class SpectrumBank:
    """Device-side multi-stream spectrum analyzer (row f1): S lock-step SpectrumProcessors, one launch chain per push."""

    def __init__(self, config, n_streams: int, api=None):
        self._api = api or _default_api()
        self._h = C.c_void_p()
        self.n_streams = int(n_streams)
        c = config.to_c()
        _check(self._api, self._api.spectrum_bank_create(C.byref(c), self.n_streams, C.byref(self._h)), "spectrum_bank_create")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._api.spectrum_bank_destroy(h)

    def reset_audio(self) -> None:
        _check(self._api, self._api.spectrum_bank_reset_audio(self._h), "spectrum_bank_reset_audio")

    @property
    def pending(self) -> int:
        return int(self._api.spectrum_bank_pending(self._h))

    def push(self, blocks, channels: int = 2, sample_rate: float = 48000.0, positions=None, copy: bool = True):
        """blocks: (S, frames * channels) float32 -> None, or (trace_index, frequency_bins, weighted[S, T, bins], raw[S, T, bins])."""
        x = np.ascontiguousarray(blocks, np.float32)
        assert x.ndim == 2 and x.shape[0] == self.n_streams
        frames = x.shape[1] // max(channels, 1)
        snap = capi.SpectrumBankSnapshot()
        rc = _check(self._api, self._api.spectrum_bank_push(self._h, _ptr(x), x.shape[1], frames, channels, sample_rate,
                                                            capi.positions_array(positions), C.byref(snap)), "spectrum_bank_push")
        if rc == capi.NO_DATA:
            return None
        S, T, bins = snap.n_streams, snap.n_traces, snap.bins
        w = np.ctypeslib.as_array(snap.weighted, shape=(S, T, bins))
        r = np.ctypeslib.as_array(snap.raw, shape=(S, T, bins))
        f = np.ctypeslib.as_array(snap.frequency_bins, shape=(bins,))
        if copy:
            w, r, f = w.copy(), r.copy(), f.copy()
        return tuple(snap.trace_index[:T]), f, w, r
