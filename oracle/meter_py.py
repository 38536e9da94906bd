"""Pure-Python restatement of the reference's ordered audio timeline (rows f1 / f4 of SURVEY.md §8).

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): imported by tests/, never by openmeters_b200.
Integer / control logic only, so parity with the product is exact (same chunk boundaries, same spans).

  RefDspBatcher  — meter.rs:27-84 `DspBatcher`, meter.rs:143-165 `ingest_silence`,
                   visuals/registry.rs:396-418 `VisualManager::ingest_samples` (generation change -> reset_audio)
  RefTimeline    — infra/pipewire/transport.rs:573-657 `AudioReader::{accept, switch, flush, reset_timeline}`
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Tuple

import numpy as np

SILENCE_CHUNK_FRAMES = 4096          # meter.rs:15
DSP_BATCH_FRAMES_AT_48K = 256        # meter.rs:16
MAX_DSP_INGEST_FRAMES_AT_48K = 1024  # meter.rs:17
MAX_SILENCE_SECONDS = 2              # meter.rs:18
MAX_CAPTURE_CHANNELS = 8             # infra/pipewire.rs:34
DEFAULT_SAMPLE_RATE = 48000.0


def _round_half_away(x: float) -> float:  # f64::round
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


@dataclass(frozen=True)
class RefFormat:  # dsp.rs:79-85
    channels: int
    sample_rate: float  # an f32 value
    generation: int
    positions: Tuple[int, ...] = ()

    def rate(self) -> int:  # dsp.rs:103-105 (f32::round)
        return max(int(_round_half_away(float(np.float32(self.sample_rate)))), 1)


def scaled_samples(frames_at_48k: int, f: RefFormat) -> int:  # meter.rs:20-25
    frames = max(_round_half_away(frames_at_48k * float(np.float32(f.sample_rate)) / DEFAULT_SAMPLE_RATE), 1.0)
    return int(frames) * max(f.channels, 1)


class RefDspBatcher:
    """`ingest(samples, format)` is called per ingest_samples; `reset_audio()` when the manager resets."""

    def __init__(self, ingest: Callable[[np.ndarray, RefFormat], None], reset_audio: Callable[[], None] = lambda: None):
        self.samples: List[float] = []
        self.format: Optional[RefFormat] = None
        self._ingest_cb = ingest
        self._reset_cb = reset_audio
        self._generation: Optional[int] = None  # VisualManager.format_generation

    # visuals/registry.rs:396-418
    def _ingest(self, chunk: np.ndarray, f: RefFormat) -> None:
        if chunk.size == 0:
            return
        if self._generation is not None and self._generation != f.generation:
            self._reset_cb()
        self._generation = f.generation
        self._ingest_cb(np.asarray(chunk, np.float32), f)

    # meter.rs:40-73
    def push(self, samples, f: RefFormat) -> int:
        samples = np.asarray(samples, np.float32).reshape(-1)
        if self.format is not None and self.format != f:
            self.samples = []
        self.format = f
        batch = scaled_samples(DSP_BATCH_FRAMES_AT_48K, f)
        count = 0
        if self.samples:
            take = min(batch - len(self.samples), samples.size)
            self.samples.extend(samples[:take].tolist())
            samples = samples[take:]
            if len(self.samples) == batch:
                self._ingest(np.array(self.samples, np.float32), f)
                self.samples = []
                count += 1
        ready = samples.size // batch * batch
        step = scaled_samples(MAX_DSP_INGEST_FRAMES_AT_48K, f)
        for o in range(0, ready, step):
            self._ingest(samples[o:min(o + step, ready)], f)
            count += 1
        self.samples.extend(samples[ready:].tolist())
        return count

    def clear(self) -> None:  # meter.rs:80-83
        self.samples = []
        self.format = None

    def reset(self) -> None:  # meter.rs:75-78
        self.clear()
        self._reset_cb()

    # meter.rs:143-165
    def ingest_silence(self, frames: int, f: RefFormat) -> int:
        limit = int(max(_round_half_away(MAX_SILENCE_SECONDS * float(np.float32(f.sample_rate))), 1.0))
        if frames > limit:
            self.reset()
            return 0
        capacity = SILENCE_CHUNK_FRAMES * MAX_CAPTURE_CHANNELS // max(f.channels, 1)
        remaining, count = frames, 0
        while remaining > 0:
            chunk = min(remaining, capacity)
            count += self.push(np.zeros(chunk * f.channels, np.float32), f)
            remaining -= chunk
        return count


def ns_frames(ns: int, rate: int) -> int:       # transport.rs:109-111
    return min(ns * rate // 1_000_000_000, 2**64 - 1)


def ns_frames_ceil(ns: int, rate: int) -> int:  # transport.rs:113-117
    return min(-(-ns * rate // 1_000_000_000), 2**64 - 1)


def frames_ns(frames: int, rate: int) -> int:   # transport.rs:105-107
    return min(frames * 1_000_000_000 // max(rate, 1), 2**64 - 1)


class RefTimeline:
    """Spans are returned as tuples: ("pcm", np.ndarray, format) / ("silence", frames, format)."""

    def __init__(self, fmt: RefFormat):
        self.scratch: List[float] = []
        self.format = fmt
        self.cursor = 0
        self.align_next_packet = True

    def flush(self) -> list:  # transport.rs:634-643
        if not self.scratch:
            return []
        out = [("pcm", np.array(self.scratch, np.float32), self.format)]
        self.scratch = []
        return out

    def reset_timeline(self, cursor: int) -> None:  # transport.rs:645-656
        self.scratch = []
        self.cursor = cursor
        self.align_next_packet = True

    # transport.rs:573-632
    def accept(self, samples, frames: int, f: RefFormat, start: int, end: int) -> list:
        out = []
        if self.format != f:  # ::switch
            out += self.flush()
            self.format = f
        if self.align_next_packet:
            self.align_next_packet = False
            self.cursor = start
        rate = f.rate()
        gap = ns_frames(start - self.cursor, rate) if start > self.cursor else None
        skip = min(ns_frames_ceil(min(self.cursor, end) - start, rate), frames) if self.cursor > start else 0
        self.cursor = max(self.cursor, end)
        if gap is not None and gap > 0:
            out += self.flush()
            out.append(("silence", gap, f))
        if samples is not None:
            if skip < frames:
                s = np.asarray(samples, np.float32).reshape(-1)
                self.scratch.extend(s[skip * f.channels: frames * f.channels].tolist())
        elif skip < frames:
            out += self.flush()
            out.append(("silence", frames - skip, f))
        return out
