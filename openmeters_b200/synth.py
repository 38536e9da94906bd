"""Synthetic PCM for the BASELINE configs: generated in f64, cast once to f32.

Every implementation (oracle, CUDA, bench) reads the same bytes produced here — a sine is never
re-evaluated per implementation (1-ulp sinf differences would pollute parity).
"""
from __future__ import annotations

import numpy as np


def chirp(n: int, sr: float, f0: float, f1: float, amp: float) -> np.ndarray:
    t = np.arange(n, dtype=np.float64) / sr
    dur = n / sr
    phase = 2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / dur * t * t)
    return amp * np.sin(phase)


def cfg1_stereo(seconds: float = 10.0, sr: float = 48000.0) -> np.ndarray:
    """2 ch: L = 0.5*chirp 20 Hz->20 kHz, R = 0.25 * same delayed 480 samples. Interleaved f32."""
    n = int(seconds * sr)
    left = chirp(n, sr, 20.0, 20000.0, 0.5)
    right = np.zeros(n)
    right[480:] = 0.5 * left[:-480]
    return np.stack([left, right], 1).astype(np.float32).reshape(-1)


def lane_signal(n: int, sr: float, f1: float, seed: int) -> np.ndarray:
    """0.4*chirp(20 Hz -> f1) + 0.05*uniform[-1,1) noise from default_rng(seed); n samples, f32."""
    rng = np.random.default_rng(seed)
    return (chirp(n, sr, 20.0, f1, 0.4) + 0.05 * rng.uniform(-1.0, 1.0, n)).astype(np.float32)


def cfg2_lanes(n_lanes: int = 8, seconds: float = 60.0, sr: float = 48000.0, seed0: int = 1000) -> np.ndarray:
    """lane c: 0.4*chirp(20 Hz -> (c%8+1)*2.5 kHz) + 0.05*uniform[-1,1) (default_rng(seed0+c)). (L, S) f32."""
    n = int(round(seconds * sr))
    out = np.empty((n_lanes, n), np.float32)
    for c in range(n_lanes):
        out[c] = lane_signal(n, sr, (c % 8 + 1) * 2500.0, seed0 + c)
    return out


def cfg3_surround(seconds: float = 30.0, sr: float = 48000.0) -> np.ndarray:
    """8 ch SURROUND: ch c = 0.3*sine(997*(c+1)/4 Hz) + 0.05 noise(seed 2000+c); first 0.5 s of ch 5 exactly
    zero (lazy activation); one 17 kHz 0.9 burst on ch 1 (inter-sample peaks). Interleaved f32."""
    n = int(seconds * sr)
    t = np.arange(n, dtype=np.float64) / sr
    ch = np.empty((n, 8))
    for c in range(8):
        rng = np.random.default_rng(2000 + c)
        ch[:, c] = 0.3 * np.sin(2 * np.pi * 997.0 * (c + 1) / 4.0 * t) + 0.05 * rng.uniform(-1.0, 1.0, n)
    ch[: int(0.5 * sr), 5] = 0.0
    b0 = min(int(2.0 * sr), max(n - 480, 0))
    ch[b0:b0 + 480, 1] = 0.9 * np.sin(2 * np.pi * 17000.0 * t[: min(480, n - b0)])
    return ch.astype(np.float32).reshape(-1)


def cfg4_streams(n_streams: int = 64, seconds: float = 20.0, sr: float = 48000.0, first: int = 0) -> np.ndarray:
    """stream s: two sines at 100(s+1) Hz and 55(s+3) Hz, amplitude-modulated at 0.5 Hz, + noise(seed 3000+s).
    Returns (n_streams, 2, S) planar f32: lane 0 = Left, lane 1 = Right (Right = 0.7*Left + own noise)."""
    n = int(seconds * sr)
    t = np.arange(n, dtype=np.float64) / sr
    out = np.empty((n_streams, 2, n), np.float32)
    am = 0.5 * (1 + np.sin(2 * np.pi * 0.5 * t))
    for i in range(n_streams):
        s = first + i  # global stream index: streams [first, first + n_streams) of the set
        rng = np.random.default_rng(3000 + s)
        base = am * (0.3 * np.sin(2 * np.pi * 100.0 * (s + 1) * t) + 0.2 * np.sin(2 * np.pi * 55.0 * (s + 3) * t))
        out[i, 0] = (base + 0.01 * rng.uniform(-1, 1, n)).astype(np.float32)
        out[i, 1] = (0.7 * base + 0.01 * rng.uniform(-1, 1, n)).astype(np.float32)
    return out


def cfg5_lanes(n_lanes: int, samples: int, sr: float = 96000.0, seed0: int = 5000) -> np.ndarray:
    out = np.empty((n_lanes, samples), np.float32)
    for c in range(n_lanes):
        rng = np.random.default_rng(seed0 + c)
        out[c] = (chirp(samples, sr, 20.0, (c % 8 + 1) * 2500.0, 0.4) + 0.05 * rng.uniform(-1.0, 1.0, samples)).astype(np.float32)
    return out
