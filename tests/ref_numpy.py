"""Independent float64 numpy restatement of one reassigned column — a second opinion on the C++
oracle and the yard-stick that calibrates the f32 noise model used by tests/parity.py."""
from __future__ import annotations

import numpy as np

COEFFS = {0: [1.0], 1: [0.5, -0.5], 2: [25 / 46, -21 / 46], 3: [0.42, -0.5, 0.08], 4: [0.35875, -0.48829, 0.14128, -0.01168]}


def window(kind: int, n: int) -> np.ndarray:
    ph = 2 * np.pi * np.arange(n) / n
    return sum(c * np.cos(ph * k) for k, c in enumerate(COEFFS[kind]))


def reassigned_column(frame: np.ndarray, kind: int, n: int, hop: int, sr: float, zp: int = 1, win: np.ndarray | None = None):
    """frame: H = 2n samples. Returns dict of per-bin arrays (no thresholds applied): power, freq, time."""
    H = 2 * n
    F = n * zp
    h = window(kind, n) if win is None else win.astype(np.float64)
    A = np.fft.fft(frame.astype(np.float64))
    A[0] = 0
    A[H // 2 + 1:] = 0
    a = np.fft.ifft(A) * H
    off = (H - n) // 2
    c = a[off:off + n]
    k = np.arange(n)
    omega = 2 * np.pi / n * np.where(k > n // 2, k - n, k)
    Wf = np.fft.fft(h)
    Wf[0] = 0
    Wf[n // 2] = 0
    dh = np.real(np.fft.ifft(1j * omega * Wf))
    th = (np.arange(n) - (n - 1) / 2) * h
    S = np.fft.fft(c * h, F)[: F // 2 + 1]
    D = np.fft.fft(c * dh, F)[: F // 2 + 1]
    T = np.fft.fft(c * th, F)[: F // 2 + 1]
    inv = 1.0 / np.sum(h)
    norm = np.full(F // 2 + 1, 4 * inv * inv)
    norm[0] = norm[-1] = inv * inv
    norm /= float(H) ** 2
    pw = np.abs(S) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        d_omega = -(D.imag * S.real - D.real * S.imag) / pw
        freq = np.arange(F // 2 + 1) * sr / F + d_omega * sr / (2 * np.pi)
        time = (T.real * S.real + T.imag * S.imag) / pw / hop - off / hop
    return dict(power=pw * norm, freq=freq, time=time)
