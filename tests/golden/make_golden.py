#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle on small seeded inputs.

The reference ships no golden vectors (its tests are tolerance bands) and cannot run here (no Rust
toolchain), so these fixtures are produced by the oracle that tests/test_oracle_kat.py pins against the reference's
own known-answer tests.  They freeze today's oracle outputs: the CPU suite checks the oracle still reproduces them,
the GPU suite checks the CUDA build against them without needing the oracle at all.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from openmeters_b200 import _capi as capi  # noqa: E402
from openmeters_b200 import batch, synth  # noqa: E402
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig  # noqa: E402
from oracle import oracle_py  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "cfg1_classic": dict(cfg=SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False), seconds=0.25, lanes=2),
    "cfg2_reassigned": dict(cfg=SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True), seconds=0.30, lanes=2),
    # the product's default spectrogram configuration (spectrogram/processor.rs:47-59) and N = 1024 reassigned: the sizes
    # served by the interleaved-frame kernels (stft_fast2k.cu / stft_fast1k.cu)
    "default_2048_64": dict(cfg=SpectrogramConfig(fft_size=2048, hop_size=64, window=capi.WINDOW_HANN, use_reassignment=True), seconds=0.10, lanes=2),
    "reassigned_1024_256": dict(cfg=SpectrogramConfig(fft_size=1024, hop_size=256, window=capi.WINDOW_HANN, use_reassignment=True), seconds=0.08, lanes=2),
    "small_reassigned_zp4": dict(cfg=SpectrogramConfig(fft_size=256, hop_size=64, window=capi.WINDOW_BLACKMAN, use_reassignment=True, zero_padding_factor=4), seconds=0.03, lanes=1),
}


def inputs(name):
    c = CASES[name]
    return synth.cfg2_lanes(c["lanes"], c["seconds"], seed0=4242)


def main():
    for name, c in CASES.items():
        lanes = inputs(name)
        if c["cfg"].use_reassignment:
            pts, cnt = oracle_py.stft_batch(c["cfg"], lanes, threads=1)
            flat = np.concatenate([pts[l, f, :cnt[l, f]] for l in range(cnt.shape[0]) for f in range(cnt.shape[1])])
            np.savez_compressed(os.path.join(HERE, name + ".npz"), counts=cnt, points=flat)
        else:
            np.savez_compressed(os.path.join(HERE, name + ".npz"), codes=oracle_py.stft_batch(c["cfg"], lanes, threads=1))
    scfg = SpectrumConfig(fft_size=2048, hop_size=256, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0)
    w, r, pk = oracle_py.spectrum_batch(scfg, synth.cfg4_streams(1, 0.2).reshape(2, -1), threads=1)
    freqs = oracle_py.spectrum_frequency_bins(scfg.sample_rate, scfg.fft_size)
    pk2, pf, pl = oracle_py.spectrum_peaks(freqs, w)  # default peak label spec: A-weighted trace, 20 Hz .. last bin
    assert np.array_equal(pk, pk2)
    np.savez_compressed(os.path.join(HERE, "spectrum_peakhold.npz"), weighted=w, raw=r, peak=pk, peak_freq=pf, peak_level=pl)
    x = synth.cfg3_surround(0.7)
    snaps, nb = oracle_py.loudness_batch(LoudnessConfig(), 8, capi.SURROUND, x[None, :], 1024, threads=1)
    np.savez_compressed(os.path.join(HERE, "loudness_surround.npz"), **batch.snapshots_to_arrays(snaps, nb))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
