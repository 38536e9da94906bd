"""CPU: the f32 oracle (and the kernel sources under the emulator) against the float64 restatements written from the Rust
(tests/ref_f64.py) — all four paths.  This is the pin the reference cannot give (no golden vectors, rustfft unobtainable):
the oracle's distance from exact math, per level class, and the measured table that the tolerances of tests/parity.py
rest on."""
import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200 import batch, synth
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig
from oracle import oracle_py
from tests import exact, parity


def _cfg2():
    return SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)


def test_oracle_reassigned_against_float64_cfg2():
    """Power meets the flat 1e-5 rule in every class; frequency / time offsets meet the flat SURVEY tolerances down to
    -40 dB re the column peak and provably cannot below: they are ratios of f32 bins whose noise floor is ~1e-7 of the
    column's amplitude.  The measured table is what parity.py's widening (x peak*1e-4/p below -40 dB) is derived from."""
    lanes = synth.cfg2_lanes(3, 8.0)
    pts, cnt = oracle_py.stft_batch(_cfg2(), lanes)
    t = exact.reassigned_table(pts, cnt, lanes, n=4096, hop=1024, kind=4, sr=48000.0)
    flat = exact.flat_tolerances(n=4096, hop=1024, sr=48000.0)
    assert t.unaligned == 0 and t.columns == cnt.size
    assert np.all(t.mx[:, 0] <= flat[0]), t.to_json()           # power: flat rule everywhere
    assert np.all(t.mx[:4, 1] <= flat[1]) and np.all(t.mx[:4, 2] <= flat[2]), t.to_json()  # >= -40 dB: flat
    # below -40 dB the f32 oracle itself exceeds the flat time tolerance (so no f32 FFT can be held to it) ...
    assert t.mx[5, 2] > flat[2] and t.mx[5, 1] > flat[1], t.to_json()
    # ... and stays inside the widened rule tests/parity.py applies between two f32 implementations
    for c in range(exact.N_CLASS):
        widen = max(1.0, 10.0 ** ((c + 1) * exact.CLASS_DB / 10.0) * 1e-4)
        assert t.mx[c, 1] <= flat[1] * widen and t.mx[c, 2] <= flat[2] * widen, (c, t.to_json())


def test_emulated_kernel_reassigned_against_float64(emu):
    """The cfg2 kernel source under the emulator is as close to exact math as the oracle is."""
    lanes = synth.cfg2_lanes(2, 0.6)
    cfg = _cfg2()
    pa, ca = batch.StftPlan(cfg, api=emu.api).execute_host(lanes)
    pb, cb = oracle_py.stft_batch(cfg, lanes)
    kw = dict(n=4096, hop=1024, kind=4, sr=48000.0)
    exact.assert_reassigned(exact.reassigned_table(pa, ca, lanes, **kw), exact.reassigned_table(pb, cb, lanes, **kw),
                            exact.flat_tolerances(n=4096, hop=1024, sr=48000.0), "emulated cfg2 kernel")


@pytest.mark.parametrize("n,hop,kind,sr", [(4096, 1000, capi.WINDOW_BLACKMAN_HARRIS, 48000.0),    # stft_r64.cu (radix-64 teams, TMEM park)
                                          (16384, 4096, capi.WINDOW_BLACKMAN_HARRIS, 48000.0),  # stft_r64x.cu<256> (64 x 64 x 4)
                                          (8192, 1000, capi.WINDOW_HANN, 96000.0)])             # stft_r64x.cu<128> (64 x 64 x 2)
def test_emulated_team_kernels_against_float64(emu, n, hop, kind, sr):
    """The round-2 team kernels under the emulator against exact math, with the oracle measured beside them (rule of tests/exact.py)."""
    lanes = synth.cfg2_lanes(2, (2 * n + 6 * hop + 16) / 48000.0)[:, :2 * n + 6 * hop]
    cfg = SpectrogramConfig(sample_rate=sr, fft_size=n, hop_size=hop, window=kind, use_reassignment=True)
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=emu.api)
    assert plan.kernel_generation in (7, 8)
    pa, ca = plan.execute_host(lanes)
    pb, cb = oracle_py.stft_batch(cfg, lanes)
    kw = dict(n=n, hop=hop, kind=kind, sr=sr)
    exact.assert_reassigned(exact.reassigned_table(pa, ca, lanes, **kw), exact.reassigned_table(pb, cb, lanes, **kw),
                            exact.flat_tolerances(n=n, hop=hop, sr=sr), f"emulated team kernel N={n}")


@pytest.mark.parametrize("n,hop,kind,zp", [(1024, 512, capi.WINDOW_HANN, 1), (2048, 256, capi.WINDOW_BLACKMAN, 2)])
def test_oracle_classic_against_float64(n, hop, kind, zp):
    st2 = synth.cfg1_stereo(4.0).reshape(-1, 2)
    mid = ((st2[:, 0] + st2[:, 1]) * np.float32(0.5)).astype(np.float32)[None, :]
    cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=kind, use_reassignment=False, zero_padding_factor=zp)
    codes = oracle_py.stft_batch(cfg, mid)
    st = exact.classic_stats(codes, mid, n=n, hop=hop, kind=kind, zp=zp)
    assert st["worst_excess"] <= 1.0 and st["exact_strong"] >= 0.98 and st["max_diff_strong"] <= 1, st


@pytest.mark.parametrize("mode,param", [(capi.AVG_NONE, 0.0), (capi.AVG_EXPONENTIAL, 0.7), (capi.AVG_PEAK_HOLD, 12.0)])
def test_oracle_spectrum_against_float64(mode, param):
    lanes = synth.cfg4_streams(2, 2.0).reshape(4, -1)
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=mode, averaging_param=param, floor_db=-100.0)
    w, r, _ = oracle_py.spectrum_batch(cfg, lanes)
    st = exact.spectrum_stats(w, r, lanes, n=16384, hop=1024, kind=capi.WINDOW_HANN, sr=48000.0, mode=mode, param=param, floor_db=-100.0)
    assert st["worst_raw"] <= 1.0 and st["worst_weighted"] <= 1.0 and st["floor_mismatch"] < 1e-3, st


@pytest.mark.parametrize("sr,ch,positions", [(48000.0, 8, capi.SURROUND), (44100.0, 2, None), (96000.0, 6, None)])
def test_oracle_loudness_against_float64(sr, ch, positions):
    x = synth.cfg3_surround(6.0, sr).reshape(-1, 8)[:, :ch].reshape(-1)
    snaps, nb = oracle_py.loudness_batch(LoudnessConfig(sample_rate=sr), ch, positions, x[None, :], 1024)
    st = exact.loudness_stats(batch.snapshots_to_arrays(snaps, nb), x, ch, positions, sr, 1024)
    # 1e-5 relative on mean squares = 4.3e-5 dB; the true-peak FIR runs in f32 in the reference (36 products of ~1e-7
    # relative error each against a float64 convolution), so its budget is the f32 one: 2e-5 dB
    assert max(st[k] for k in ("short_term", "momentary", "rms_fast", "rms_slow")) <= 5e-5, st
    assert st["true_peak"] <= 5e-5, st


def test_parity_widening_is_peak_relative_and_measured():
    """tests/parity.py compares two f32 implementations; its frequency/time rule is the flat SURVEY tolerance down to
    -40 dB re the column PEAK and widens ∝ peak/p below.  Two oracle runs on inputs that differ by one ulp-scale dither
    stay inside it (sanity of the rule itself)."""
    lanes = synth.cfg2_lanes(1, 1.0)
    cfg = _cfg2()
    pa, ca = oracle_py.stft_batch(cfg, lanes)
    pb, cb = oracle_py.stft_batch(cfg, (lanes * np.float32(1.0 + 2 ** -22)).astype(np.float32))
    parity.compare_reassigned(pa, ca, pb, cb, sr=48000.0, fft_len=4096, window=4096, hop=1024)
