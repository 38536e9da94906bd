"""Row f3 — the spectrum view's peak label: peak_bin (spectrum/state.rs:321-325) and interpolated_peak
(state.rs:327-356).

The reference has no #[test] for these two functions, so the oracle restatement is pinned by hand-derived
known answers (each case cites the source lines it exercises); the CUDA kernel `k_peak_interpolate` is then checked
against the oracle on the same dB values, bit for bit (the function is plain f32 arithmetic), under the emulator
on the CPU and on the real build with `-m gpu`.  The fused arg-max is covered by tests/cases.py::spectrum_parity."""
import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200 import batch
from openmeters_b200.processors import SpectrumConfig
from oracle import oracle_py

F32 = np.float32
BINS = np.arange(8, dtype=np.float32) * F32(10.0)  # 0, 10, ..., 70 Hz


def _peaks(db, min_f=20.0, max_f=None, bins=BINS):
    b, f, m = oracle_py.spectrum_peaks(bins, np.asarray(db, np.float32)[None, :], min_f, max_f)
    return int(b[0]), float(f[0]), float(m[0])


# ------------------------------------------------------------------ oracle known answers
def test_peak_bin_skips_edges_and_range():
    """:322-323 — candidates are 1..len-2 and must lie in [min_f, max_f] (inclusive on both ends)."""
    db = [0.0, -1.0, -5.0, -6.0, -7.0, -8.0, -2.0, 0.0]  # bins 0 and 7 hold the largest values but are never candidates
    assert _peaks(db, min_f=0.0)[0] == 1
    assert _peaks(db, min_f=20.0)[0] == 6        # bin 1 (10 Hz) is below min_f; bin 2 (20 Hz) is inside (inclusive)
    assert _peaks(db, min_f=20.0, max_f=50.0)[0] == 2
    assert _peaks(db, min_f=20.0, max_f=60.0)[0] == 6   # 60 Hz inclusive
    assert _peaks(db, min_f=61.0, max_f=65.0)[0] == -1  # no candidate -> None


def test_peak_bin_last_maximum_wins_and_non_finite_skipped():
    """:324 Iterator::max_by returns the LAST maximum; :323 non-finite values are filtered out."""
    assert _peaks([0, -9.0, -3.0, -9.0, -3.0, -9.0, -9.0, 0], min_f=0.0)[0] == 4
    assert _peaks([0, 0, np.inf, np.nan, -3.0, -9.0, -9.0, 0], min_f=0.0)[0] == 1
    assert _peaks([0, np.nan, np.nan, np.nan, np.nan, np.nan, np.nan, 0], min_f=0.0)[0] == -1
    # total_cmp orders -0.0 below +0.0
    assert _peaks([0, -1.0, 0.0, -0.0, -1.0, -1.0, -1.0, 0], min_f=0.0)[0] == 2
    assert _peaks([0, -1.0, -0.0, 0.0, -1.0, -1.0, -1.0, 0], min_f=0.0)[0] == 3


def test_default_max_f_is_last_bin_or_just_above_min():
    """:107 max_f = frequency_bins[last].max(min_f * 1.02)."""
    db = [0, -9.0, -9.0, -9.0, -9.0, -9.0, -1.0, 0]
    assert _peaks(db)[0] == 6
    tiny = np.arange(4, dtype=np.float32) * F32(5.0)  # 0, 5, 10, 15 Hz: last bin < 20 Hz, so max_f = 20.4 and nothing qualifies
    b, _, _ = oracle_py.spectrum_peaks(tiny, np.array([[0, -1.0, -2.0, 0]], np.float32))
    assert b[0] == -1


def test_interpolated_peak_parabola():
    """:339-355 — offset = 0.5 (l - r) / (l - 2c + r), level = c - 0.25 (l - r) offset, freq = f_c + offset bin_hz."""
    db = [0, -20.0, -10.0, -4.0, -6.0, -20.0, -20.0, 0]
    b, f, m = _peaks(db, min_f=0.0)
    assert b == 3
    # denom = -10 + 8 - 6 = -8; offset = 0.5 * (-4) / (-8) = 0.25; level = -4 - (0.25 * -4) * 0.25 = -3.75
    assert f == 30.0 + 0.25 * 10.0 and m == -3.75
    # mirrored: the vertex moves the other way
    b, f, m = _peaks([0, -20.0, -6.0, -4.0, -10.0, -20.0, -20.0, 0], min_f=0.0)
    assert (b, f, m) == (3, 30.0 - 2.5, -3.75)


def test_interpolated_peak_flat_top_clamp_and_non_finite_neighbours():
    # :342 denom >= -EPSILON (flat or convex) -> offset 0, level = centre
    # two equal maxima: the last one (bin 3) is the peak; l = c = -4, r = -9: denom = -5, offset = 0.5 * 5 / -5 = -0.5,
    # level = -4 - (0.25 * 5) * -0.5 = -3.375
    assert _peaks([0, -9.0, -4.0, -4.0, -9.0, -9.0, -9.0, 0], min_f=0.0) == (3, 25.0, -3.375)
    b, f, m = _peaks([0, -5.0, -5.0, -5.0, -5.0, -5.0, -5.0, 0], min_f=0.0)
    assert (b, f, m) == (6, 60.0, -5.0)  # right neighbour is 0 dB: denom = -5 + 10 + 0 = 5 > 0 -> no offset
    # :343 clamp to +-0.5 — candidates limited to bin 3 whose right neighbour is higher
    b, f, m = _peaks([0, -30.0, -30.0, -10.0, -9.0, -30.0, -30.0, 0], min_f=30.0, max_f=30.0)
    # denom = -30 + 20 - 9 = -19; 0.5 * (-21) / -19 = 0.5526 -> clamped to 0.5; level = -10 - (0.25 * -21) * 0.5 = -7.375
    assert (b, f, m) == (3, 35.0, -7.375)
    # :339-349 a non-finite neighbour disables the interpolation but not the peak
    assert _peaks([0, -30.0, np.nan, -10.0, -12.0, -30.0, -30.0, 0], min_f=0.0) == (3, 30.0, -10.0)
    assert _peaks([0, -30.0, -12.0, -10.0, -np.inf, -30.0, -30.0, 0], min_f=0.0) == (3, 30.0, -10.0)


def test_interpolated_peak_none_cases():
    """:328-337 — bin 0, missing right neighbour, non-finite centre."""
    db = np.array([[-1.0, -2.0, -3.0, -4.0]], np.float32)
    bins = BINS[:4]
    for bad in (0, 3, -1, 7):
        f, m = oracle_py.spectrum_interpolate_peaks(bins, db, np.array([bad], np.int32))
        assert np.isnan(f[0]) and np.isnan(m[0]), bad
    f, m = oracle_py.spectrum_interpolate_peaks(bins, np.array([[-1.0, np.inf, -3.0, -4.0]], np.float32), np.array([1], np.int32))
    assert np.isnan(f[0]) and np.isnan(m[0])
    f, m = oracle_py.spectrum_interpolate_peaks(bins, db, np.array([1], np.int32))
    # l = -1, c = -2, r = -3: denom = -1 + 4 - 3 = 0 -> no offset
    assert (f[0], m[0]) == (10.0, -2.0)


# ------------------------------------------------------------------ the CUDA kernel vs the oracle
def _random_traces(rows, bins, seed):
    rng = np.random.default_rng(seed)
    db = (rng.standard_normal((rows, bins)) * 12.0 - 40.0).astype(np.float32)
    # smooth every other row so that real parabolic peaks (small curvature) occur as well as ragged ones
    sm = np.cumsum(db, axis=1, dtype=np.float32) / np.arange(1, bins + 1, dtype=np.float32)
    db[::2] = sm[::2]
    db[rng.random((rows, bins)) < 0.01] = np.float32(-100.0)       # floor plateaus
    db[rng.random((rows, bins)) < 0.002] = np.nan
    db[rng.random((rows, bins)) < 0.002] = -np.inf
    pk = rng.integers(1, bins - 1, size=rows).astype(np.int32)
    pk[-16:] = np.tile(np.array([-1, 0, bins - 1, bins], np.int32), 4)  # every out-of-range None case
    db[np.arange(rows - 32, rows - 16), pk[-32:-16]] = np.nan               # non-finite centre -> None
    pk[: rows // 2] = np.nanargmax(np.where(np.isfinite(db[: rows // 2, 1:-1]), db[: rows // 2, 1:-1], -np.inf), axis=1) + 1
    return db, pk


def _check_kernel(api, to_device, from_device):
    cfg = SpectrumConfig(sample_rate=44100.0, fft_size=512, hop_size=128)
    plan = batch.SpectrumPlan(cfg, api=api)
    bins = cfg.fft_size // 2 + 1
    freqs = oracle_py.spectrum_frequency_bins(cfg.sample_rate, cfg.fft_size)
    db, pk = _random_traces(300, bins, 7)
    d_db, d_pk = to_device(db), to_device(pk)
    d_f, d_m = to_device(np.zeros(300, np.float32)), to_device(np.zeros(300, np.float32))
    plan.interpolate_peaks_device(d_db[1], d_pk[1], 300, d_f[1], d_m[1])
    f, m = from_device(d_f), from_device(d_m)
    ef, em = oracle_py.spectrum_interpolate_peaks(freqs, db, pk)
    assert np.array_equal(f.view(np.uint32), ef.view(np.uint32))
    assert np.array_equal(m.view(np.uint32), em.view(np.uint32))
    assert np.isnan(ef).sum() > 10 and np.isfinite(ef).sum() > 150 and (ef != freqs[np.clip(pk, 0, bins - 1)])[np.isfinite(ef)].sum() > 50


def test_interpolate_kernel_under_emulator(emu):
    # the emulator's "device" memory is host memory: numpy buffers are passed straight through
    _check_kernel(emu.api, lambda a: (a, a.ctypes.data), lambda d: d[0])


def test_peak_spec_validation(emu):
    plan = batch.SpectrumPlan(SpectrumConfig(fft_size=256, hop_size=64), api=emu.api)
    assert plan.peak_spec() == (0, 20.0, 0.0)  # reference defaults: A-weighted, MIN_FREQUENCY, up to the last bin
    with pytest.raises(Exception):
        plan.set_peak_spec(trace=2)
    plan.set_peak_spec(trace=1, min_hz=1.0e6)  # empty candidate range: every hop reports None
    x = np.sin(np.arange(1024, dtype=np.float32) * 0.3).astype(np.float32)[None, :]
    _, _, pk, f, m = plan.execute_host_peaks(x)
    assert np.all(pk == -1) and np.all(np.isnan(f)) and np.all(np.isnan(m))


@pytest.mark.gpu
def test_interpolate_kernel_on_gpu(product):
    import torch

    def to_device(a):
        t = torch.from_numpy(a).cuda()
        return t, t.data_ptr()

    def from_device(d):
        torch.cuda.synchronize()
        return d[0].cpu().numpy()

    _check_kernel(product.api, to_device, from_device)


@pytest.mark.gpu
def test_peak_label_of_a_pure_tone_on_gpu(product):
    """End to end: a 1 kHz tone between two bins — the interpolated label must land within 0.1 bin of 1 kHz (Hann window,
    parabolic interpolation of dB values), in both trace modes."""
    sr, n = 48000.0, 4096
    t = np.arange(3 * n, dtype=np.float64)
    x = (0.5 * np.sin(2 * np.pi * 1000.0 * t / sr)).astype(np.float32)[None, :]
    plan = batch.SpectrumPlan(SpectrumConfig(sample_rate=sr, fft_size=n, hop_size=n // 4), api=product.api)
    for trace in (0, 1):
        plan.set_peak_spec(trace=trace)
        w, r, pk, f, m = plan.execute_host_peaks(x)
        assert np.all(np.abs(f - 1000.0) < 0.1 * sr / n), f
        tr = (w, r)[trace]
        assert np.all(m >= tr[0, np.arange(pk.shape[1]), pk[0]])
