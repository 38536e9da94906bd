"""Row f2 of SURVEY.md §8: splat accumulation + resolve (CUDA scatter-add) against the float64 numpy restatement of
spectrogram.wgsl, on synthetic point clouds and on real reassigned columns."""
import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200 import splat
from oracle import splat_py


def random_rings(seed, n_rings, hl, stride, sr=48000.0):
    rng = np.random.default_rng(seed)
    rings = np.zeros((n_rings, hl, stride, 3), np.float32)
    rings[..., 0] = rng.uniform(-3.0, 1.0, rings.shape[:-1])                   # time_offset in hops
    rings[..., 1] = np.exp(rng.uniform(np.log(0.5), np.log(sr / 2 * 1.05), rings.shape[:-1]))
    rings[..., 2] = 10.0 ** rng.uniform(-15.0, 0.0, rings.shape[:-1])
    rings[..., 2][rng.random(rings.shape[:-1]) < 0.02] = 0.0                   # culled: !(power > 0)
    counts = rng.integers(0, stride + 1, (n_rings, hl)).astype(np.uint32)
    counts[:, ::7] = stride
    return rings, counts


def check(api, rings, counts, p):
    acc, db = splat.render_host(rings, counts, p, api=api)
    ref = splat_py.render(rings, counts, p)
    assert acc.shape == ref.accum.shape
    # every pixel: the non-tie reference power, plus at most the tie power that may land there, within f32 summation noise
    lo = ref.accum * (1 - 2e-5) - 1e-30
    hi = (ref.accum + ref.ties) * (1 + 2e-5) + 1e-30
    assert np.all(acc >= lo) and np.all(acc <= hi), (float(np.max(lo - acc)), float(np.max(acc - hi)))
    # total energy is conserved up to ties that fall off the image edge
    assert abs(acc.sum() - ref.accum.sum()) <= ref.ties.sum() + 1e-5 * ref.accum.sum()
    clean = ref.ties == 0
    lit = clean & (ref.accum > 0)
    assert np.array_equal(np.isneginf(db[clean]), ~(ref.accum[clean] > 0))
    assert np.max(np.abs(db[lit] - ref.db[lit])) < 2e-4
    assert ref.n_drawn > 0
    return ref


CASES = [
    dict(freq_scale=capi.FREQ_LOG, ext_w=300.0, ext_h=200.0, scale_factor=1.0, tilt_db=0.0, uv_y_range=(0.0, 1.0)),
    dict(freq_scale=capi.FREQ_ERB, ext_w=257.5, ext_h=130.25, scale_factor=2.0, tilt_db=3.0, uv_y_range=(0.2, 0.7)),
    dict(freq_scale=capi.FREQ_LINEAR, ext_w=128.0, ext_h=96.0, scale_factor=1.5, tilt_db=-4.5, uv_y_range=(0.0, 0.5)),
]


@pytest.mark.parametrize("case", range(len(CASES)))
def test_splat_emulated_kernel_matches_restatement(emu, case):
    """The kernel source under the CPU emulator (development aid) on a small cloud: all three frequency scales,
    fractional extents, scale factors > 1 (multi-pixel quads), tilt on/off, zoom windows, partial slots, ring wrap."""
    hl, stride = 24, 40
    fmin, fmax = splat.display_axis(48000.0)
    p = splat.SplatParams(freq_min=fmin, freq_max=fmax, ring_capacity=hl, newest_col=5, col_count=hl + 9,
                          reassigned_power_scale=0.37, **CASES[case])
    rings, counts = random_rings(100 + case, 2, hl, stride)
    check(emu.api, rings, counts, p)


def test_splat_emulated_partial_history(emu):
    hl, stride = 16, 12
    fmin, fmax = splat.display_axis(44100.0)
    p = splat.SplatParams(freq_min=fmin, freq_max=fmax, ring_capacity=hl, newest_col=6, col_count=7, ext_w=40.0, ext_h=30.0)
    rings, counts = random_rings(7, 1, hl, stride, 44100.0)
    ref = check(emu.api, rings, counts, p)
    # slots >= col_count are never drawn (render.rs:113)
    p2 = splat.SplatParams(**{**p.__dict__, "col_count": 0})
    acc, db = splat.render_host(rings, counts, p2, api=emu.api)
    assert not acc.any() and np.all(np.isneginf(db)) and ref.n_drawn > 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(CASES)))
def test_splat_gpu_matches_restatement(product, case):
    hl, stride = 64, 257
    fmin, fmax = splat.display_axis(48000.0)
    p = splat.SplatParams(freq_min=fmin, freq_max=fmax, ring_capacity=hl, newest_col=17, col_count=hl + 3,
                          reassigned_power_scale=1.25, **CASES[case])
    rings, counts = random_rings(200 + case, 3, hl, stride)
    check(product.api, rings, counts, p)


@pytest.mark.gpu
def test_splat_of_real_reassigned_columns(product):
    """cfg2 columns straight from the STFT kernel into the splat: a chirp must draw a ridge whose per-column energy
    equals the column's total scaled power (Parseval-style conservation through the scatter-add)."""
    from openmeters_b200 import batch, synth
    from openmeters_b200.processors import SpectrogramConfig

    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    lanes = synth.cfg2_lanes(2, (8192 + 199 * 1024) / 48000.0)
    plan = batch.StftPlan(cfg, api=product.api)
    pts, cnt = plan.execute_host(lanes)            # (lanes*frames, bins, 3), (lanes*frames,)
    frames = cnt.size // 2
    rings = pts.reshape(2, frames, plan.bins, 3)
    counts = cnt.reshape(2, frames).astype(np.uint32)
    fmin, fmax = splat.display_axis(48000.0)
    p = splat.SplatParams(freq_min=fmin, freq_max=fmax, ring_capacity=frames, newest_col=frames - 1, col_count=frames,
                          ext_w=float(frames + 8), ext_h=400.0, reassigned_power_scale=plan.power_scale)
    ref = check(product.api, rings, counts, p)
    acc, _ = splat.render_host(rings, counts, p, api=product.api)
    total_points = sum(float(rings[r, s, :counts[r, s], 2].astype(np.float64).sum()) for r in range(2) for s in range(frames))
    assert acc.sum() <= total_points * (1 + 1e-5)
    assert acc.sum() >= 0.9 * total_points         # only points reassigned off the visible time / frequency range are lost


def _render_vs_two_step(api, n, hop, window, lanes, ext_h=96.0, scale=1.0):
    """omb_stft_render_host (STFT -> splat -> resolve chained on the device, only images come back) against the two public
    steps it fuses: omb_stft_execute_host, then omb_splat_render_host on the returned points."""
    from openmeters_b200 import batch
    from openmeters_b200.processors import SpectrogramConfig

    cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=window, use_reassignment=True)
    plan = batch.StftPlan(cfg, api=api)
    pts, cnt = plan.execute_host(lanes)
    L, F = cnt.shape
    fmin, fmax = splat.display_axis(48000.0)
    view = splat.SplatParams(freq_min=fmin, freq_max=fmax, ext_w=float(F) * scale + 3.0, ext_h=ext_h, scale_factor=scale)
    db, cnt2 = plan.render_host(lanes, view)
    assert np.array_equal(cnt, cnt2)
    full = splat.SplatParams(**{**view.__dict__, "ring_capacity": F, "newest_col": F - 1, "col_count": F,
                                "reassigned_power_scale": plan.power_scale})
    acc_ref, db_ref = splat.render_host(pts, cnt, full, api=api)
    assert db.shape == db_ref.shape
    assert np.array_equal(np.isneginf(db), np.isneginf(db_ref))
    lit = ~np.isneginf(db_ref)
    assert lit.mean() > 0.05
    assert np.max(np.abs(db[lit] - db_ref[lit])) < 1e-3      # f32 atomics: summation order differs between the two runs
    return db


def test_stft_render_host_emulated(emu):
    from openmeters_b200 import synth

    lanes = synth.cfg2_lanes(5, (8192 + 6 * 1024) / 48000.0)
    _render_vs_two_step(emu.api, 4096, 1024, capi.WINDOW_BLACKMAN_HARRIS, lanes)


@pytest.mark.gpu
@pytest.mark.parametrize("n,hop,lanes,secs", [(4096, 1024, 13, 2.0), (2048, 64, 2, 0.4)])
def test_stft_render_host_gpu(product, n, hop, lanes, secs):
    from openmeters_b200 import synth

    x = synth.cfg2_lanes(lanes, secs)
    db = _render_vs_two_step(product.api, n, hop, capi.WINDOW_BLACKMAN_HARRIS if n == 4096 else capi.WINDOW_HANN, x, ext_h=300.0, scale=1.5)
    # a chirp draws a ridge: the brightest pixel of the newest column region is far above the median lit pixel
    lit = db[np.isfinite(db)]
    assert lit.max() > np.median(lit) + 30.0
