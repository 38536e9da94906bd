"""`-m gpu`: CUDA path vs the CPU oracle on the BASELINE configs (same seeded input bytes), through the C ABI,
plus size-independent properties at larger sizes."""
import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200 import batch, synth
from openmeters_b200.processors import AudioBlock, LoudnessConfig, SpectrogramConfig, SpectrumConfig
from tests import cases, parity

pytestmark = pytest.mark.gpu

KERNELS = [capi.KERNEL_GENERIC, capi.KERNEL_AUTO]


# ---------------------------------------------------------------- cfg1: classic, streaming, stereo -> Mid
def test_cfg1_streaming_classic_stereo(product, oracle):
    x = synth.cfg1_stereo(10.0)
    cfg = SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False, history_length=8192)
    cols = {}
    for b in (product, oracle):
        p = b.Spectrogram(cfg)
        out = []
        first_reset = None
        for s in range(0, x.size, 2 * 1024):
            up = p.process_block(AudioBlock(x[s:s + 2 * 1024], 2, 48000.0))
            if up is not None:
                if first_reset is None:
                    first_reset = up.reset
                assert up.kind == capi.COLUMN_CLASSIC
                out.extend(up.new_columns)
        assert first_reset is True
        cols[b.name] = np.stack(out)
    assert cols["product"].shape == cols["oracle"].shape == (936, 513)
    st = parity.compare_classic(cols["product"][None], cols["oracle"][None])
    assert st["exact"] >= 0.98


@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg1_batch_classic(product, kernel):
    st2 = synth.cfg1_stereo(10.0).reshape(-1, 2)
    mid = ((st2[:, 0] + st2[:, 1]) * np.float32(0.5)).astype(np.float32)
    cfg = SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False)
    st = cases.stft_parity(product.api, cfg, mid[None, :], kernel=kernel)
    assert st["exact"] >= 0.98


# ---------------------------------------------------------------- cfg2: the metric path
@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg2_reassigned_batch(product, kernel):
    lanes = synth.cfg2_lanes(8, 12.0)
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    st = cases.stft_parity(product.api, cfg, lanes, kernel=kernel)
    assert st["cols"] == 8 * 555 and st["checked"] > 0.5 * st["pts"]


def test_cfg2_streaming_matches_batch(product):
    """The streaming processor is a thin wrapper over the batch kernel: same columns, block-partition independent."""
    lane = synth.cfg2_lanes(1, 3.0)[0]
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True, history_length=8192)
    plan = batch.StftPlan(cfg, api=product.api)
    pts, cnt = plan.execute_host(lane[None, :])
    for block in (1024, 3000):
        p = product.Spectrogram(cfg)
        cols = []
        for s in range(0, lane.size, block):
            up = p.process_block(AudioBlock(lane[s:s + block], 1, 48000.0))
            if up is not None:
                cols.extend(up.new_columns)
        assert len(cols) == cnt.shape[1]
        for f, c in enumerate(cols):
            assert np.array_equal(c, pts[0, f, :cnt[0, f]])


@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg2_properties_large(product, kernel):
    """Size-independent properties on a batch the oracle would take minutes for:
    frame independence (a frame's result does not depend on where it sits in the batch), exact power-of-two
    scaling (x2 input -> x4 power bit-exactly, same time/frequency), physical ranges."""
    L, S = 16, 1 << 20
    lanes = synth.cfg2_lanes(L, S / 48000.0)[:, :S]
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    plan = batch.StftPlan(cfg, kernel=kernel, api=product.api)
    pts, cnt = plan.execute_host(lanes)
    F = cnt.shape[1]
    assert F == (S - 8192) // 1024 + 1
    assert cnt.max() <= 2049 and cnt.min() > 0
    # frame independence: recompute a shifted sub-batch
    sub = np.ascontiguousarray(lanes[3:5, 100 * 1024: 100 * 1024 + 8192 + 49 * 1024])
    p2, c2 = plan.execute_host(sub)
    assert np.array_equal(c2, cnt[3:5, 100:150])
    for l in range(2):
        for f in range(50):
            assert np.array_equal(p2[l, f, :c2[l, f]], pts[3 + l, 100 + f, :cnt[3 + l, 100 + f]])
    # scaling
    p4, c4 = plan.execute_host((sub * np.float32(2.0)).astype(np.float32))
    same = c4 == c2
    assert same.mean() > 0.99
    for l, f in zip(*np.nonzero(same)):
        n = c2[l, f]
        assert np.array_equal(p4[l, f, :n, :2], p2[l, f, :n, :2])
        assert np.array_equal(p4[l, f, :n, 2], p2[l, f, :n, 2] * np.float32(4.0))
    # ranges
    for l in range(0, L, 5):
        for f in range(0, F, 97):
            q = pts[l, f, :cnt[l, f]]
            assert np.all(q[:, 1] > 0) and np.all(q[:, 1] < 24000.0) and np.all(q[:, 2] >= 1e-14) and np.all(np.isfinite(q))


def test_generic_and_fast_kernels_agree(product):
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    fast = batch.StftPlan(cfg, kernel=capi.KERNEL_AUTO, api=product.api)
    if not fast.is_fast:
        pytest.skip("no specialised kernel for this size in this build")
    gen = batch.StftPlan(cfg, kernel=capi.KERNEL_GENERIC, api=product.api)
    lanes = synth.cfg2_lanes(4, 4.0, seed0=77)
    pa, ca = fast.execute_host(lanes)
    pb, cb = gen.execute_host(lanes)
    parity.compare_reassigned(pa, ca, pb, cb, sr=48000.0, fft_len=4096, window=4096, hop=1024)


# ---------------------------------------------------------------- cfg3: loudness
def test_cfg3_loudness_batch(product):
    x = synth.cfg3_surround(30.0)
    st = cases.loudness_parity(product.api, LoudnessConfig(), 8, capi.SURROUND, x[None, :], 1024)
    assert max(st.values()) <= 5e-5


def test_cfg3_loudness_streaming(product, oracle):
    x = synth.cfg3_surround(8.0)
    snaps = {}
    for b in (product, oracle):
        p = b.Loudness(LoudnessConfig())
        out = []
        for s in range(0, x.size, 8 * 1024):
            out.append(p.process_block(AudioBlock(x[s:s + 8 * 1024], 8, 48000.0, capi.SURROUND)))
        snaps[b.name] = out
    for a, o in zip(snaps["product"], snaps["oracle"]):
        assert abs(a.short_term_loudness - o.short_term_loudness) <= 5e-5
        assert abs(a.momentary_loudness - o.momentary_loudness) <= 5e-5
        assert np.max(np.abs(a.rms_fast_db - o.rms_fast_db)) <= 5e-5 and np.max(np.abs(a.rms_slow_db - o.rms_slow_db)) <= 5e-5
        assert np.array_equal(a.true_peak_db, o.true_peak_db) or np.max(np.abs(a.true_peak_db - o.true_peak_db)) <= 1e-5
        assert a.channel_count == 8 and a.positions == o.positions


def test_loudness_rates_and_layouts(product):
    for sr, ch in ((44100.0, 2), (96000.0, 6), (192000.0, 1)):
        x = synth.cfg3_surround(1.5, sr).reshape(-1, 8)[:, :ch].reshape(-1)
        cases.loudness_parity(product.api, LoudnessConfig(sample_rate=sr), ch, None, x[None, :], 777)


# ---------------------------------------------------------------- cfg4: spectrum analyzer
def test_cfg4_spectrum_batch(product):
    lanes = synth.cfg4_streams(6, 5.0).reshape(12, -1)
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD,
                         averaging_param=12.0, floor_db=-100.0)
    cases.spectrum_parity(product.api, cfg, lanes)


@pytest.mark.parametrize("mode,param", [(capi.AVG_PEAK_HOLD, 12.0), (capi.AVG_EXPONENTIAL, 0.7), (capi.AVG_NONE, 0.0)])
def test_cfg4_spectrum_fused_kernel(product, mode, param, monkeypatch):
    """The whole-batch fused kernel (one CTA per lane: FFT + smoothing + dB + arg-max), pinned on for a small lane
    count, against the oracle; and the auto-selected path with enough lanes to fill the GPU against the two-kernel
    path on the same data (same power values up to the epilogue twiddle rounding)."""
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=mode, averaging_param=param, floor_db=-100.0)
    monkeypatch.setenv("OMB_SPECTRUM_FUSED", "1")
    lanes = synth.cfg4_streams(3, 3.0).reshape(6, -1)
    cases.spectrum_parity(product.api, cfg, lanes)
    if mode != capi.AVG_PEAK_HOLD:
        return
    many = np.ascontiguousarray(np.concatenate([np.roll(lanes, 331 * r, axis=1) * np.float32(1 - 0.03 * r) for r in range(16)], 0)[:, :16384 + 40 * 1024])
    monkeypatch.delenv("OMB_SPECTRUM_FUSED")
    plan = batch.SpectrumPlan(cfg, api=product.api)
    wf, rf, pf = plan.execute_host(many)            # 96 lanes: fused kernel
    monkeypatch.setenv("OMB_SPECTRUM_FUSED", "0")
    w2, r2, p2 = plan.execute_host(many)            # two-kernel path
    parity.compare_db(rf, r2, cfg.floor_db)
    parity.compare_db(wf, w2, cfg.floor_db)
    assert np.mean(pf == p2) > 0.999


@pytest.mark.parametrize("mode,param", [(capi.AVG_NONE, 0.0), (capi.AVG_EXPONENTIAL, 0.5)])
def test_spectrum_modes(product, mode, param):
    lanes = synth.cfg4_streams(2, 1.5).reshape(4, -1)
    cfg = SpectrumConfig(fft_size=4096, hop_size=512, window=capi.WINDOW_BLACKMAN, averaging=mode, averaging_param=param)
    cases.spectrum_parity(product.api, cfg, lanes)


def test_cfg4_streaming_two_sources(product, oracle):
    x = synth.cfg4_streams(1, 3.0)[0]  # (2, S) planar
    inter = np.stack([x[0], x[1]], 1).reshape(-1).astype(np.float32)
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0,
                         source=capi.CHANNEL_LEFT, secondary_source=capi.CHANNEL_RIGHT)
    last = {}
    for b in (product, oracle):
        p = b.Spectrum(cfg)
        snap = None
        for s in range(0, inter.size, 2 * 1024):
            r = p.process_block(AudioBlock(inter[s:s + 2 * 1024], 2, 48000.0))
            snap = r if r is not None else snap
        last[b.name] = snap
    a, o = last["product"], last["oracle"]
    assert np.array_equal(a.frequency_bins, o.frequency_bins)
    for t in range(2):
        for w in range(2):
            parity.compare_db(a.traces[t][w][None], o.traces[t][w][None], -100.0)


def test_cfg4_two_kernel_path_ring_power(product, monkeypatch):
    """Few lanes -> two-kernel path with the front half of the fused kernel as its power stage ((lane, hop segment) work items, each
    priming its own ring): same results as the frame-per-CTA power kernel within the parity metric; peak-hold state carries across
    the chunk boundaries (4 chunks here)."""
    monkeypatch.setenv("OMB_SPECTRUM_FUSED", "0")
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, floor_db=-100.0)
    lanes = synth.cfg4_streams(8, 22.0).reshape(16, -1)
    monkeypatch.setenv("OMB_SPECTRUM_RING_POWER", "1")
    a = batch.SpectrumPlan(cfg, api=product.api).execute_host(lanes)
    monkeypatch.setenv("OMB_SPECTRUM_RING_POWER", "0")
    b = batch.SpectrumPlan(cfg, api=product.api).execute_host(lanes)
    assert np.max(np.abs(a[0] - b[0])) < 5e-3 and np.max(np.abs(a[1] - b[1])) < 5e-3  # dB, two f32 transforms
    assert np.mean(a[2] == b[2]) > 0.999                                              # peak bins, up to near-ties


# ---------------------------------------------------------------- cfg5: 8192-pt reassigned at 96 kHz
@pytest.mark.parametrize("kernel", KERNELS)
def test_cfg5_reassigned_batch(product, kernel):
    lanes = synth.cfg5_lanes(4, 96000 * 2)
    cfg = SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=2048, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    st = cases.stft_parity(product.api, cfg, lanes, kernel=kernel)
    assert st["cols"] == 4 * ((96000 * 2 - 16384) // 2048 + 1)


# ---------------------------------------------------------------- N = 16384 reassigned, the UI's largest size, on chip (stft_r64x.cu)
@pytest.mark.gpu
@pytest.mark.parametrize("hop,window", [(4096, capi.WINDOW_BLACKMAN_HARRIS), (256, capi.WINDOW_HANN), (2732, capi.WINDOW_BLACKMAN)])
def test_size_16384_reassigned(product, hop, window):
    """One CTA per frame, 64 x 64 x 4 transforms with the analysis input / S / nd parked in tensor memory; 300+ frames over 148 SMs so
    every CTA walks several frames; vs the oracle, and the specialised kernel must be the one that ran."""
    cfg = SpectrogramConfig(fft_size=16384, hop_size=hop, window=window, use_reassignment=True)
    frames = 101
    n = 32768 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(3, (n + 64) / 48000.0)[:, :n]
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=product.api)
    assert plan.kernel_generation == 8
    st = cases.stft_parity(product.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 3 * frames and st["checked"] > 100000


@pytest.mark.gpu
def test_size_16384_specialised_and_generic_agree(product, monkeypatch):
    """The on-chip kernel against the global-scratch generic tier on the GPU (pairwise metric)."""
    cfg = SpectrogramConfig(fft_size=16384, hop_size=2048, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    n = 32768 + 20 * 2048
    lanes = synth.cfg2_lanes(2, (n + 64) / 48000.0)[:, :n]
    pa, ca = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=product.api).execute_host(lanes)
    pb, cb = batch.StftPlan(cfg, kernel=capi.KERNEL_GENERIC, api=product.api).execute_host(lanes)
    st = parity.compare_reassigned(pa, ca, pb, cb, sr=48000.0, fft_len=16384, window=16384, hop=2048)
    assert st["cols"] == 42 and st["checked"] > 100000


@pytest.mark.parametrize("n,hop,sr", [(4096, 1000, 48000.0), (16384, 4096, 48000.0), (8192, 1000, 96000.0)])
def test_team_kernels_streaming_matches_batch(product, n, hop, sr):
    """Streaming processor over stft_r64.cu / stft_r64x.cu (hops off the ring kernels' grid): ragged blocks -> calls with first_frame > 0;
    every column equals the batch path's bit for bit (so the same kernel served the stream), and the batch path matches the oracle."""
    cfg = SpectrogramConfig(sample_rate=sr, fft_size=n, hop_size=hop, window=capi.WINDOW_HANN, use_reassignment=True, history_length=64)
    frames = 23
    S = 2 * n + (frames - 1) * hop
    lanes = synth.cfg2_lanes(1, (S + 64) / 48000.0)[:, :S]
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=product.api)
    assert plan.kernel_generation in (7, 8)
    pts, cnt = plan.execute_host(lanes)
    p = product.Spectrogram(cfg)
    cols = []
    step = 4 * 997
    for s0 in range(0, S, step):
        up = p.process_block(AudioBlock(lanes[0, s0:min(s0 + step, S)], 1, sr))
        if up is not None:
            cols += list(up.new_columns)
    assert len(cols) == frames
    for f, c in enumerate(cols):
        assert np.array_equal(np.asarray(c), pts[0, f, :cnt[0, f]]), f
    cases.stft_parity(product.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)


# ---------------------------------------------------------------- N = 4096 at the UI's small hops (N/16 ... N/128)
@pytest.mark.parametrize("hop", [256, 64, 32])
def test_cfg2_small_hops(product, hop):
    cfg = SpectrogramConfig(fft_size=4096, hop_size=hop, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    frames = 601
    n = 8192 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(3, (n + 64) / 48000.0)[:, :n]
    st = cases.stft_parity(product.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 3 * frames
    # the streaming processor takes the same kernel (its FIFO stays 16-byte aligned at these hops)
    p = product.Spectrogram(SpectrogramConfig(fft_size=4096, hop_size=hop, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True,
                                              history_length=8192))
    cols = []
    for s0 in range(0, 8192 + 40 * hop, 1000):
        up = p.process_block(AudioBlock(lanes[0, s0:min(s0 + 1000, 8192 + 40 * hop)], 1, 48000.0))
        if up is not None:
            cols += list(up.new_columns)
    ref_pts, ref_cnt = cases.oracle_py.stft_batch(cfg, lanes[:1, :8192 + 40 * hop])
    assert len(cols) == ref_cnt.shape[1] == 41
    for f, c in enumerate(cols):
        parity.compare_reassigned_column(np.asarray(c), ref_pts[0, f, :ref_cnt[0, f]], sr=48000.0, fft_len=4096, window=4096, hop=hop)


# ---------------------------------------------------------------- N = 2048: the product's default analysis size
@pytest.mark.parametrize("hop,window", [(64, capi.WINDOW_HANN), (512, capi.WINDOW_BLACKMAN_HARRIS), (16, capi.WINDOW_HANN)])
def test_default_size_2048(product, hop, window):
    """spectrogram/processor.rs:47-59 defaults (2048 / 64 / Hann / reassigned) and neighbours through stft_fast2k.cu, against the
    oracle and against the shared-memory tier on the same device."""
    cfg = SpectrogramConfig(fft_size=2048, hop_size=hop, window=window, use_reassignment=True)
    frames = 803
    n = 4096 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(3, (n + 64) / 48000.0)[:, :n]
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=product.api)
    assert plan.kernel_generation == 5
    st = cases.stft_parity(product.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 3 * frames
    pa, ca = plan.execute_host(lanes)
    pb, cb = batch.StftPlan(cfg, kernel=capi.KERNEL_GENERIC, api=product.api).execute_host(lanes)
    parity.compare_reassigned(pa, ca, pb, cb, sr=48000.0, fft_len=2048, window=2048, hop=hop)


@pytest.mark.parametrize("hop,window", [(256, capi.WINDOW_HANN), (32, capi.WINDOW_BLACKMAN_HARRIS)])
def test_size_1024_reassigned(product, hop, window):
    """N = 1024 reassigned through stft_fast1k.cu (four interleaved frames per transform), against the oracle and the generic kernel."""
    cfg = SpectrogramConfig(fft_size=1024, hop_size=hop, window=window, use_reassignment=True)
    frames = 1203
    n = 2048 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(3, (n + 64) / 48000.0)[:, :n]
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=product.api)
    assert plan.kernel_generation == 6
    st = cases.stft_parity(product.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 3 * frames
    pa, ca = plan.execute_host(lanes)
    pb, cb = batch.StftPlan(cfg, kernel=capi.KERNEL_GENERIC, api=product.api).execute_host(lanes)
    parity.compare_reassigned(pa, ca, pb, cb, sr=48000.0, fft_len=1024, window=1024, hop=hop)


# ---------------------------------------------------------------- edge cases
def test_edge_cases(product):
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    plan = batch.StftPlan(cfg, api=product.api)
    # too short: no frames, no launch, no error
    pts, cnt = plan.execute_host(np.zeros((2, 8191), np.float32))
    assert cnt.shape == (2, 0)
    # exactly one frame; all-zero lane -> empty column; DC lane -> empty column (bin 0 removed by the Hilbert mask)
    lanes = np.zeros((3, 8192), np.float32)
    lanes[1] = 0.25
    lanes[2] = synth.cfg2_lanes(1, 8192 / 48000.0)[0, :8192]
    pts, cnt = plan.execute_host(lanes)
    assert cnt.shape == (3, 1) and cnt[0, 0] == 0 and cnt[1, 0] == 0 and cnt[2, 0] > 1000
    # ragged stride
    big = np.zeros((2, 20000), np.float32)
    big[:, :8192 + 1024] = lanes[2:3, :1].repeat(2, 0) * 0 + synth.cfg2_lanes(2, 0.2)[:, :8192 + 1024]
    p1, c1 = plan.execute_host(np.ascontiguousarray(big[:, :8192 + 1024]))
    assert c1.shape == (2, 2)


# ---------------------------------------------------------------- the product's config space
@pytest.mark.parametrize("n,hop,zp,window,reassign", cases.settings_grid())
def test_settings_grid(product, n, hop, zp, window, reassign):
    """Every reachable (size, hop, zero padding, window, mode) combination has a kernel and matches the oracle —
    whichever tier (specialised / shared-memory / generic) the plan picks."""
    cases.settings_grid_case(product.api, n, hop, zp, window, reassign)


# ---------------------------------------------------------------- size-independent properties of the other rows, at full BASELINE sizes
def test_cfg1_properties_large(product):
    """cfg1 at a size the oracle is not run on (256 lanes x 2^18 samples = 1.3e5 frames): frame independence — a column depends
    only on its 1024 samples, so a batch shifted by whole hops reproduces the same u16 codes bit for bit — plus the code range."""
    L, S = 256, 1 << 18
    base = synth.cfg2_lanes(8, S / 48000.0)[:, :S]
    lanes = np.concatenate([np.roll(base, 977 * r, axis=1) * np.float32(1.0 - 0.01 * r) for r in range(L // 8)], 0)
    cfg = SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False)
    plan = batch.StftPlan(cfg, api=product.api)
    codes = plan.execute_host(lanes)
    F = (S - 1024) // 512 + 1
    assert codes.shape == (L, F, 513)
    sub = np.ascontiguousarray(lanes[40:43, 200 * 512: 200 * 512 + 1024 + 99 * 512])
    assert np.array_equal(plan.execute_host(sub), codes[40:43, 200:300])
    assert codes.min() >= 1680  # -140 dB floor = code round(4 * 65535 / 156); nothing below it (spectrogram/processor.rs:103-108)
    # a louder copy never codes lower, and +6.0206 dB is 2529.2 codes wherever neither side clips at the floor
    c2 = plan.execute_host((sub * np.float32(2.0)).astype(np.float32)).astype(np.int64)
    c1 = codes[40:43, 200:300].astype(np.int64)
    assert np.all(c2 >= c1)
    free = c1 > 1680
    assert np.all(np.abs((c2 - c1)[free] - 2529.24) <= 1.1)


def test_cfg3_properties_large(product):
    """cfg3 at BASELINE size (16 streams x 30 s x 8 ch): gain linearity of every output (x0.5 -> -6.0206 dB exactly where the
    floor is not hit; f64 mean squares scale exactly by 0.25, true-peak FIR sums scale exactly), stream independence."""
    x = synth.cfg3_surround(30.0)
    streams = np.stack([x * np.float32(1.0 - 0.02 * i) for i in range(16)]).astype(np.float32)
    plan = batch.LoudnessPlan(LoudnessConfig(), 8, capi.SURROUND, api=product.api)
    sa, nb = plan.execute_host(streams, 1024)
    A = batch.snapshots_to_arrays(sa, nb * 16)
    sh, nb2 = plan.execute_host((streams[3:5] * np.float32(0.5)).astype(np.float32), 1024)
    Hh = batch.snapshots_to_arrays(sh, nb * 2)
    assert nb2 == nb
    shift = np.float32(20.0 * np.log10(0.5))
    for k in A:
        a = A[k].reshape(16, nb, -1)[3:5].reshape(Hh[k].shape)
        live = (a > -99.0) & (Hh[k] > -99.0)  # away from the floors (-99.9 LUFS / -140 dB)
        assert live.mean() > 0.5, k
        assert np.max(np.abs(Hh[k][live] - (a[live] + shift))) <= 2e-5, (k, float(np.max(np.abs(Hh[k][live] - (a[live] + shift)))))
    # stream independence: the same two streams alone give the same snapshots bit for bit
    s2, _ = plan.execute_host(np.ascontiguousarray(streams[3:5]), 1024)
    B = batch.snapshots_to_arrays(s2, nb * 2)
    for k in A:
        assert np.array_equal(B[k], A[k].reshape(16, nb, -1)[3:5].reshape(B[k].shape)), k


def test_cfg4_properties_large(product):
    """cfg4 at BASELINE width (128 lanes x 5 s): the peak-hold trace dominates the unsmoothed one hop by hop and never falls
    faster than the configured decay; lane independence; x2 gain = +6.0206 dB on every bin above the floor."""
    L, S = 128, 240000
    base = synth.cfg4_streams(4, S / 48000.0).reshape(8, -1)[:, :S]
    lanes = np.concatenate([np.roll(base, 977 * r, axis=1) * np.float32(1.0 - 0.01 * r) for r in range(L // 8)], 0)
    hold_cfg = SpectrumConfig(fft_size=16384, hop_size=1024, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, floor_db=-100.0)
    none_cfg = SpectrumConfig(fft_size=16384, hop_size=1024, averaging=capi.AVG_NONE, floor_db=-100.0)
    hold = batch.SpectrumPlan(hold_cfg, api=product.api)
    wh, rh, pkh = hold.execute_host(lanes)
    sub = np.ascontiguousarray(lanes[17:20])
    wn, rn, _ = batch.SpectrumPlan(none_cfg, api=product.api).execute_host(sub)
    rs = rh[17:20]
    # hold = max(hold * decay, p) >= p — across two kernels (fused batch vs the two-kernel path), so up to the parity budget
    ph, pn = 10.0 ** (rs.astype(np.float64) / 10.0), 10.0 ** (rn.astype(np.float64) / 10.0)
    assert np.all(ph >= pn - 2e-5 * np.maximum(pn, pn.max(axis=-1, keepdims=True) * 1e-3))
    step = 12.0 * 1024.0 / 48000.0                               # dB per hop (spectrum/processor.rs:380-388)
    drop = rs[:, :-1] - rs[:, 1:]
    falling = rs[:, 1:] > -100.0
    assert np.max(drop[falling]) <= step + 1e-3
    # lane independence (the fused kernel serves the big batch, the two-kernel path the small one): same dB within the budget
    w3, r3, pk3 = hold.execute_host(sub)
    parity.compare_db(r3, rs, -100.0)
    parity.compare_db(w3, wh[17:20], -100.0)
    # gain
    w2, r2, _ = hold.execute_host((sub * np.float32(2.0)).astype(np.float32))
    live = (r3 > -90.0)
    assert np.max(np.abs(r2[live] - r3[live] - 6.0206)) <= 1e-3
