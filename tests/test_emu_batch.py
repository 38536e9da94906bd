"""Batched entry points of the product's kernel sources (under the CPU emulator) vs the oracle, on small
seeded inputs.  Development aid; the -m gpu suite repeats these at full size on the real build."""
import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200 import batch, synth
from openmeters_b200.processors import AudioBlock, LoudnessConfig, SpectrogramConfig, SpectrumConfig
from oracle import oracle_py
from tests import cases, parity


@pytest.mark.parametrize("window,n,hop,zp", [(capi.WINDOW_HANN, 64, 16, 1), (capi.WINDOW_BLACKMAN_HARRIS, 256, 64, 2),
                                             (capi.WINDOW_HAMMING, 128, 200, 1)])
def test_classic_batch(emu, window, n, hop, zp):
    cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=window, use_reassignment=False, zero_padding_factor=zp)
    lanes = synth.cfg2_lanes(3, 0.05)
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_GENERIC)
    assert st["exact"] >= 0.98
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_AUTO)  # shared-memory Stockham kernel
    assert st["exact"] >= 0.98


@pytest.mark.parametrize("window,n,hop,zp", [(capi.WINDOW_HANN, 64, 16, 1), (capi.WINDOW_BLACKMAN_HARRIS, 256, 64, 1),
                                             (capi.WINDOW_BLACKMAN, 128, 32, 4)])
def test_reassigned_batch(emu, window, n, hop, zp):
    cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=window, use_reassignment=True, zero_padding_factor=zp)
    lanes = synth.cfg2_lanes(2, 0.04)
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_GENERIC)
    assert st["checked"] > 100
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_AUTO)  # shared-memory Stockham kernel
    assert st["checked"] > 100


@pytest.mark.parametrize("mode,param", [(capi.AVG_NONE, 0.0), (capi.AVG_EXPONENTIAL, 0.5), (capi.AVG_PEAK_HOLD, 12.0)])
def test_spectrum_batch(emu, mode, param):
    cfg = SpectrumConfig(fft_size=256, hop_size=64, averaging=mode, averaging_param=param, floor_db=-100.0)
    lanes = synth.cfg4_streams(2, 0.05)[:, 0, :]
    cases.spectrum_parity(emu.api, cfg, lanes)


def test_loudness_batch(emu):
    x = synth.cfg3_surround(0.62)  # 29760 frames: crosses the 0.5 s lazy-activation edge and two window lengths
    st = cases.loudness_parity(emu.api, LoudnessConfig(), 8, capi.SURROUND, x[None, :], 1024)
    assert st["short_term"] < 5e-5


def test_loudness_batch_two_streams_stereo(emu):
    a = synth.cfg1_stereo(0.2)
    b = (a * np.float32(0.25)).astype(np.float32)
    cases.loudness_parity(emu.api, LoudnessConfig(), 2, None, np.stack([a, b]), 500)


@pytest.mark.parametrize("gen", ["1", "2", "3"])
def test_specialised_reassigned_kernels(emu, gen, monkeypatch):
    """All three generations of the specialised N=4096 kernel (stft_fast.cu / stft_fast2.cu / stft_r64.cu) under the emulator:
    multi-run work split, odd frame counts (one idle group in the last pair), 3 lanes."""
    monkeypatch.setenv("OMB_FAST_KERNEL", gen)
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    lanes = synth.cfg2_lanes(3, (8192 + 40 * 1024) / 48000.0)  # 41 frames per lane
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 3 * 41 and st["unmatched"] <= 4


def test_specialised_kernel_contiguous_ranges(emu, monkeypatch):
    """Generation 2 at a batch large enough for the contiguous work assignment (>= 96 frames per SM; the emulator has 4): every CTA
    walks one range of the linearised (lane, frame) sequence; the ranges cut lanes at odd frames and span a lane boundary (three
    lanes of 131 frames over 4 CTAs of 100).  Must equal the round-robin runs bit for bit, and match the oracle."""
    monkeypatch.setenv("OMB_FAST_KERNEL", "2")
    cfg = SpectrogramConfig(fft_size=4096, hop_size=512, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    n = 8192 + 130 * 512
    lanes = synth.cfg2_lanes(3, (n + 64) / 48000.0)[:, :n]
    pa, ca = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=emu.api).execute_host(lanes)
    pb, cb = oracle_py.stft_batch(cfg, lanes)
    st = parity.compare_reassigned(pa, ca, pb, cb, sr=48000.0, fft_len=4096, window=4096, hop=512)
    assert st["cols"] == 3 * 131


@pytest.mark.parametrize("n,hop,frames,sr", [(8192, 512, 101, 96000.0), (2048, 64, 403, 48000.0), (1024, 32, 805, 48000.0)])
def test_ring_kernels_contiguous_ranges(emu, n, hop, frames, sr):
    """stft_fast8k.cu / stft_fast2k.cu / stft_fast1k.cu at batches large enough for the contiguous work assignment (the emulator has 4
    SMs): ranges that cut lanes at frames that are not multiples of the kernels' frames-per-iteration and span the lane boundary."""
    cfg = SpectrogramConfig(sample_rate=sr, fft_size=n, hop_size=hop, window=capi.WINDOW_HANN, use_reassignment=True)
    S = 2 * n + (frames - 1) * hop
    lanes = synth.cfg2_lanes(2, (S + 64) / 48000.0)[:, :S]
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 2 * frames


def test_specialised_kernel_other_hops_and_windows(emu):
    for hop, win in ((512, capi.WINDOW_HANN), (2048, capi.WINDOW_BLACKMAN)):
        cfg = SpectrogramConfig(fft_size=4096, hop_size=hop, window=win, use_reassignment=True)
        lanes = synth.cfg2_lanes(1, (8192 + 6 * hop) / 48000.0)
        cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)


@pytest.mark.parametrize("hop", [256, 128, 64, 32, 100])
def test_specialised_kernel_small_hops(emu, hop):
    """The settings UI's N/16 ... N/128 hops (and any other multiple of 4 below 512) at N = 4096: frames whose ring origin is
    not 512-aligned; enough frames that a run walks past the end of the 16384-sample ring (frames x hop + 8192 > 16384) and
    splits into several runs; odd frame count."""
    cfg = SpectrogramConfig(fft_size=4096, hop_size=hop, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    frames = max(45, 9000 // hop) | 1
    n = 8192 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(2, (n + 64) / 48000.0)[:, :n]
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 2 * frames and st["unmatched"] <= 2e-3 * st["pts"] + 2


@pytest.mark.parametrize("hop", [256, 64])
def test_8k_kernel_small_hops(emu, hop):
    """stft_fast8k.cu at hops that are not multiples of 512: per-thread ring wrap (ring = 16384 + hop samples, so a run
    of more than 1 + 16384 / hop ... frames wraps; 70 / 270 frames here)."""
    cfg = SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=hop, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    frames = (17000 // hop + 4) | 1
    n = 16384 + (frames - 1) * hop
    lanes = synth.cfg5_lanes(1, n)[:, :n]
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == frames


@pytest.mark.parametrize("hop,frames,window", [(512, 9, capi.WINDOW_HANN), (64, 150, capi.WINDOW_BLACKMAN_HARRIS), (100, 45, capi.WINDOW_HAMMING),
                                               (1024, 6, capi.WINDOW_HANN), (16, 300, capi.WINDOW_HANN)])
def test_2k_kernel_two_frames_per_transform(emu, hop, frames, window):
    """stft_fast2k.cu (N = 2048, the product's default size): two interleaved frames per 4096-point transform, separated in
    registers.  Frame counts of every residue mod 4 (missing second frame / idle second group at the tail of a run), ring
    wrap-around (ring 8192 or 16384 samples), hops that are not powers of two."""
    cfg = SpectrogramConfig(fft_size=2048, hop_size=hop, window=window, use_reassignment=True)
    n = 4096 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(2, (n + 64) / 48000.0)[:, :n]
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 2 * frames and st["unmatched"] <= 2e-3 * st["pts"] + 2


@pytest.mark.parametrize("hop,frames,window", [(256, 17, capi.WINDOW_HANN), (32, 203, capi.WINDOW_BLACKMAN_HARRIS), (100, 45, capi.WINDOW_HAMMING),
                                               (512, 12, capi.WINDOW_HANN), (8, 301, capi.WINDOW_BLACKMAN), (32, 6, capi.WINDOW_HANN)])
def test_1k_kernel_four_frames_per_transform(emu, hop, frames, window):
    """stft_fast1k.cu (N = 1024): four interleaved frames per 4096-point transform, separated by a radix-4 butterfly in
    registers.  Frame counts of several residues mod 8 (zero-fed missing frames, idle second group), ring wrap-around."""
    cfg = SpectrogramConfig(fft_size=1024, hop_size=hop, window=window, use_reassignment=True)
    n = 2048 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(2, (n + 64) / 48000.0)[:, :n]
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=emu.api)
    assert plan.kernel_generation == 6
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 2 * frames and st["unmatched"] <= 2e-3 * st["pts"] + 2


def test_2k_kernel_single_frame_and_silence(emu):
    cfg = SpectrogramConfig(fft_size=2048, hop_size=64, window=capi.WINDOW_HANN, use_reassignment=True)
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=emu.api)
    assert plan.kernel_generation == 5
    lanes = np.zeros((3, 4096), np.float32)
    lanes[1] = 0.25
    lanes[2] = synth.cfg2_lanes(1, 4096 / 48000.0)[0, :4096]
    pts, cnt = plan.execute_host(lanes)
    assert cnt.shape == (3, 1) and cnt[0, 0] == 0 and cnt[1, 0] == 0 and cnt[2, 0] > 500  # silence, DC -> empty columns
    assert np.all(np.isfinite(pts[2, 0, :cnt[2, 0]]))


def test_host_path_pipelined_lane_chunks(emu):
    """execute_host pipelines lane chunks over three streams when a specialised kernel is active (>= 4 lanes)."""
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    lanes = synth.cfg2_lanes(5, (8192 + 2 * 1024) / 48000.0)
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 15


def test_spectrum_fast_16k(emu):
    """spectrum_fast.cu (N = 16384: two parallel 4096-point FFTs + fused combine / real split) vs the oracle."""
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=capi.AVG_PEAK_HOLD,
                         averaging_param=12.0, floor_db=-100.0)
    lanes = synth.cfg4_streams(1, (16384 + 2 * 1024) / 48000.0).reshape(2, -1)
    cases.spectrum_parity(emu.api, cfg, lanes)


def test_streaming_survives_misaligned_fifo(emu):
    """After an odd-sized drain (config change) the device FIFO is no longer 16-byte aligned: the specialised
    N=4096 kernel must hand over to the shared-memory tier instead of failing, with identical columns."""
    from openmeters_b200.processors import AudioBlock, SpectrogramProcessor
    from oracle import oracle_py

    lane = synth.cfg2_lanes(1, 1.0)[0]
    outs = {}
    for name, api in (("emu", emu.api), ("oracle", oracle_py.api())):
        p = SpectrogramProcessor(SpectrogramConfig(fft_size=256, hop_size=65, window=capi.WINDOW_BLACKMAN_HARRIS,
                                                   use_reassignment=True, history_length=64), api=api)
        p.process_block(AudioBlock(lane[:1001], 1, 48000.0))          # 12 frames * 65 = 780... +65 -> odd FIFO offset below
        p.process_block(AudioBlock(lane[1001:1101], 1, 48000.0))      # 13th frame: FIFO begins at sample 845 (odd)
        c = p.config()
        c.fft_size, c.hop_size = 4096, 1024                            # rebuild: keeps the newest 2*H samples (odd offset)
        p.update_config(c)
        cols = []
        for s in range(1101, 1101 + 12 * 1024 + 9000, 1024):
            up = p.process_block(AudioBlock(lane[s:s + 1024], 1, 48000.0))
            if up is not None:
                cols.extend(up.new_columns)
        outs[name] = cols
    assert len(outs["emu"]) == len(outs["oracle"]) > 3
    for a, b in zip(outs["emu"], outs["oracle"]):
        parity.compare_reassigned_column(a, b, sr=48000.0, fft_len=4096, window=4096, hop=1024)


def test_host_path_ragged_lane_length(emu):
    """Lane lengths that are not a multiple of 4 samples: the host path pads the device stride so every lane stays
    16-byte aligned for the specialised kernel (KERNEL_FAST would fail loudly otherwise)."""
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    lanes = np.ascontiguousarray(synth.cfg2_lanes(5, (8192 + 2 * 1024 + 4) / 48000.0)[:, : 8192 + 2 * 1024 + 3])
    assert lanes.shape[1] % 4 == 3
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 15
    st = cases.stft_parity(emu.api, cfg, lanes[:2], kernel=capi.KERNEL_FAST, expect_fast=True)  # non-pipelined branch
    assert st["cols"] == 6


def test_classic_warp_kernel_1024(emu):
    """stft_classic_fast.cu (cfg1: N = 1024 classic, one warp per frame) vs the oracle; several hops, a DC offset
    (exercises the mean removal) and more frames than one pass of the persistent grid covers per warp."""
    for hop, win in ((512, capi.WINDOW_HANN), (256, capi.WINDOW_BLACKMAN_HARRIS), (1536, capi.WINDOW_HAMMING)):
        cfg = SpectrogramConfig(fft_size=1024, hop_size=hop, window=win, use_reassignment=False)
        lanes = synth.cfg1_stereo(0.4).reshape(-1, 2).T if hop == 512 else synth.cfg2_lanes(3, 0.25) + np.float32(0.125)
        st = cases.stft_parity(emu.api, cfg, np.ascontiguousarray(lanes, np.float32), kernel=capi.KERNEL_FAST, expect_fast=True)
        assert st["exact"] >= 0.98, st


def test_specialised_reassigned_kernel_8192(emu):
    """stft_fast8k.cu (cfg5: N = 8192, two 4096-point sub-transforms per CTA) vs the oracle: several hops, a run
    split across CTAs, ring wrap-around (more frames than the ring holds hops)."""
    for hop, win, frames, nl in ((2048, capi.WINDOW_BLACKMAN_HARRIS, 21, 2), (512, capi.WINDOW_HANN, 5, 1), (1536, capi.WINDOW_BLACKMAN, 13, 1)):
        cfg = SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=hop, window=win, use_reassignment=True)
        lanes = synth.cfg5_lanes(nl, 16384 + (frames - 1) * hop)
        st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
        assert st["cols"] == nl * frames and st["unmatched"] <= 4, st


@pytest.mark.parametrize("mode,param", [(capi.AVG_PEAK_HOLD, 12.0), (capi.AVG_EXPONENTIAL, 0.6), (capi.AVG_NONE, 0.0)])
def test_spectrum_fused_16k(emu, mode, param, monkeypatch):
    """k_spectrum_fused_16k (FFT + smoothing + dB + arg-max in one kernel, one CTA per lane; pinned on with
    OMB_SPECTRUM_FUSED=1 because two lanes would not select it) vs the oracle: ring wrap-around across 20 hops."""
    monkeypatch.setenv("OMB_SPECTRUM_FUSED", "1")
    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, window=capi.WINDOW_HANN, averaging=mode, averaging_param=param, floor_db=-100.0)
    lanes = synth.cfg4_streams(1, (16384 + 19 * 1024) / 48000.0).reshape(2, -1)
    st = cases.spectrum_parity(emu.api, cfg, lanes)
    assert st is not None


@pytest.mark.parametrize("mode,param,hop,lanes", [(capi.AVG_PEAK_HOLD, 12.0, 1024, 2), (capi.AVG_EXPONENTIAL, 0.6, 2048, 3), (capi.AVG_NONE, 0.0, 512, 1)])
def test_spectrum_two_kernel_path_ring_power(emu, mode, param, hop, lanes, monkeypatch):
    """The two-kernel path (few lanes: OMB_SPECTRUM_FUSED=0) with the front half of the fused kernel as its power stage: (lane, hop
    segment) work items, each priming its own ring — 23 hops split into segments of 8, so segments start mid-lane — against the
    oracle, and bit-for-bit against the same kernel run as ONE segment per lane is not observable from here, so: against the
    frame-per-CTA power kernel (OMB_SPECTRUM_RING_POWER=0) within the parity metric."""
    monkeypatch.setenv("OMB_SPECTRUM_FUSED", "0")
    cfg = SpectrumConfig(fft_size=16384, hop_size=hop, window=capi.WINDOW_HANN, averaging=mode, averaging_param=param, floor_db=-100.0)
    x = synth.cfg4_streams(2, (16384 + 22 * hop) / 48000.0).reshape(4, -1)[:lanes]
    monkeypatch.setenv("OMB_SPECTRUM_RING_POWER", "1")
    cases.spectrum_parity(emu.api, cfg, x)
    wa, ra, pka = batch.SpectrumPlan(cfg, api=emu.api).execute_host(x)
    monkeypatch.setenv("OMB_SPECTRUM_RING_POWER", "0")
    wb, rb, pkb = batch.SpectrumPlan(cfg, api=emu.api).execute_host(x)
    assert wa.shape == wb.shape and np.max(np.abs(wa - wb)) < 2e-3 and np.max(np.abs(ra - rb)) < 2e-3  # dB; both within 1e-5 of the oracle in power


@pytest.mark.parametrize("mode,param,hop", [(capi.AVG_PEAK_HOLD, 12.0, 1024), (capi.AVG_EXPONENTIAL, 0.6, 2048), (capi.AVG_NONE, 0.0, 512)])
def test_spectrum_fused_16k_planar_ring(emu, mode, param, hop, monkeypatch):
    """The experimental planar staging ring of the fused kernel (OMB_SPECTRUM_PLANAR=1, off by default): even / odd float2 of
    every quad in two planes filled by 8-byte async copies, so each transform group reads at unit stride.  Same outputs."""
    monkeypatch.setenv("OMB_SPECTRUM_FUSED", "1")
    monkeypatch.setenv("OMB_SPECTRUM_PLANAR", "1")
    cfg = SpectrumConfig(fft_size=16384, hop_size=hop, window=capi.WINDOW_HANN, averaging=mode, averaging_param=param, floor_db=-100.0)
    lanes = synth.cfg4_streams(1, (16384 + 19 * hop) / 48000.0).reshape(2, -1)
    plan = batch.SpectrumPlan(cfg, api=emu.api)
    wa, ra, pka = plan.execute_host(lanes)
    monkeypatch.setenv("OMB_SPECTRUM_PLANAR", "0")
    wb, rb, pkb = batch.SpectrumPlan(cfg, api=emu.api).execute_host(lanes)
    assert np.array_equal(wa, wb) and np.array_equal(ra, rb) and np.array_equal(pka, pkb)  # same arithmetic, different staging
    monkeypatch.setenv("OMB_SPECTRUM_PLANAR", "1")
    cases.spectrum_parity(emu.api, cfg, lanes)


@pytest.mark.parametrize("n,hop,zp,window,reassign", cases.settings_grid())
def test_settings_grid_emulated(emu, n, hop, zp, window, reassign):
    """the (size, hop, zero padding, window, mode) grid the settings UI can reach, through whichever kernel
    tier the plan picks — caught a 341-thread launch (partial warp under full-mask collectives) for F = 4096."""
    cases.settings_grid_case(emu.api, n, hop, zp, window, reassign)

@pytest.mark.parametrize("hop,frames,window", [(4096, 5, capi.WINDOW_BLACKMAN_HARRIS), (256, 7, capi.WINDOW_HANN), (1000, 3, capi.WINDOW_HAMMING)])
def test_16k_kernel_on_chip(emu, hop, frames, window):
    """stft_r64x.cu (N = 16384 reassigned: one CTA per frame, 64 x 64 x 4 transforms, TMEM park, bin-staged ordered column) vs the
    oracle: more frames than emulated SMs (a CTA walks several frames: TMA phase bookkeeping), two lanes, hops off every grid."""
    cfg = SpectrogramConfig(fft_size=16384, hop_size=hop, window=window, use_reassignment=True)
    n = 32768 + (frames - 1) * hop
    lanes = synth.cfg2_lanes(2, (n + 64) / 48000.0)[:, :n]
    st = cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
    assert st["cols"] == 2 * frames and st["checked"] > 10000

@pytest.mark.parametrize("n,hop,sr", [(4096, 1000, 48000.0), (16384, 4096, 48000.0), (8192, 1000, 96000.0)])
def test_team_kernels_streaming_matches_batch(emu, n, hop, sr):
    """The streaming processor over the team kernels (stft_r64.cu / stft_r64x.cu: hops off the ring kernels' grid): ragged blocks, so
    calls start at first_frame > 0 and at FIFO offsets that are multiples of 4 floats only when the hop is; every column equals the
    batch path's (same kernel, same frame) bit for bit, and the batch path matches the oracle."""
    cfg = SpectrogramConfig(sample_rate=sr, fft_size=n, hop_size=hop, window=capi.WINDOW_HANN, use_reassignment=True, history_length=64)
    frames = 7
    S = 2 * n + (frames - 1) * hop
    lanes = synth.cfg2_lanes(1, (S + 64) / 48000.0)[:, :S]
    plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=emu.api)
    assert plan.kernel_generation in (7, 8)
    pts, cnt = plan.execute_host(lanes)
    p = emu.Spectrogram(cfg)
    cols = []
    step = 4 * 997  # a multiple of 4: the device FIFO stays 16-byte aligned, so the specialised kernel keeps serving the stream
    for s0 in range(0, S, step):
        up = p.process_block(AudioBlock(lanes[0, s0:min(s0 + step, S)], 1, sr))
        if up is not None:
            cols += list(up.new_columns)
    assert len(cols) == frames
    for f, c in enumerate(cols):
        assert np.array_equal(np.asarray(c), pts[0, f, :cnt[0, f]]), f
    cases.stft_parity(emu.api, cfg, lanes, kernel=capi.KERNEL_FAST, expect_fast=True)
