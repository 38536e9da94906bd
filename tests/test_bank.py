"""Row f1, device side: the multi-stream spectrogram ring (omb_spectrogram_bank_*) must return, column for column, what
S independent SpectrogramProcessors return when fed the same blocks (here: the oracle's processors)."""
import numpy as np
import pytest

from openmeters_b200 import _capi as capi
from openmeters_b200 import synth
from openmeters_b200.meter import SpectrogramBank
from openmeters_b200.processors import AudioBlock, SpectrogramConfig
from tests import parity


def run_bank_vs_processors(api, oracle, cfg, S, seconds, channels, block_choices, seed, sr=48000.0):
    rng = np.random.default_rng(seed)
    if channels == 1:
        x = synth.cfg2_lanes(S, seconds, sr)                        # (S, n) mono
    else:
        base = synth.cfg1_stereo(seconds, sr).reshape(-1, 2)
        x = np.stack([(np.roll(base, 97 * s, axis=0) * np.float32(1 - 0.05 * s)).reshape(-1) for s in range(S)])  # (S, n*2)
    bank = SpectrogramBank(cfg, S, api=api)
    procs = [oracle.Spectrogram(cfg) for _ in range(S)]
    n_frames = x.shape[1] // channels
    o, total_cols, updates = 0, 0, 0
    classic_got = [[] for _ in range(S)]   # classic codes are judged over the whole run: the ">= 98 % exact" criterion is
    classic_want = [[] for _ in range(S)]  # statistical and means nothing on the handful of strong bins of one update
    while o < n_frames:
        nf = min(int(rng.choice(block_choices)), n_frames - o)
        blk = x[:, o * channels:(o + nf) * channels]
        got = bank.push(blk, channels, sr)
        want = [p.process_block(AudioBlock(blk[s], channels, sr)) for s, p in enumerate(procs)]
        assert (got is None) == all(w is None for w in want)
        if got is not None:
            updates += 1
            for s, w in enumerate(want):
                assert w is not None and len(got.columns[s]) == len(w.new_columns) and got.reset == w.reset
                if not w.new_columns:
                    continue
                if cfg.use_reassignment:
                    for a, b in zip(got.columns[s], w.new_columns):
                        parity.compare_reassigned_column(a, b, sr=sr, fft_len=cfg.fft_size * cfg.zero_padding_factor, window=cfg.fft_size, hop=cfg.hop_size)
                else:
                    classic_got[s] += got.columns[s]
                    classic_want[s] += w.new_columns
                total_cols += len(w.new_columns)
        o += nf
    if not cfg.use_reassignment:
        parity.compare_classic(np.stack([np.stack(c) for c in classic_got]), np.stack([np.stack(c) for c in classic_want]))
    return total_cols, updates


def test_bank_emulated_classic_and_reassigned(emu, oracle):
    """Kernel sources under the CPU emulator (development aid): small transforms, stereo fold-down and mono bypass,
    ragged block sizes (ring growth, relocation for alignment), history retention (history_length small -> skipped columns)."""
    cfg = SpectrogramConfig(fft_size=128, hop_size=32, window=capi.WINDOW_HANN, use_reassignment=False, history_length=6)
    cols, ups = run_bank_vs_processors(emu.api, oracle, cfg, 3, 0.1, 2, [17, 64, 250, 1000], seed=1)
    assert cols > 40 and ups > 5
    cfg = SpectrogramConfig(fft_size=64, hop_size=24, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True, history_length=64)
    cols, ups = run_bank_vs_processors(emu.api, oracle, cfg, 4, 0.05, 1, [5, 33, 200], seed=2)
    assert cols > 40 and ups > 5


def test_bank_emulated_hop_larger_than_window(emu, oracle):
    cfg = SpectrogramConfig(fft_size=64, hop_size=150, window=capi.WINDOW_HAMMING, use_reassignment=False, history_length=4)
    cols, ups = run_bank_vs_processors(emu.api, oracle, cfg, 2, 0.08, 2, [100, 256, 999], seed=3)
    assert cols > 10


@pytest.mark.gpu
def test_bank_gpu_cfg2_streams(product, oracle):
    """48 lock-step mono streams through the cfg2 kernel (4096-pt reassigned) with DspBatcher-sized blocks."""
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True, history_length=32)
    cols, ups = run_bank_vs_processors(product.api, oracle, cfg, 48, 0.9, 1, [256, 512, 1024, 1000], seed=4)
    assert cols > 48 * 20 and ups > 10


@pytest.mark.gpu
def test_bank_gpu_cfg1_stereo_streams(product, oracle):
    cfg = SpectrogramConfig(fft_size=1024, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False, history_length=64)
    cols, ups = run_bank_vs_processors(product.api, oracle, cfg, 20, 0.5, 2, [256, 768, 1024], seed=5)
    assert cols > 20 * 30


# ---------------------------------------------------------------- loudness bank
def run_loudness_bank_vs_processors(api, oracle, S, seconds, channels, positions, block_choices, seed, sr=48000.0):
    """omb_loudness_bank_push must return, snapshot for snapshot, what S independent LoudnessProcessors return."""
    from openmeters_b200.meter import LoudnessBank
    from openmeters_b200.processors import LoudnessConfig

    rng = np.random.default_rng(seed)
    if channels == 8:
        base = synth.cfg3_surround(seconds, sr)
    else:
        base = synth.cfg1_stereo(seconds, sr)
    x = np.stack([(np.roll(base.reshape(-1, channels), 131 * s, axis=0) * np.float32(1 - 0.07 * s)).reshape(-1) for s in range(S)])
    x[S - 1, : int(0.05 * sr) * channels] = 0.0                       # one stream starts silent: lazy activation (processor.rs:264-274)
    bank = LoudnessBank(LoudnessConfig(sample_rate=sr), S, api=api)
    procs = [oracle.Loudness(LoudnessConfig(sample_rate=sr)) for _ in range(S)]
    n_frames = x.shape[1] // channels
    o, n_snap = 0, 0
    worst = 0.0
    while o < n_frames:
        nf = min(int(rng.choice(block_choices)), n_frames - o)
        blk = np.ascontiguousarray(x[:, o * channels:(o + nf) * channels])
        got = bank.push(blk, channels, sr, positions)
        want = [p.process_block(AudioBlock(blk[s], channels, sr, positions)) for s, p in enumerate(procs)]
        assert got is not None and all(w is not None for w in want)
        for s, w in enumerate(want):
            g = got[s]
            assert g.channel_count == channels
            for a, b in ((g.short_term_loudness, w.short_term_loudness), (g.momentary_loudness, w.momentary_loudness)):
                worst = max(worst, abs(a - b))
            for name in ("rms_fast_db", "rms_slow_db", "true_peak_db"):
                ga = np.array(getattr(g, name)[:channels], np.float32)
                wa = np.asarray(getattr(w, name), np.float32)[:channels]
                tol = 1e-5 if name == "true_peak_db" else 5e-5
                assert np.max(np.abs(ga - wa)) <= tol, (name, s, o, ga, wa)
            n_snap += 1
        o += nf
    assert worst <= 5e-5, worst
    return n_snap


def test_loudness_bank_emulated(emu, oracle):
    n = run_loudness_bank_vs_processors(emu.api, oracle, S=3, seconds=0.35, channels=8, positions=capi.SURROUND,
                                        block_choices=(256, 512, 1024, 700), seed=5)
    assert n >= 3 * 16
    run_loudness_bank_vs_processors(emu.api, oracle, S=2, seconds=0.2, channels=2, positions=None, block_choices=(480, 1000), seed=6, sr=96000.0)


def test_loudness_bank_reset_and_single_stream_equivalence(emu):
    from openmeters_b200.meter import LoudnessBank
    from openmeters_b200.processors import LoudnessConfig, LoudnessProcessor

    x = synth.cfg1_stereo(0.1)
    blocks = np.stack([x, x * np.float32(0.5)])
    bank = LoudnessBank(LoudnessConfig(), 2, api=emu.api)
    a = bank.push(blocks[:, :2048], 2, 48000.0)
    m0 = (a[0].momentary_loudness, a[1].momentary_loudness)
    bank.reset_audio()
    b = bank.push(blocks[:, :2048], 2, 48000.0)
    assert (b[0].momentary_loudness, b[1].momentary_loudness) == m0          # reset restores the initial state of every stream
    single = LoudnessProcessor(LoudnessConfig(), api=emu.api).process_block(AudioBlock(blocks[1, :2048], 2, 48000.0))
    assert single.momentary_loudness == b[1].momentary_loudness             # same kernel, same bits


@pytest.mark.gpu
def test_loudness_bank_gpu(product, oracle):
    n = run_loudness_bank_vs_processors(product.api, oracle, S=16, seconds=1.5, channels=8, positions=capi.SURROUND,
                                        block_choices=(256, 512, 768, 1024), seed=7)
    assert n >= 16 * 60


# ---------------------------------------------------------------- spectrum bank
def run_spectrum_bank_vs_processors(api, oracle, cfg, S, seconds, channels, block_choices, seed, sr=48000.0):
    """omb_spectrum_bank_push must return, trace for trace, what S independent SpectrumProcessors return."""
    from openmeters_b200.meter import SpectrumBank

    rng = np.random.default_rng(seed)
    base = synth.cfg4_streams(S, seconds, sr)                         # (S, 2, n)
    if channels == 2:
        x = np.stack([np.stack([base[s, 0], base[s, 1]], 1).reshape(-1) for s in range(S)]).astype(np.float32)
    else:
        x = np.ascontiguousarray(base[:, 0, :], np.float32)
    bank = SpectrumBank(cfg, S, api=api)
    procs = [oracle.Spectrum(cfg) for _ in range(S)]
    n_frames = x.shape[1] // channels
    o, n_snap = 0, 0
    while o < n_frames:
        nf = min(int(rng.choice(block_choices)), n_frames - o)
        blk = np.ascontiguousarray(x[:, o * channels:(o + nf) * channels])
        got = bank.push(blk, channels, sr)
        want = [p.process_block(AudioBlock(blk[s], channels, sr)) for s, p in enumerate(procs)]
        assert (got is None) == all(w is None for w in want), (o, nf)
        if got is not None:
            tidx, freqs, w, r = got
            for s, snap in enumerate(want):
                assert snap is not None
                assert np.array_equal(freqs, snap.frequency_bins)
                for i, t in enumerate(tidx):
                    parity.compare_db(r[s, i][None], np.asarray(snap.traces[t][1])[None], cfg.floor_db)
                    parity.compare_db(w[s, i][None], np.asarray(snap.traces[t][0])[None], cfg.floor_db)
            n_snap += 1
        o += nf
    return n_snap


@pytest.mark.parametrize("mode,param", [(capi.AVG_PEAK_HOLD, 12.0), (capi.AVG_EXPONENTIAL, 0.6), (capi.AVG_NONE, 0.0)])
def test_spectrum_bank_emulated(emu, oracle, mode, param):
    from openmeters_b200.processors import SpectrumConfig

    cfg = SpectrumConfig(fft_size=512, hop_size=128, averaging=mode, averaging_param=param, source=capi.CHANNEL_LEFT,
                         secondary_source=capi.CHANNEL_SIDE, floor_db=-100.0)
    n = run_spectrum_bank_vs_processors(emu.api, oracle, cfg, S=3, seconds=0.12, channels=2, block_choices=(100, 256, 700, 1024), seed=3)
    assert n >= 5


def test_spectrum_bank_single_trace_and_large_hop(emu, oracle):
    from openmeters_b200.processors import SpectrumConfig

    # only the secondary source is active (processor.rs:174-177); hop larger than the window exercises the skip accounting
    cfg = SpectrumConfig(fft_size=256, hop_size=600, averaging=capi.AVG_PEAK_HOLD, averaging_param=6.0, source=capi.CHANNEL_NONE,
                         secondary_source=capi.CHANNEL_MID)
    n = run_spectrum_bank_vs_processors(emu.api, oracle, cfg, S=2, seconds=0.15, channels=2, block_choices=(256, 512, 999), seed=4)
    assert n >= 4
    cfg = SpectrumConfig(fft_size=256, hop_size=64, source=capi.CHANNEL_MID)   # mono input, default averaging
    run_spectrum_bank_vs_processors(emu.api, oracle, cfg, S=2, seconds=0.08, channels=1, block_choices=(333, 512), seed=5)


@pytest.mark.gpu
def test_spectrum_bank_gpu(product, oracle):
    from openmeters_b200.processors import SpectrumConfig

    cfg = SpectrumConfig(fft_size=16384, hop_size=1024, averaging=capi.AVG_PEAK_HOLD, averaging_param=12.0, source=capi.CHANNEL_LEFT,
                         secondary_source=capi.CHANNEL_RIGHT, floor_db=-100.0)
    n = run_spectrum_bank_vs_processors(product.api, oracle, cfg, S=8, seconds=0.8, channels=2, block_choices=(512, 1024, 2048), seed=8)
    assert n >= 10


# ------------------------------------------------------------------ streaming loudness: phase-parallel kernel vs the sequential port
def _loudness_stream_sequences_equal(api, monkeypatch, sr, channels, sizes, seed):
    """k_loudness_stream (IIR / true peak / window sums as separate phases) must give the SAME BYTES as k_loudness_stream_seq
    (the reference's per-sample loop, OMB_LOUDNESS_STREAM_SEQ=1) snapshot after snapshot: ragged blocks, a channel that
    starts silent and wakes up mid-block, blocks longer than the shortest windows, non-finite samples."""
    import ctypes as C

    from openmeters_b200.processors import AudioBlock, LoudnessConfig, LoudnessProcessor

    rng = np.random.default_rng(seed)
    total = sum(sizes)
    x = (0.4 * np.sin(np.arange(total)[:, None] * (0.01 + 0.003 * np.arange(channels))[None, :]) +
         0.05 * rng.uniform(-1, 1, (total, channels))).astype(np.float32)
    if channels > 1:
        x[: sizes[0] + sizes[1] // 2, 1] = 0.0          # channel 1 wakes up in the middle of the second block
    x[total // 2, 0] = np.float32(np.inf)               # y^2 non-finite -> pushed as 0 (dsp.rs:324-333); |s| = inf is a legal peak
    x[total // 2 + 3, channels - 1] = np.float32(np.nan)
    snaps = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("OMB_LOUDNESS_STREAM_SEQ", mode)
        p = LoudnessProcessor(LoudnessConfig(sample_rate=sr), api=api)
        out, pos = [], 0
        for n in sizes:
            s = p.process_block(AudioBlock(x[pos:pos + n].reshape(-1), channels, sr))
            pos += n
            out.append((np.float32(s.short_term_loudness), np.float32(s.momentary_loudness), np.array(s.rms_fast_db, np.float32),
                        np.array(s.rms_slow_db, np.float32), np.array(s.true_peak_db, np.float32)))
        snaps[mode] = out
    for a, b in zip(snaps["0"], snaps["1"]):
        for u, v in zip(a, b):
            assert np.array_equal(np.asarray(u).view(np.uint32), np.asarray(v).view(np.uint32)), (u, v)


@pytest.mark.parametrize("sr,channels,sizes", [(8000.0, 2, [5, 700, 1, 3000, 64, 2500]), (48000.0, 3, [256, 999, 17, 1024]),
                                               (96000.0, 1, [300, 40, 1000]), (192000.0, 2, [100, 333])])
def test_loudness_stream_kernels_agree_emulated(emu, monkeypatch, sr, channels, sizes):
    _loudness_stream_sequences_equal(emu.api, monkeypatch, sr, channels, sizes, 11)


@pytest.mark.gpu
@pytest.mark.parametrize("sr,channels,sizes", [(8000.0, 2, [5, 700, 1, 30000, 64, 2500]), (48000.0, 8, [256, 9600, 17, 1024, 200000, 3]),
                                               (44100.0, 6, [1024] * 12), (96000.0, 2, [300, 40, 100000]), (192000.0, 2, [100, 33333])])
def test_loudness_stream_kernels_agree_gpu(product, monkeypatch, sr, channels, sizes):
    _loudness_stream_sequences_equal(product.api, monkeypatch, sr, channels, sizes, 12)
