#!/usr/bin/env python
"""One small launch of every specialised kernel, for `compute-sanitizer --tool racecheck|memcheck|synccheck`:

    compute-sanitizer --tool racecheck python tools/sanitizer_cases.py

The CPU emulator (tests/emu) runs threads cooperatively in a fixed order, so it cannot see shared-memory races; this does.
Inputs are a few frames per kernel so that the instrumented run stays within a couple of minutes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openmeters_b200 import _capi as capi  # noqa: E402
from openmeters_b200 import batch, synth  # noqa: E402
from openmeters_b200._lib import api as lib_api  # noqa: E402
from openmeters_b200.processors import LoudnessConfig, SpectrogramConfig, SpectrumConfig  # noqa: E402


def main():
    api = lib_api()
    assert api.device_count() >= 1
    api.set_device(0)
    done = []
    # reassigned: (N, hop) -> fast2 (aligned / small hop), 8k (aligned / small hop), 2k, 1k; frames chosen to leave a ragged tail group
    # N = 4096: generation 2 at the hops its ring handles with warp-uniform rows, generation 3 (stft_r64.cu: TMEM park, TMA frame fetch)
    # at every other hop.  `--only new` runs just the N = 4096 kernels (the environment picks the variant and is read once per
    # process: OMB_FAST_KERNEL=2|3, OMB_FAST2_BULK=7 = generation 2 with its tables in tensor memory, OMB_R64_PARK=global).
    only_new = "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "new"
    pin = os.environ.get("OMB_FAST_KERNEL")
    kind = {None: None, "1": 1, "2": 2, "3": 7}[pin]  # fast_kind of the pinned generation (stft.h)
    cases = [(4096, 1024, kind or 2), (4096, 64, kind or 2)]
    if pin != "2":
        cases.append((4096, 1000, 7))
    cases.append((16384, 4096, 8))  # stft_r64x.cu
    if not only_new:
        cases += [(8192, 2048, 4), (8192, 256, 8), (2048, 64, 5), (2048, 512, 5), (1024, 32, 6), (1024, 256, 6)]
    for n, hop, gen in cases:
        cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
        frames = 11
        S = 2 * n + (frames - 1) * hop
        lanes = synth.cfg2_lanes(2, (S + 64) / 48000.0)[:, :S]
        plan = batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=api)
        assert plan.kernel_generation == gen, (n, hop, plan.kernel_generation)
        pts, cnt = plan.execute_host(lanes)
        assert cnt.shape == (2, frames) and cnt.min() > 100
        done.append(f"reassigned {n}/{hop} gen {gen}")
    if only_new:
        print("\n".join(done))
        print("sanitizer cases ok")
        return
    # classic warp kernel + shared-memory tier
    for n, hop in [(1024, 512), (2048, 64)]:
        cfg = SpectrogramConfig(fft_size=n, hop_size=hop, window=capi.WINDOW_HANN, use_reassignment=False)
        S = n + 20 * hop
        codes = batch.StftPlan(cfg, api=api).execute_host(synth.cfg2_lanes(3, (S + 64) / 48000.0)[:, :S])
        assert codes.shape[1] == 21
        done.append(f"classic {n}/{hop}")
    # shared-memory reassigned tier
    cfg = SpectrogramConfig(fft_size=1024, hop_size=128, window=capi.WINDOW_HANN, use_reassignment=True, zero_padding_factor=2)
    S = 2048 + 5 * 128
    pts, cnt = batch.StftPlan(cfg, api=api).execute_host(synth.cfg2_lanes(2, (S + 64) / 48000.0)[:, :S])
    done.append("reassigned smem 1024 zp2")
    # residue decomposition of long zero-padded transforms (reassigned F = 16384 -> 2 residues; classic F = 32768 -> 4)
    cfg = SpectrogramConfig(fft_size=2048, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=True, zero_padding_factor=8)
    S = 4096 + 3 * 512
    pts, cnt = batch.StftPlan(cfg, api=api).execute_host(synth.cfg2_lanes(2, (S + 64) / 48000.0)[:, :S])
    assert cnt.shape == (2, 4) and cnt.min() > 1000
    cfg = SpectrogramConfig(fft_size=2048, hop_size=512, window=capi.WINDOW_HANN, use_reassignment=False, zero_padding_factor=16)
    S = 2048 + 3 * 512
    codes = batch.StftPlan(cfg, api=api).execute_host(synth.cfg2_lanes(2, (S + 64) / 48000.0)[:, :S])
    assert codes.shape == (2, 4, 16385)
    done.append("residue tiers")
    # spectrum: fused kernel (pinned on for few lanes) in the three modes, and the two-kernel path
    lanes = synth.cfg4_streams(2, (16384 + 4 * 1024) / 48000.0).reshape(4, -1)
    for mode, param in [(capi.AVG_PEAK_HOLD, 12.0), (capi.AVG_EXPONENTIAL, 0.7), (capi.AVG_NONE, 0.0)]:
        scfg = SpectrumConfig(fft_size=16384, hop_size=1024, averaging=mode, averaging_param=param, floor_db=-100.0)
        os.environ["OMB_SPECTRUM_FUSED"] = "1"
        batch.SpectrumPlan(scfg, api=api).execute_host_peaks(lanes)
        os.environ["OMB_SPECTRUM_FUSED"] = "0"
        batch.SpectrumPlan(scfg, api=api).execute_host_peaks(lanes)
        os.environ.pop("OMB_SPECTRUM_FUSED")
        done.append(f"spectrum mode {mode}")
    # loudness batch
    x = synth.cfg3_surround(0.4)
    batch.LoudnessPlan(LoudnessConfig(), 8, capi.SURROUND, api=api).execute_host(x[None, :], 1024)
    done.append("loudness batch")
    print("\n".join(done))
    print("sanitizer cases ok")


if __name__ == "__main__":
    main()
