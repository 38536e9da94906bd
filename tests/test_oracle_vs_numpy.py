"""Second, independent pin of the C++ oracle: a float64 numpy restatement of the reassigned column.
Also documents the f32 noise the parity tolerances are calibrated on."""
import numpy as np

from openmeters_b200 import _capi as capi
from openmeters_b200 import synth
from openmeters_b200.processors import SpectrogramConfig
from oracle import oracle_py
from tests import ref_numpy


def test_oracle_reassigned_column_matches_float64_numpy():
    cfg = SpectrogramConfig(fft_size=4096, hop_size=1024, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
    lanes = synth.cfg2_lanes(3, 0.6)
    pts, cnt = oracle_py.stft_batch(cfg, lanes)
    checked = 0
    for l in range(3):
        for f in range(0, cnt.shape[1], 3):
            ref = ref_numpy.reassigned_column(lanes[l, f * 1024: f * 1024 + 8192], 4, 4096, 1024, 48000.0)
            keep = (ref["power"] >= 1e-14) & (ref["freq"] > 0) & (ref["freq"] < 24000.0)
            if keep.sum() != cnt[l, f]:
                # membership may differ only on threshold bins
                assert abs(int(keep.sum()) - int(cnt[l, f])) <= 2
                continue
            p = pts[l, f, : cnt[l, f]]
            peak = ref["power"].max()
            rel = ref["power"][keep] / peak
            widen = np.maximum(1.0, 1e-4 / rel)
            assert np.all(np.abs(p[:, 2] - ref["power"][keep]) <= 1e-5 * np.maximum(ref["power"][keep], peak * 1e-3))
            strong = rel >= 1e-6
            assert np.all(np.abs(p[strong, 0] - ref["time"][keep][strong]) <= 4e-5 * widen[strong])
            assert np.all(np.abs(p[strong, 1] - ref["freq"][keep][strong]) <= 0.24 * widen[strong])
            checked += int(strong.sum())
    assert checked > 20000


def test_oracle_windows_match_closed_forms():
    for kind in range(5):
        w = np.zeros(4096, np.float32)
        import ctypes as C
        oracle_py.api().window_coefficients(kind, 4096, w.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.max(np.abs(w - ref_numpy.window(kind, 4096))) < 3e-6
