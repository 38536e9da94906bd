"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the
entry points include/omb200.h declares (no compute calls — there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

from openmeters_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "omb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(omb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from openmeters_b200 import build

    path = build.build()
    return C.CDLL(path, mode=C.RTLD_LOCAL)


def test_header_and_ctypes_table_agree():
    assert sorted("omb_" + k for k in capi.HEADER_SYMBOLS) == header_functions()


def test_library_exports_every_declared_symbol(lib):
    missing = [f for f in header_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_library_is_built_for_sm_100a():
    import subprocess

    from openmeters_b200 import _lib

    out = subprocess.run(["cuobjdump", "--list-elf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_compute_without_gpu_fails_loudly(lib):
    """No CPU fallback: on a GPU-less host every compute entry point reports OMB_ERR_CUDA."""
    api = capi.bind(lib, "omb_")
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    cfg = capi.SpectrogramConfig()
    api.spectrogram_default_config(C.byref(cfg))
    h = C.c_void_p()
    assert api.spectrogram_create(C.byref(cfg), C.byref(h)) == 0  # creation never rejects (processor.rs:71-82)
    assert api.spectrogram_prepare(h) == capi.ERR_CUDA
    assert b"no CPU fallback" in api.last_error()
    plan = C.c_void_p()
    assert api.stft_plan_create(C.byref(cfg), 0, C.byref(plan)) == capi.ERR_CUDA
    api.spectrogram_destroy(h)


def test_host_side_tables_match_oracle(lib, oracle):
    """Plan set-up runs on the host in the product too: pin it against the oracle bit for bit where the
    arithmetic is identical (windows, norms, K-weighting, FIRs, layouts) and tightly where an FFT is involved."""
    import numpy as np

    from tests.backends import _utils

    p, o = _utils(capi.bind(lib, "omb_")), oracle.u
    for kind in range(5):
        for n in (1, 2, 8, 64, 1024, 4096):
            assert np.array_equal(p.window(kind, n), o.window(kind, n)), (kind, n)
    w = o.window(capi.WINDOW_BLACKMAN_HARRIS, 4096)
    assert np.array_equal(p.bin_norm(w, 8192), o.bin_norm(w, 8192))
    dp, tp = p.reassignment_windows(w)
    do, to = o.reassignment_windows(w)
    assert np.array_equal(tp, to)
    assert np.max(np.abs(dp - do)) <= 2e-7 * np.max(np.abs(do))
    assert p.power_scale(w, 4096) == o.power_scale(w, 4096)
    for sr in (44100.0, 48000.0, 96000.0, 192000.0):
        assert all(np.array_equal(x, y) for x, y in zip(p.k_weighting(sr), o.k_weighting(sr)))
    assert np.array_equal(p.true_peak_fir(4), o.true_peak_fir(4)) and np.array_equal(p.true_peak_fir(2), o.true_peak_fir(2))
    for f in (0.0, 1.0, 31.5, 1000.0, 20000.0):
        assert p.a_weight(f) == o.a_weight(f)
    for db in (-150.0, -140.0, -23.456, 0.0, 13.0):
        assert p.pack_classic_db(db) == o.pack_classic_db(db)
    for ch in range(1, 9):
        assert p.fallback_positions(ch) == o.fallback_positions(ch)
        assert np.array_equal(p.stereo_matrix(ch, p.fallback_positions(ch)), o.stereo_matrix(ch, o.fallback_positions(ch)))
    odd = [capi.POS_LOW_FREQUENCY, capi.POS_AUX0, capi.POS_FRONT_RIGHT, capi.POS_UNKNOWN]
    assert np.array_equal(p.stereo_matrix(4, odd), o.stereo_matrix(4, odd))


def test_cpp_mirror_compiles_and_runs(lib, tmp_path):
    """include/omb200.hpp (C++ host mirror) builds against the library and its host-side classes behave."""
    import subprocess

    from openmeters_b200 import _lib

    exe = tmp_path / "mirror_smoke"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-std=c++20", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "mirror_smoke.cpp"),
                    "-L", libdir, "-lomb200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stdout + out.stderr


# ------------------------------------------------------------------ the Rust side (rust/: cannot be compiled here, no cargo)
def _gen():
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_sys", os.path.join(ROOT, "tools", "gen_rust_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_rust_sys_crate_is_generated_from_the_header_and_current():
    g = _gen()
    assert open(g.OUT).read() == g.render(g.parse_header()), "rust/omb200-sys/src/lib.rs is stale: python tools/gen_rust_sys.py"


def test_rust_externs_match_header_names_and_arity(lib):
    """Every `extern "C"` declaration of the -sys crate: declared in the header with the same number of parameters, exported
    by libomb200.so, and with a ctypes twin of the same arity (tests and bench bind through that table)."""
    g = _gen()
    rust = g.parse_rust_externs()
    hdr = {name: len(params) for name, _, params in g.parse_header()["functions"]}
    assert sorted(rust) == sorted(hdr) == header_functions()
    assert rust == hdr
    for name, n in rust.items():
        assert hasattr(lib, name), name
        assert len(capi.HEADER_SYMBOLS[name[4:]][1]) == n, (name, n, capi.HEADER_SYMBOLS[name[4:]][1])


def test_rust_wrapper_exposes_the_registry_method_set():
    """src/visuals/registry.rs:100-118 calls new / config / update_config / prepare / process_block / reset_audio on the
    spectrogram and spectrum processors and new / process_block / reset_audio on loudness: the wrapper must define them,
    and use only functions the -sys crate declares."""
    g = _gen()
    rust = g.parse_rust_externs()
    want = {"spectrogram.rs": ["new", "config", "update_config", "prepare", "reset_audio", "process_block"],
            "spectrum.rs": ["new", "config", "update_config", "prepare", "reset_audio", "process_block"],
            "loudness.rs": ["new", "reset_audio", "process_block"]}
    for fname, methods in want.items():
        src = open(os.path.join(ROOT, "rust", "omb200", "src", fname)).read()
        for m in methods:
            assert re.search(rf"pub fn {m}\(", src), (fname, m)
        for used in set(re.findall(r"sys::(omb_\w+)\(", src)):
            assert used in rust, (fname, used)
    # struct layouts the wrapper copies field by field: same field names as the header's structs
    hdr_structs = dict(g.parse_header()["structs"])
    src = open(os.path.join(ROOT, "rust", "omb200", "src", "spectrogram.rs")).read()
    for field, _ in hdr_structs["omb_spectrogram_config"]:
        assert field in src, field
