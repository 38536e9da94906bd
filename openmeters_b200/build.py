"""Builds openmeters_b200/libomb200.so in-tree with nvcc for sm_100a (B200).

    python -m openmeters_b200.build [--force] [--verbose]

No torch dependency: the library is plain CUDA runtime + C ABI. The emulator build used by the CPU
test-suite lives in tests/emu/build_emu.py (separate artefact, never loaded by this package).
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libomb200.so")
OBJ = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-math-errno",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


# Per-file extra flags (none at the moment).
EXTRA_FLAGS = {}


def nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name: str, defines: list[str]) -> str:
    """A/B build of the same sources with extra -D flags -> openmeters_b200/build_ab/libomb200_<name>.so (select it with
    OMB_LIB=<path>; measurement only)."""
    obj_dir = os.path.join(HERE, "build_ab", name)
    os.makedirs(obj_dir, exist_ok=True)
    out = os.path.join(HERE, "build_ab", f"libomb200_{name}.so")
    procs, objs = [], []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc(), *[f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")], *[f"-D{d}" for d in defines], "-c", src, "-o", obj]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(log)
    subprocess.run([nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out, *objs, "-cudart", "static",
                    "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"], check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(HERE, "..", "include", "omb200.h")]
    objs = []
    logs = []
    procs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc(), *NVCC_FLAGS, *EXTRA_FLAGS.get(os.path.basename(src), []), "-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        logs.append(f"==== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            failed = True
            sys.stderr.write(logs[-1])
    with open(os.path.join(OBJ, "ptxas.log"), "a" if not force else "w") as f:
        f.write("\n".join(logs))
    if failed:
        raise RuntimeError("nvcc failed")
    if verbose:
        print("\n".join(logs))
    if force or procs or _stale(OUT, objs):
        cmd = [nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", OUT, *objs, "-cudart", "static",
               "-Xlinker", "--no-undefined", "-lpthread", "-ldl", "-lrt"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m openmeters_b200.build --variant adds_only OMB_F32X2_LEVEL=1
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
