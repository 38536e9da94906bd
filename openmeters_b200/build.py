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
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
