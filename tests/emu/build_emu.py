"""Builds tests/emu/libomb200_emu.so: the product's kernel sources compiled with g++ against the
fiber-based CUDA emulator (tests/emu/cuda_emu.h).  TEST/DEVELOPMENT AID ONLY — lets the CPU suite
execute the real kernel code thread-for-thread; openmeters_b200 never loads it."""
from __future__ import annotations

import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "openmeters_b200", "csrc")
OUT = os.path.join(HERE, "libomb200_emu.so")
OBJ = os.path.join(HERE, "build")

FLAGS = ["-O2", "-g", "-std=c++17", "-fPIC", "-DOMB_EMU", "-ffp-contract=off", "-fno-math-errno", "-x", "c++",
         "-I", HERE, "-I", CSRC, "-pthread", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-unused-variable",
         "-Wno-unused-but-set-variable", "-Wno-unused-parameter"]


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    deps = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + [
        os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "omb200.h")]
    newest_dep = max(os.path.getmtime(d) for d in deps)
    objs, procs = [], []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(newest_dep, os.path.getmtime(src)):
            procs.append((src, subprocess.Popen(["g++", *FLAGS, "-c", src, "-o", obj], stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT, text=True)))
    bad = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            bad = True
            print(f"==== {src}\n{out}")
        elif out.strip():
            print(f"==== {os.path.basename(src)} (warnings)\n{out}")
    if bad:
        raise RuntimeError("emulator build failed")
    if force or procs or not os.path.exists(OUT):
        subprocess.run(["g++", "-shared", "-o", OUT, *objs, "-pthread"], check=True)
    return OUT


if __name__ == "__main__":
    import sys
    print(build("--force" in sys.argv))
