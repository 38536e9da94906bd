#!/usr/bin/env python
"""Per-source-line totals of an ncu report's source page (CUDA + SASS view): stall samples, executed warp instructions and shared-memory
wavefronts per line of the kernel's .cu / .cuh files.   python tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur_file, hdr = None, None
    acc = collections.defaultdict(lambda: [0, 0, 0, ""])
    line_key = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            isamp = r.index("# Samples")
            iexe = r.index("Instructions Executed")
            iwav = r.index("L1 Wavefronts Shared")
            continue
        if hdr is None or len(r) < 10:
            continue
        if r[0] != "":  # a source line
            line_key = (cur_file, int(r[0]))
            acc[line_key][3] = r[1].strip()[:110]
        elif line_key is not None:  # a SASS line under it
            a = acc[line_key]
            num = lambda x: int(x) if x.strip().isdigit() else 0
            a[0] += num(r[isamp])
            a[1] += num(r[iexe])
            a[2] += num(r[iwav])
    tot = sum(a[0] for a in acc.values()) or 1
    tote = sum(a[1] for a in acc.values()) or 1
    print(f"total samples {tot}, executed warp instructions {tote}")
    for k, a in sorted(acc.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * a[0] / tot:6.2f}% samp {100 * a[1] / tote:6.2f}% exec {a[2]:>12} wav  {k[0]}:{k[1]}  {a[3]}")


if __name__ == "__main__":
    main()
