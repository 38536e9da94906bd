#!/usr/bin/env python
"""Prints registers / spills per kernel from an nvcc -Xptxas -v log (openmeters_b200/build/ptxas.log)."""
import re
import subprocess
import sys


def parse(path):
    out = {}
    name = None
    for line in open(path):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void omb::", "")
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and name:
            out.setdefault(name, {})["spill"] = (int(m.group(2)), int(m.group(3)))
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            out.setdefault(name, {})["regs"] = int(m.group(1))
    return out


if __name__ == "__main__":
    a = parse(sys.argv[1])
    b = parse(sys.argv[2]) if len(sys.argv) > 2 else None
    for k in sorted(a):
        row = f"{k:60s} regs {a[k].get('regs')} spill {a[k].get('spill')}"
        if b and k in b:
            row += f"   -> regs {b[k].get('regs')} spill {b[k].get('spill')}"
        print(row)
