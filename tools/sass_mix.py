#!/usr/bin/env python
"""Static SASS opcode mix of one kernel of libomb200.so (cuobjdump -sass): proves which instructions the build contains.

    python tools/sass_mix.py <kernel-substring> [--lib path] [--dump out.sass]
"""
import collections
import re
import subprocess
import sys


def main():
    pat = sys.argv[1]
    lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else "openmeters_b200/libomb200.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if pat not in dem:
            continue
        ops = collections.Counter()
        n = 0
        for line in b.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                ops[m.group(1).split(".")[0]] += 1
                n += 1
        print(f"{dem}\n  {n} instructions: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(24)))
        if "--dump" in sys.argv:
            open(sys.argv[sys.argv.index("--dump") + 1], "w").write("Function : " + b)
            break


if __name__ == "__main__":
    main()
