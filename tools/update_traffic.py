#!/usr/bin/env python
"""Writes profiles/traffic.json from an `ncu --set full` summary (tools/ncu_summary.py output) of the bench command:
DRAM bytes per unit of the dominant kernel, stamped with the digest of the libomb200.so it was captured from (bench.py reports
`roofline.traffic` only when the digest matches the library it is running).

    python tools/update_traffic.py profiles/r02z_ncu_full_fast2.json k_reassigned_fast2 frame"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    summary, kernel, unit = sys.argv[1], sys.argv[2], sys.argv[3]
    d = json.load(open(summary))
    h = hashlib.sha256()
    with open(os.path.join(ROOT, "openmeters_b200", "libomb200.so"), "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    out = {k: v for k, v in out.items() if isinstance(v, dict)}
    out[kernel] = {"dram_bytes_per_unit": d["dram_bytes_per_" + unit], "unit": unit, "library_sha256_16": h.hexdigest()[:16],
                   "source": os.path.relpath(summary, ROOT)}
    json.dump(out, open(path, "w"), indent=1)
    print(path, out[kernel])


if __name__ == "__main__":
    main()
