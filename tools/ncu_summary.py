#!/usr/bin/env python
"""Summarises an `ncu --set full` report (.ncu-rep) into a small JSON for profiles/: the raw-page metrics the design
discussion uses plus per-opcode shares and stall-reason shares from the source page.

    python tools/ncu_summary.py report.ncu-rep out.json [--units N --unit-name frames] [--note "..."]
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    args = sys.argv[3:]
    opt = {args[i]: args[i + 1] for i in range(0, len(args) - 1, 2)}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    res = {"report": rep.split("/")[-1], "kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else None, "metrics": {}}
    for k in KEYS:
        if k in col:
            res["metrics"][k] = [vals[col[k]], units[col[k]]]
    dram = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        if k in col:
            dram += float(vals[col[k]].replace(",", "")) * SCALE.get(units[col[k]], 1)
    res["dram_bytes"] = dram
    if "--units" in opt:
        n = float(opt["--units"])
        name = opt.get("--unit-name", "units")
        res[name + "_in_launch"] = n
        res["dram_bytes_per_" + name[:-1]] = dram / n
        res["warp_instructions_per_" + name[:-1]] = float(vals[col["smsp__inst_executed.sum"]].replace(",", "")) / n
    if "--note" in opt:
        res["note"] = opt["--note"]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    byop, exe = collections.Counter(), collections.Counter()
    for r in data:
        t = r[ix["Source"]].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        byop[op] += int(r[ix["# Samples"]])
        exe[op] += int(r[ix["Instructions Executed"]])
    te = sum(exe.values()) or 1
    res["opcodes"] = {op: {"sample_pct": round(100 * s / tot, 2), "executed_pct": round(100 * exe[op] / te, 2)} for op, s in byop.most_common(14)}
    res["stall_sample_pct"] = {c: round(100 * sum(int(r[ix[c]]) for r in data) / tot, 2) for c in hdr
                               if c.startswith("stall_") and "Not Issued" not in c and sum(int(r[ix[c]]) for r in data) * 200 > tot}
    res["local_memory_instructions_static"] = sum(1 for r in data if "LDL" in r[ix["Source"]] or "STL" in r[ix["Source"]])
    json.dump(res, open(out, "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
