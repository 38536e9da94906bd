#!/usr/bin/env python
"""Prints the host topology facts the multi-GPU host path depends on (NUMA nodes, GPU <-> node affinity, allowed CPUs)
and measures pinned D2H bandwidth per GPU alone and with all GPUs copying at once."""
import glob
import os
import subprocess
import sys
import threading
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


print(sh("nvidia-smi topo -m"))
print("allowed cpus:", sorted(os.sched_getaffinity(0)))
print(sh("lscpu | grep -i -E 'numa|socket|model name'"))
for d in sorted(glob.glob("/sys/devices/system/node/node*")):
    print(d, open(d + "/cpulist").read().strip(), sh(f"grep MemTotal {d}/meminfo"))
n = torch.cuda.device_count()
for i in range(n):
    bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    print("gpu", i, bus)
print(sh("for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/class 2>/dev/null)\" = 0x030200 ]; then echo $d $(cat $d/numa_node); fi; done"))
print("mempolicy:", sh("cat /proc/self/numa_maps | head -3"))

SZ = 1 << 30
bufs = []
for i in range(n):
    torch.cuda.set_device(i)
    d = torch.empty(SZ, dtype=torch.uint8, device=f"cuda:{i}")
    h = torch.empty(SZ, dtype=torch.uint8).pin_memory()
    bufs.append((d, h, torch.cuda.Stream(device=i)))


def run(idx, reps=4, direction="d2h"):
    outs = {}

    def work(i):
        d, h, s = bufs[i]
        torch.cuda.set_device(i)
        with torch.cuda.stream(s):
            (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        s.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s):
            for _ in range(reps):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
        s.synchronize()
        outs[i] = reps * SZ / (time.perf_counter() - t0) / 1e9

    th = [threading.Thread(target=work, args=(i,)) for i in idx]
    [t.start() for t in th]
    [t.join() for t in th]
    return outs


for i in range(n):
    print("alone d2h gpu", i, run([i]))
print("all d2h", run(list(range(n))))
print("all h2d", run(list(range(n)), direction="h2d"))
if n >= 2:
    print("pair 0,1 d2h", run([0, 1]))
    print("pair 0,%d d2h" % (n - 1), run([0, n - 1]))
