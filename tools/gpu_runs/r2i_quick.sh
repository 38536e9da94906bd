#!/bin/bash
# quick A/B of the cfg2 kernels (bench only)
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
TAG=${1:-r2i}
for rep in 1 2; do
timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/${TAG}_bench_r64_$rep.json 2> $O/${TAG}_bench_r64.err; b $O/${TAG}_bench_r64_$rep.json r64
done
OMB_FAST_KERNEL=2 timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/${TAG}_bench_gen2.json 2> $O/${TAG}_bench_gen2.err; b $O/${TAG}_bench_gen2.json gen2
