#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
run() { local name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2k_bench_$name.json 2> $O/r2k_bench_$name.err; b $O/r2k_bench_$name.json $name; }
run helper A=1
run nohelper OMB_R64_HELPER=0
run helper_b A=1
run nohelper_b OMB_R64_HELPER=0
