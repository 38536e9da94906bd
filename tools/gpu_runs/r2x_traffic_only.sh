#!/bin/bash
# ncu capture of the bench command on the in-tree library (for profiles/traffic.json) + one bench line
set -u
O=gpurun_out
mkdir -p $O
sha256sum openmeters_b200/libomb200.so > $O/r2x_lib_sha256.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2x_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2x_ncu_fast2.log 2>&1; tail -1 $O/r2x_ncu_fast2.log
timeout 400 python bench.py --no-cpu-baseline > $O/r2x_bench_n1_nocpu.json 2> $O/r2x_bench_n1_nocpu.err
python -c "import json; d=json.loads([l for l in open('$O/r2x_bench_n1_nocpu.json') if l.startswith('{')][-1]); print('value', d['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic'))"
