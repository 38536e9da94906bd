#!/bin/bash
# Round-2 last GPU run (1 GPU): full GPU suite, smoke, bench, ncu capture of the bench command for profiles/traffic.json — on the final build
set -u
O=gpurun_out
mkdir -p $O
sha256sum openmeters_b200/libomb200.so > $O/r2x_lib_sha256.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2x_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2x_pytest_gpu.log ); tail -3 $O/r2x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2x_smoke.log 2>&1; tail -1 $O/r2x_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2x_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2x_ncu_fast2.log 2>&1; tail -1 $O/r2x_ncu_fast2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2x_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2x_launches_bench.log 2>&1
timeout 400 python bench.py > $O/r2x_bench_n1.json 2> $O/r2x_bench_n1.err
python -c "import json; d=json.loads([l for l in open('$O/r2x_bench_n1.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_image', d['e2e_image']['value'], 'frac', d['roofline']['frac']); print({k: (v['value'], v['hbm_frac']) for k, v in d['secondary'].items()})"


