#!/bin/bash
# Round-2 end-of-round validation (1 GPU): full GPU suite, smoke, full-size parity sets, both bench arms, launch list and ncu capture of the
# bench command (-> profiles/traffic.json via tools/update_traffic.py), secondary configs, settings grid.
set -u
O=gpurun_out
mkdir -p $O
sha256sum openmeters_b200/libomb200.so > $O/r2z_lib_sha256.txt
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2z_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2z_pytest_gpu.log ); tail -3 $O/r2z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2z_smoke.log 2>&1; tail -1 $O/r2z_smoke.log
timeout 400 python bench.py > $O/r2z_bench_n1.json 2> $O/r2z_bench_n1.err
python -c "import json; d=json.loads([l for l in open('$O/r2z_bench_n1.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_image', d['e2e_image']['value'], 'frac', d['roofline']['frac'], 'traffic', d['roofline'].get('traffic')); print({k: (v['value'], v['hbm_frac']) for k, v in d['secondary'].items()})"
timeout 300 python bench.py --impl reference > $O/r2z_bench_reference_arm.json 2> $O/r2z_bench_reference_arm.err; tail -c 400 $O/r2z_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2z_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2z_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2z_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2z_ncu_fast2.log 2>&1; tail -1 $O/r2z_ncu_fast2.log
timeout 300 python tools/bench_configs.py > $O/r2z_configs.json 2> $O/r2z_configs.err
timeout 600 python tools/bench_grid.py > $O/r2z_grid.json 2> $O/r2z_grid.err
timeout 900 python tools/parity_fullsize.py --out $O/r2z_parity_fullsize.json > $O/r2z_parity_fullsize.log 2>&1; tail -7 $O/r2z_parity_fullsize.log
