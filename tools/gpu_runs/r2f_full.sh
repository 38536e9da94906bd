#!/bin/bash
# Round-2 GPU run F (1 GPU): full GPU suite on the current code, spec-size parity sets, secondary configs, ncu captures of the
# shipped cfg2 kernel (bench size) and of the cfg5 kernel, launch list of the bench command.
set -u
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2f_pytest_gpu.log )
tail -4 $O/r2f_pytest_gpu.log
timeout 300 python tools/bench_configs.py --only cfg5,cfg4,loudbank > $O/r2f_configs.json 2> $O/r2f_configs.err; cat $O/r2f_configs.json
timeout 900 python tools/parity_fullsize.py --out $O/r2f_parity_fullsize.json > $O/r2f_parity_fullsize.log 2>&1; tail -8 $O/r2f_parity_fullsize.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2f_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2f_ncu_fast2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_8k -s 2 -c 1 -f -o $O/r2f_8k python tools/bench_configs.py --only cfg5 > $O/r2f_ncu_8k.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2f_launches_bench.log 2>&1
sha256sum openmeters_b200/libomb200.so > $O/r2f_lib_sha256.txt
timeout 400 python bench.py > $O/r2f_bench_n1.json 2> $O/r2f_bench_n1.err
python -c "import json; d=json.loads([l for l in open('$O/r2f_bench_n1.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_image', d['e2e_image']['value']); print({k: (v['value'], v['hbm_frac']) for k, v in d['secondary'].items()})"
