#!/bin/bash
# Round-2 GPU run V (1 GPU): fused spectrum kernel after keeping the smoothing modes' whole-lane loop compile-time (cfg4 rate), spectrum tests,
# then the ncu capture of the bench command for profiles/traffic.json on this build
set -u
O=gpurun_out
mkdir -p $O
sha256sum openmeters_b200/libomb200.so > $O/r2v_lib_sha256.txt
( timeout 600 python -m pytest tests -m gpu -x -q -k "spectrum or cfg4 or peaks" > $O/r2v_pytest.log 2>&1; echo "exit $?" >> $O/r2v_pytest.log ); tail -3 $O/r2v_pytest.log
timeout 300 python tools/bench_configs.py --only cfg4 > $O/r2v_cfg4.json 2> $O/r2v_cfg4.err; cat $O/r2v_cfg4.json; echo
timeout 300 python tools/bench_spectrum_lanes.py 32 64 128 > $O/r2v_lanes.json 2> $O/r2v_lanes.err; cat $O/r2v_lanes.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2v_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2v_ncu_fast2.log 2>&1; tail -1 $O/r2v_ncu_fast2.log
timeout 400 python bench.py > $O/r2v_bench_n1.json 2> $O/r2v_bench_n1.err
python -c "import json; d=json.loads([l for l in open('$O/r2v_bench_n1.json') if l.startswith('{')][-1]); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_image', d['e2e_image']['value'], 'frac', d['roofline']['frac']); print({k: (v['value'], v['hbm_frac']) for k, v in d['secondary'].items()})"
