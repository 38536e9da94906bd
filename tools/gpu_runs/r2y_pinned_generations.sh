#!/bin/bash
# Round-2 GPU run Y (1 GPU): the GPU suite's spectrogram tests with each N = 4096 kernel generation pinned, and N = 8192 through the team kernel
set -u
O=gpurun_out
mkdir -p $O
K="cfg2 or exact or kat or golden or bank or meter or render or generic_and_fast or small_hops or edge or settings_grid"
( OMB_FAST_KERNEL=3 timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > $O/r2y_pytest_gen3.log 2>&1; echo "exit $?" >> $O/r2y_pytest_gen3.log ); tail -3 $O/r2y_pytest_gen3.log
( OMB_FAST_KERNEL=3 OMB_R64_PARK=global timeout 900 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or small_hops" > $O/r2y_pytest_gen3_global.log 2>&1; echo "exit $?" >> $O/r2y_pytest_gen3_global.log ); tail -3 $O/r2y_pytest_gen3_global.log
( OMB_FAST_KERNEL=2 OMB_FAST2_BULK=7 timeout 900 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or small_hops" > $O/r2y_pytest_gen2_tmemtab.log 2>&1; echo "exit $?" >> $O/r2y_pytest_gen2_tmemtab.log ); tail -3 $O/r2y_pytest_gen2_tmemtab.log
( OMB_R64X_8K=1 timeout 900 python -m pytest tests -m gpu -x -q -k "cfg5 or 8192 or exact" > $O/r2y_pytest_r64x_8k.log 2>&1; echo "exit $?" >> $O/r2y_pytest_r64x_8k.log ); tail -3 $O/r2y_pytest_r64x_8k.log
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > $O/r2y_bench_n1.json 2> $O/r2y_bench_n1.err
python -c "import json; d=json.loads([l for l in open('$O/r2y_bench_n1.json') if l.startswith('{')][-1]); print('value', d['value'], 'traffic', d['roofline'].get('traffic'), d['roofline'].get('traffic_source'))"
