#!/bin/bash
# Round-2 GPU run L (1 GPU): generation-2 cfg2 kernel with its tables in tensor memory (OMB_FAST2_BULK=7) vs the shipped variant
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
run() { local name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2l_bench_$name.json 2> $O/r2l_bench_$name.err; b $O/r2l_bench_$name.json $name; }
( OMB_FAST_KERNEL=2 OMB_FAST2_BULK=7 timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact" > $O/r2l_pytest.log 2>&1; echo "exit $?" >> $O/r2l_pytest.log ); tail -3 $O/r2l_pytest.log
run gen2_tmemtab OMB_FAST_KERNEL=2 OMB_FAST2_BULK=7
run gen2 OMB_FAST_KERNEL=2
run gen2_tmemtab_b OMB_FAST_KERNEL=2 OMB_FAST2_BULK=7
run gen2_b OMB_FAST_KERNEL=2
run r64 A=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2l_fast2_tmemtab env OMB_FAST_KERNEL=2 OMB_FAST2_BULK=7 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2l_ncu.log 2>&1; tail -2 $O/r2l_ncu.log
