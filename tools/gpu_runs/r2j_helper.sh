#!/bin/bash
# Round-2 GPU run J (1 GPU): stft_r64.cu with the helper warpgroup (setmaxnreg 232 / 40) vs without, A/B of pair-step chunk and pruning
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2j_bench_$name.json 2> $O/r2j_bench_$name.err; b $O/r2j_bench_$name.json $name
}
( timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or generic_and_fast" > $O/r2j_pytest.log 2>&1; echo "exit $?" >> $O/r2j_pytest.log ); tail -3 $O/r2j_pytest.log
run helper_tmem A=1
run helper_global OMB_R64_PARK=global
run nohelper_tmem OMB_R64_HELPER=0
run helper_pc16 OMB_LIB=$PWD/openmeters_b200/build_ab/libomb200_pc16.so
run helper_prune OMB_LIB=$PWD/openmeters_b200/build_ab/libomb200_prune.so
run nohelper_pc16 OMB_R64_HELPER=0 OMB_LIB=$PWD/openmeters_b200/build_ab/libomb200_pc16.so
run nohelper_prune OMB_R64_HELPER=0 OMB_LIB=$PWD/openmeters_b200/build_ab/libomb200_prune.so
run gen2 OMB_FAST_KERNEL=2
run helper_tmem_b A=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_r64 -s 3 -c 1 -f -o $O/r2j_r64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2j_ncu_r64.log 2>&1; tail -2 $O/r2j_ncu_r64.log
