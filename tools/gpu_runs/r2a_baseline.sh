#!/bin/bash
# Round-2 GPU run A: state of the shipped code before any round-2 kernel work.
# full GPU suite, headline bench, ncu --set full of k_reassigned_fast2<2> at bench size, cfg4 planar-ring A/B.
set -u
O=gpurun_out
mkdir -p $O
nproc > $O/r2a_env.txt; nvidia-smi -L >> $O/r2a_env.txt
( timeout 900 python -m pytest tests -m gpu -x -q > $O/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2a_pytest_gpu.log )
tail -3 $O/r2a_pytest_gpu.log
timeout 300 python bench.py > $O/r2a_bench_n1.json 2> $O/r2a_bench_n1.err
cut -c1-400 $O/r2a_bench_n1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2a_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/r2a_ncu_fast2.log 2>&1
timeout 200 python tools/bench_configs.py --only cfg4 > $O/r2a_cfg4_default.json 2> $O/r2a_cfg4_default.err
OMB_SPECTRUM_PLANAR=1 timeout 200 python tools/bench_configs.py --only cfg4 > $O/r2a_cfg4_planar.json 2> $O/r2a_cfg4_planar.err
cat $O/r2a_cfg4_default.json $O/r2a_cfg4_planar.json
