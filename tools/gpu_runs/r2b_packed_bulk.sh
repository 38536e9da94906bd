#!/bin/bash
# Round-2 GPU run B: packed FP32x2 complex primitives + TMA bulk ring / column store.
# GPU suite (incl. the new exact-math tests), then A/B: {round-1 packing, full packing} x {LDGSTS, bulk in, bulk in+out}.
set -u
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2b_pytest_gpu.log )
tail -5 $O/r2b_pytest_gpu.log
AB=openmeters_b200/build_ab/libomb200_adds_only.so
for rep in 1 2; do
for lib in default adds_only; do
  for bulk in 0 1 3; do
    if [ $lib = adds_only ]; then export OMB_LIB=$PWD/$AB; else unset OMB_LIB; fi
    OMB_FAST2_BULK=$bulk timeout 200 python bench.py --no-cpu-baseline --e2e-steps 1 > $O/r2b_bench_${lib}_bulk${bulk}_$rep.json 2> $O/r2b_bench_${lib}_bulk${bulk}_$rep.err
    python -c "import json,sys; d=json.load(open('$O/r2b_bench_${lib}_bulk${bulk}_$rep.json')); print('$lib bulk$bulk rep$rep', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
  done
done
done
unset OMB_LIB
timeout 300 python tools/bench_configs.py > $O/r2b_configs_default.json 2> $O/r2b_configs_default.err
OMB_LIB=$PWD/$AB timeout 300 python tools/bench_configs.py > $O/r2b_configs_adds_only.json 2> $O/r2b_configs_adds_only.err
cat $O/r2b_configs_default.json; echo; cat $O/r2b_configs_adds_only.json; echo
timeout 200 python tools/bench_grid.py > $O/r2b_grid_default.json 2> $O/r2b_grid_default.err
OMB_LIB=$PWD/$AB timeout 200 python tools/bench_grid.py > $O/r2b_grid_adds_only.json 2> $O/r2b_grid_adds_only.err
( OMB_SPECTRUM_PLANAR=1 timeout 600 python -m pytest tests -m gpu -x -q -k "spectrum or cfg4" > $O/r2b_pytest_planar.log 2>&1; echo "pytest exit $?" >> $O/r2b_pytest_planar.log )
tail -3 $O/r2b_pytest_planar.log
