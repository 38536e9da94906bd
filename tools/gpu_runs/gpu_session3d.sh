#!/bin/bash
# Session-3 GPU run D: loudness after the packed-multiply true peak / constant-bank zero-state pass.
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -q -k "loud or Loud or kat or golden or meter" > $O/s3d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3d_pytest_gpu.log )
timeout 300 python tools/bench_configs.py --only cfg3 > $O/s3d_configs.json 2> $O/s3d_configs.err
OMB_KW_ZERO_RECURRENCE=1 timeout 300 python tools/bench_configs.py --only cfg3 > $O/s3d_configs_recurrence.json 2>> $O/s3d_configs.err
OMB_NO_TRUE_PEAK4=1 timeout 300 python tools/bench_configs.py --only cfg3 > $O/s3d_configs_generic_tp.json 2>> $O/s3d_configs.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_true_peak4|k_kw_zero_state|k_kw_chunks" -s 9 -c 3 -f -o $O/s3d_loud python tools/bench_configs.py --only cfg3 > $O/s3d_ncu_loud.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/s3d_launches_cfg3.csv python tools/bench_configs.py --only cfg3 > $O/s3d_launches3.log 2>&1
tail -4 $O/s3d_pytest_gpu.log
cat $O/s3d_configs.json $O/s3d_configs_recurrence.json $O/s3d_configs_generic_tp.json
