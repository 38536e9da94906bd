#!/bin/bash
# Session-3 GPU run B: settings grid + new tests, cfg1 after the lane remap, ncu of cfg1 / cfg5 kernels.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q > $O/s3b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3b_pytest_gpu.log )
timeout 300 python tools/bench_configs.py --only cfg1,cfg4 > $O/s3b_configs.json 2> $O/s3b_configs.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_classic_1024 -s 3 -c 1 -f -o $O/s3b_classic python tools/bench_configs.py --only cfg1 > $O/s3b_ncu_classic.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_8k -s 2 -c 1 -f -o $O/s3b_8k python tools/bench_configs.py --only cfg5 > $O/s3b_ncu_8k.log 2>&1
tail -3 $O/s3b_pytest_gpu.log
cat $O/s3b_configs.json
