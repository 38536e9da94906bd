#!/bin/bash
# Session-3 GPU run: tests, secondary configs, ncu captures of the cfg1 / cfg4 / cfg3 kernels, headline bench.
set -u
O=gpurun_out
mkdir -p $O
python -c "import torch; print(torch.cuda.get_device_name(0))" > $O/s3_env.txt 2>&1
nproc >> $O/s3_env.txt
( timeout 900 python -m pytest tests -m gpu -x -q > $O/s3_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3_pytest_gpu.log )
timeout 300 python tools/bench_configs.py > $O/s3_configs.json 2> $O/s3_configs.err
timeout 300 python bench.py > $O/s3_bench_n1.json 2> $O/s3_bench_n1.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_classic_1024 -s 3 -c 1 -f -o $O/s3_classic python tools/bench_configs.py --only cfg1 > $O/s3_ncu_classic.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_spectrum_fused -s 2 -c 1 -f -o $O/s3_specfused python tools/bench_configs.py --only cfg4 > $O/s3_ncu_spec.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_true_peak4|k_kw_chunks|k_loud_snapshots" -s 10 -c 4 -f -o $O/s3_loud python tools/bench_configs.py --only cfg3 > $O/s3_ncu_loud.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s3_launches_configs.csv python tools/bench_configs.py --only cfg1,cfg3,cfg4 > $O/s3_launches.log 2>&1
tail -3 $O/s3_pytest_gpu.log
cat $O/s3_configs.json
cat $O/s3_bench_n1.json
