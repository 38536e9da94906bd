#!/bin/bash
# Session-3 GPU run C: full GPU suite, all secondary configs, headline bench + its launch list, loudness captures.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q > $O/s3c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3c_pytest_gpu.log )
timeout 400 python tools/bench_configs.py > $O/s3c_configs.json 2> $O/s3c_configs.err
timeout 300 python bench.py > $O/s3c_bench_n1.json 2> $O/s3c_bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/s3c_bench_reference_arm.json 2> $O/s3c_bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s3c_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/s3c_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_true_peak4|k_kw_chunks" -s 9 -c 3 -f -o $O/s3c_loud python tools/bench_configs.py --only cfg3 > $O/s3c_ncu_loud.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/s3c_launches_cfg3.csv python tools/bench_configs.py --only cfg3 > $O/s3c_launches3.log 2>&1
tail -5 $O/s3c_pytest_gpu.log
cat $O/s3c_configs.json
cat $O/s3c_bench_n1.json
cat $O/s3c_bench_reference_arm.json
