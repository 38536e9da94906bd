#!/bin/bash
# Session-3 GPU run K: full suite on the final code, all secondary configs incl. the new banks, headline bench with the FP32 probe.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q > $O/s3k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3k_pytest_gpu.log )
timeout 500 python tools/bench_configs.py > $O/s3k_configs.json 2> $O/s3k_configs.err
timeout 300 python bench.py > $O/s3k_bench_n1.json 2> $O/s3k_bench_n1.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/s3k_smoke.log 2>&1
tail -4 $O/s3k_pytest_gpu.log
cat $O/s3k_configs.json
tail -2 $O/s3k_configs.err
cat $O/s3k_bench_n1.json
tail -2 $O/s3k_smoke.log
