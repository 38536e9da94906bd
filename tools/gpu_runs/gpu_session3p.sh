#!/bin/bash
# Session-3 GPU run P (last slot): epilogue diet — headline A/B first, then parity.
set -u
O=gpurun_out
mkdir -p $O
timeout 120 python bench.py --no-cpu-baseline --e2e-steps 1 > $O/s3p_bench_n1.json 2> $O/s3p_bench_n1.err
cut -c1-200 $O/s3p_bench_n1.json
timeout 100 python tools/bench_grid.py --first 12 > $O/s3p_grid.json 2> $O/s3p_grid.err
( timeout 300 python -m pytest tests -m gpu -q -x -k "cfg2 or cfg5 or default_size or size_1024 or small_hops or golden or kat or edge or generic_and_fast or settings_grid or bank" > $O/s3p_pytest_a.log 2>&1; echo "pytest exit $?" >> $O/s3p_pytest_a.log )
tail -3 $O/s3p_pytest_a.log
( timeout 300 python -m pytest tests -m gpu -q -x -k "not (cfg2 or cfg5 or default_size or size_1024 or small_hops or golden or kat or edge or generic_and_fast or settings_grid or bank)" > $O/s3p_pytest_b.log 2>&1; echo "pytest exit $?" >> $O/s3p_pytest_b.log )
tail -3 $O/s3p_pytest_b.log
