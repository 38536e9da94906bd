#!/bin/bash
# Session-3 GPU run N: residue tiers — racecheck, full suite, grid throughput, headline sanity.
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitizer_cases.py > $O/s3n_racecheck.log 2>&1; echo "racecheck exit $?" >> $O/s3n_racecheck.log )
( timeout 900 python -m pytest tests -m gpu -q > $O/s3n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3n_pytest_gpu.log )
timeout 600 python tools/bench_grid.py > $O/s3n_grid.json 2> $O/s3n_grid.err
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > $O/s3n_bench_n1.json 2> $O/s3n_bench_n1.err
tail -3 $O/s3n_racecheck.log
tail -4 $O/s3n_pytest_gpu.log
tail -2 $O/s3n_grid.err
cut -c1-260 $O/s3n_bench_n1.json
