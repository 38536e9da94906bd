#!/bin/bash
# Session-3 GPU run M: engine sync change — racecheck, tests, headline + grid A/B.
set -u
O=gpurun_out
mkdir -p $O
( timeout 300 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python tools/sanitizer_cases.py > $O/s3m_racecheck.log 2>&1; echo "racecheck exit $?" >> $O/s3m_racecheck.log )
( timeout 900 python -m pytest tests -m gpu -q > $O/s3m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3m_pytest_gpu.log )
timeout 300 python bench.py --no-cpu-baseline --e2e-steps 2 > $O/s3m_bench_n1.json 2> $O/s3m_bench_n1.err
timeout 300 python tools/bench_grid.py --first 14 > $O/s3m_grid.json 2> $O/s3m_grid.err
tail -3 $O/s3m_racecheck.log
tail -4 $O/s3m_pytest_gpu.log
cut -c1-260 $O/s3m_bench_n1.json
