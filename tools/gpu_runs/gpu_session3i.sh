#!/bin/bash
# Session-3 GPU run I: new interleaved-frame kernels — tests, memcheck, grid throughput.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -q > $O/s3i_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/s3i_pytest_gpu.log )
( timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "default_size_2048 or size_1024_reassigned or small_hops" > $O/s3i_memcheck.log 2>&1; echo "memcheck exit $?" >> $O/s3i_memcheck.log )
timeout 600 python tools/bench_grid.py > $O/s3i_grid.json 2> $O/s3i_grid.err
tail -4 $O/s3i_pytest_gpu.log
tail -6 $O/s3i_memcheck.log
tail -2 $O/s3i_grid.err
