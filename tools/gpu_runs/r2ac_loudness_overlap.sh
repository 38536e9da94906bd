#!/bin/bash
# Round-2 GPU run AC (1 GPU): loudness batch path with the true-peak kernel on a side stream next to the K-weighting chain; then the
# ncu capture of the bench command for profiles/traffic.json on this (final) library
set -u
O=gpurun_out
mkdir -p $O
sha256sum openmeters_b200/libomb200.so > $O/r2x_lib_sha256.txt
( timeout 300 python -m pytest tests -m gpu -x -q -k "loud or cfg3" > $O/r2ac_pytest.log 2>&1; echo "exit $?" >> $O/r2ac_pytest.log ); tail -2 $O/r2ac_pytest.log
for rep in 1 2; do
timeout 200 python tools/bench_configs.py --only cfg3 > $O/r2ac_cfg3_overlap_$rep.json 2> $O/r2ac_cfg3.err; cat $O/r2ac_cfg3_overlap_$rep.json; echo
OMB_LOUDNESS_OVERLAP=0 timeout 200 python tools/bench_configs.py --only cfg3 > $O/r2ac_cfg3_serial_$rep.json 2> $O/r2ac_cfg3.err; cat $O/r2ac_cfg3_serial_$rep.json; echo
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_fast2 -s 3 -c 1 -f -o $O/r2x_fast2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2x_ncu_fast2.log 2>&1; tail -1 $O/r2x_ncu_fast2.log
