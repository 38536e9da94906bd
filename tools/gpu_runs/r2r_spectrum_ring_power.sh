#!/bin/bash
# Round-2 GPU run R (1 GPU): the two-kernel spectrum path with the fused kernel's front half as its power stage (hop segments) vs the
# frame-per-CTA power kernel, at the lane counts a 2- / 4- / 8-GPU shard of cfg4 sees; spectrum GPU tests
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "spectrum or cfg4 or peaks or bank" > $O/r2r_pytest.log 2>&1; echo "exit $?" >> $O/r2r_pytest.log ); tail -3 $O/r2r_pytest.log
timeout 300 python tools/bench_spectrum_lanes.py 8 16 32 64 128 > $O/r2r_lanes_ring.json 2> $O/r2r_lanes_ring.err; cat $O/r2r_lanes_ring.json; echo
OMB_SPECTRUM_RING_POWER=0 timeout 300 python tools/bench_spectrum_lanes.py 8 16 32 64 > $O/r2r_lanes_old.json 2> $O/r2r_lanes_old.err; cat $O/r2r_lanes_old.json; echo
OMB_SPECTRUM_FUSED=1 timeout 300 python tools/bench_spectrum_lanes.py 32 64 > $O/r2r_lanes_fused_pinned.json 2> $O/r2r_lanes_fused_pinned.err; cat $O/r2r_lanes_fused_pinned.json; echo
OMB_SPECTRUM_OVERLAP=0 timeout 300 python tools/bench_spectrum_lanes.py 8 16 32 64 > $O/r2r_lanes_serial.json 2> $O/r2r_lanes_serial.err; cat $O/r2r_lanes_serial.json; echo
timeout 300 python tools/bench_configs.py --only specbank > $O/r2r_specbank.json 2> $O/r2r_specbank.err; cat $O/r2r_specbank.json; echo
