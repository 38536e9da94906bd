#!/bin/bash
# Round-2 GPU run W (1 GPU): N = 2048 / 1024 kernels with several iterations per CTA barrier
set -u
O=gpurun_out
mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q -k "2048 or 1024 or exact or golden or settings_grid or kat" > $O/r2w_pytest.log 2>&1; echo "exit $?" >> $O/r2w_pytest.log ); tail -3 $O/r2w_pytest.log
pg() { python -c "
import json
rows=json.load(open('$1'))['settings_grid']
print('$2', [(r['fft_size'], r['hop'], r['tier'][5:12], '%.4g' % r['frames_per_s']) for r in rows][:12])"; }
timeout 300 python tools/bench_grid.py --first 12 > $O/r2w_grid.json 2> $O/r2w_grid.err; pg $O/r2w_grid.json reps_auto
OMB_FAST2K_REPS=1 OMB_FAST1K_REPS=1 timeout 300 python tools/bench_grid.py --first 12 > $O/r2w_grid_reps1.json 2> $O/r2w_grid_reps1.err; pg $O/r2w_grid_reps1.json reps_1
timeout 400 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py > $O/r2w_racecheck.log 2>&1; tail -2 $O/r2w_racecheck.log
