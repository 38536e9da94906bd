#!/bin/bash
# Round-2 GPU run U (1 GPU): generation 2 with several frame pairs per loop iteration (one CTA barrier / ring prefetch per iteration)
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
run() { local name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2u_bench_$name.json 2> $O/r2u_bench_$name.err; b $O/r2u_bench_$name.json $name; }
( timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or generic_and_fast or small_hops" > $O/r2u_pytest.log 2>&1; echo "exit $?" >> $O/r2u_pytest.log ); tail -3 $O/r2u_pytest.log
run pairs2 A=1
run pairs1 OMB_FAST2_PAIRS=1
run pairs2_b A=1
run pairs1_b OMB_FAST2_PAIRS=1
timeout 400 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py --only new > $O/r2u_racecheck.log 2>&1; tail -2 $O/r2u_racecheck.log
timeout 200 python tools/bench_grid.py --first 8 > $O/r2u_grid.json 2> $O/r2u_grid.err
python -c "
import json
rows=json.load(open('gpurun_out/r2u_grid.json'))['settings_grid']
print([(r['fft_size'], r['hop'], r['tier'][5:12], '%.4g' % r['frames_per_s']) for r in rows][2:8])"
