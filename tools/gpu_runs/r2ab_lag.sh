#!/bin/bash
# Round-2 GPU run AB (1 GPU): generation 2 with the lagged cross-group barrier (bar.arrive / bar.sync, OMB_FAST2_LAG=1) vs the CTA barrier
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
run() { local name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2ab_bench_$name.json 2> $O/r2ab_bench_$name.err; b $O/r2ab_bench_$name.json $name; }
( OMB_FAST2_LAG=1 timeout 400 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or small_hops" > $O/r2ab_pytest.log 2>&1; echo "exit $?" >> $O/r2ab_pytest.log ); tail -3 $O/r2ab_pytest.log
run lag OMB_FAST2_LAG=1
run cta A=1
run lag_b OMB_FAST2_LAG=1
run cta_b A=1
OMB_FAST2_LAG=1 timeout 300 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py --only new > $O/r2ab_racecheck.log 2>&1; tail -2 $O/r2ab_racecheck.log
OMB_FAST2_LAG=1 timeout 200 python tools/bench_grid.py --first 8 > $O/r2ab_grid.json 2> $O/r2ab_grid.err
python -c "
import json
rows=json.load(open('gpurun_out/r2ab_grid.json'))['settings_grid']
print([(r['fft_size'], r['hop'], '%.4g' % r['frames_per_s']) for r in rows][2:8])"
