#!/bin/bash
# Round-2 GPU run S (1 GPU): generation 2 with one contiguous frame range per CTA vs round-robin runs
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
run() { local name=$1; shift; env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2s_bench_$name.json 2> $O/r2s_bench_$name.err; b $O/r2s_bench_$name.json $name; }
( timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or generic_and_fast" > $O/r2s_pytest.log 2>&1; echo "exit $?" >> $O/r2s_pytest.log ); tail -3 $O/r2s_pytest.log
run contig A=1
run runs OMB_FAST2_CONTIG=0
run contig_b A=1
run runs_b OMB_FAST2_CONTIG=0
run r64 OMB_FAST_KERNEL=3
timeout 200 python tools/bench_grid.py --first 8 > $O/r2s_grid.json 2> $O/r2s_grid.err
python -c "
import json
rows=json.load(open('gpurun_out/r2s_grid.json'))['settings_grid']
print([(r['fft_size'], r['hop'], r['tier'], '%.4g' % r['frames_per_s']) for r in rows][2:8])"
