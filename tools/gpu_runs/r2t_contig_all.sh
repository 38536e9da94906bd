#!/bin/bash
# Round-2 GPU run T (1 GPU): contiguous work assignment in all four ring kernels (N = 4096 / 8192 / 2048 / 1024): parity + grid + cfg5
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "cfg2 or cfg5 or 2048 or 1024 or exact or small_hops or settings_grid" > $O/r2t_pytest.log 2>&1; echo "exit $?" >> $O/r2t_pytest.log ); tail -3 $O/r2t_pytest.log
pg() { python -c "
import json
rows=json.load(open('$1'))['settings_grid']
print('$2', [(r['fft_size'], r['hop'], r['tier'][5:12], '%.4g' % r['frames_per_s']) for r in rows][:14])"; }
timeout 300 python tools/bench_grid.py --first 14 > $O/r2t_grid_default.json 2> $O/r2t_grid_default.err; pg $O/r2t_grid_default.json default
OMB_FAST_KERNEL=2 OMB_R64X_8K=0 timeout 300 python tools/bench_grid.py --first 12 > $O/r2t_grid_ring_kernels.json 2> $O/r2t_grid_ring_kernels.err; pg $O/r2t_grid_ring_kernels.json ring_kernels_pinned
OMB_FAST2_CONTIG=0 OMB_FAST8K_CONTIG=0 OMB_FAST2K_CONTIG=0 OMB_FAST1K_CONTIG=0 OMB_FAST_KERNEL=2 OMB_R64X_8K=0 timeout 300 python tools/bench_grid.py --first 12 > $O/r2t_grid_runs.json 2> $O/r2t_grid_runs.err; pg $O/r2t_grid_runs.json round_robin_runs
timeout 300 python tools/bench_configs.py --only cfg5 > $O/r2t_cfg5.json 2> $O/r2t_cfg5.err; cat $O/r2t_cfg5.json; echo
OMB_FAST8K_CONTIG=0 timeout 300 python tools/bench_configs.py --only cfg5 > $O/r2t_cfg5_runs.json 2> $O/r2t_cfg5_runs.err; cat $O/r2t_cfg5_runs.json; echo
timeout 300 python bench.py --config cfg5 --steps 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2t_bench_cfg5.json 2> $O/r2t_bench_cfg5.err; python -c "import json; d=json.loads([l for l in open('$O/r2t_bench_cfg5.json') if l.startswith('{')][-1]); print('bench cfg5', d['value'], d['ms_per_step'])"
