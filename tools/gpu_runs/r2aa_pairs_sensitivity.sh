#!/bin/bash
# How much do more frame pairs per CTA barrier buy?  N = 4096 hop 512 (the ring holds 4 pairs) with the cap at 1 / 2 / 3 / 4
set -u
O=gpurun_out
mkdir -p $O
for p in 1 2 3 4; do
OMB_FAST2_PAIRS=$p timeout 200 python tools/bench_grid.py --first 4 > $O/r2aa_grid_p$p.json 2> $O/r2aa_grid_p$p.err
python -c "
import json
rows=json.load(open('$O/r2aa_grid_p$p.json'))['settings_grid']
print('pairs<=$p', [(r['fft_size'], r['hop'], '%.4g' % r['frames_per_s']) for r in rows][2:4])"
done
