#!/bin/bash
# Round-2 GPU run C (1 GPU): new tests (render_host, mid-stream KATs), the rewritten bench.py at N = 1 (secondary, e2e_image),
# A/B of the rotation-packed build, bulk variants on the final switches.
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "render or kat or exact or cfg2" > $O/r2c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/r2c_pytest_gpu.log )
tail -5 $O/r2c_pytest_gpu.log
timeout 600 python bench.py > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err
tail -3 $O/r2c_bench_n1.err
python -c "import json; d=json.load(open('$O/r2c_bench_n1.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'e2e_image', d.get('e2e_image',{}).get('value'), 'frac', d['roofline']['frac']); print(json.dumps(d.get('secondary'), indent=0)[:1500])"
for bulk in 0 1 3; do
  OMB_FAST2_BULK=$bulk timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 2 > $O/r2c_bench_bulk$bulk.json 2> $O/r2c_bench_bulk$bulk.err
  python -c "import json; d=json.load(open('$O/r2c_bench_bulk$bulk.json')); print('final-switches bulk$bulk', d['value'], d['ms_per_step'])"
done
OMB_LIB=$PWD/openmeters_b200/build_ab/libomb200_rot.so timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 2 > $O/r2c_bench_rot.json 2> $O/r2c_bench_rot.err
python -c "import json; d=json.load(open('$O/r2c_bench_rot.json')); print('rot-packed bulk3', d['value'], d['ms_per_step'])"
OMB_LIB=$PWD/openmeters_b200/build_ab/libomb200_rot.so timeout 200 python tools/bench_grid.py --first 12 > $O/r2c_grid_rot.json 2> $O/r2c_grid_rot.err
timeout 200 python tools/bench_grid.py --first 12 > $O/r2c_grid_final.json 2> $O/r2c_grid_final.err
