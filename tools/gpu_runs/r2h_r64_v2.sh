#!/bin/bash
# Round-2 GPU run H (1 GPU): stft_r64.cu v2 (Im c parked as complex c in TMEM, DIT pass B, 128-bit readback, double-buffered pair step)
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
( timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or generic_and_fast" > $O/r2h_pytest.log 2>&1; echo "exit $?" >> $O/r2h_pytest.log ); tail -3 $O/r2h_pytest.log
for rep in 1 2; do
OMB_FAST_KERNEL=2 timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2h_bench_gen2_$rep.json 2> $O/r2h_bench_gen2.err; b $O/r2h_bench_gen2_$rep.json gen2
OMB_R64_PARK=global timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2h_bench_r64_global_$rep.json 2> $O/r2h_bench_r64_global.err; b $O/r2h_bench_r64_global_$rep.json r64_global
timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2h_bench_r64_tmem_$rep.json 2> $O/r2h_bench_r64_tmem.err; b $O/r2h_bench_r64_tmem_$rep.json r64_tmem
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_r64 -s 3 -c 1 -f -o $O/r2h_r64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2h_ncu_r64.log 2>&1; tail -2 $O/r2h_ncu_r64.log
