#!/bin/bash
# Round-2 GPU run G (1 GPU): first contact of stft_r64.cu with the hardware.  Order: the global-park variant first (no tcgen05),
# then the TMEM park; each step under its own timeout.  Then the A/B against generation 2, an ncu capture, streaming loudness.
set -u
O=gpurun_out
mkdir -p $O
b() { python -c "import json,sys; d=json.loads([l for l in open('$1') if l.startswith('{')][-1]); print('$2', d['value'], d['ms_per_step'], d.get('clocks',{}).get('sm_mhz'))"; }
echo "== parity, global park"
( OMB_R64_PARK=global timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or generic_and_fast" > $O/r2g_pytest_global.log 2>&1; echo "exit $?" >> $O/r2g_pytest_global.log ); tail -3 $O/r2g_pytest_global.log
echo "== parity, TMEM park"
( timeout 600 python -m pytest tests -m gpu -x -q -k "cfg2 or exact or generic_and_fast" > $O/r2g_pytest_tmem.log 2>&1; echo "exit $?" >> $O/r2g_pytest_tmem.log ); tail -3 $O/r2g_pytest_tmem.log
for rep in 1 2; do
OMB_FAST_KERNEL=2 timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2g_bench_gen2_$rep.json 2> $O/r2g_bench_gen2.err; b $O/r2g_bench_gen2_$rep.json gen2
OMB_R64_PARK=global timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2g_bench_r64_global_$rep.json 2> $O/r2g_bench_r64_global.err; b $O/r2g_bench_r64_global_$rep.json r64_global
timeout 200 python bench.py --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2g_bench_r64_tmem_$rep.json 2> $O/r2g_bench_r64_tmem.err; b $O/r2g_bench_r64_tmem_$rep.json r64_tmem
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_r64 -s 3 -c 1 -f -o $O/r2g_r64 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2g_ncu_r64.log 2>&1; tail -2 $O/r2g_ncu_r64.log
timeout 300 python tools/bench_configs.py --only loudbank > $O/r2g_loudbank.json 2> $O/r2g_loudbank.err; cat $O/r2g_loudbank.json
( timeout 300 python -m pytest tests -m gpu -x -q -k "loudness" > $O/r2g_pytest_loud.log 2>&1; echo "exit $?" >> $O/r2g_pytest_loud.log ); tail -3 $O/r2g_pytest_loud.log
