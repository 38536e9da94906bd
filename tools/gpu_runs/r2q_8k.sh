#!/bin/bash
# Round-2 GPU run Q (1 GPU): N = 8192 (cfg5) through stft_r64x.cu<128> vs stft_fast8k.cu
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "cfg5 or 16384 or 8192" > $O/r2q_pytest.log 2>&1; echo "exit $?" >> $O/r2q_pytest.log ); tail -3 $O/r2q_pytest.log
for rep in 1 2; do
timeout 300 python tools/bench_configs.py --only cfg5 > $O/r2q_cfg5_r64x_$rep.json 2> $O/r2q_cfg5_r64x.err; cat $O/r2q_cfg5_r64x_$rep.json; echo
OMB_R64X_8K=0 timeout 300 python tools/bench_configs.py --only cfg5 > $O/r2q_cfg5_fast8k_$rep.json 2> $O/r2q_cfg5_fast8k.err; cat $O/r2q_cfg5_fast8k_$rep.json; echo
done
timeout 300 python bench.py --config cfg5 --steps 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2q_bench_cfg5_r64x.json 2> $O/r2q_bench_cfg5_r64x.err; python -c "import json; d=json.loads([l for l in open('$O/r2q_bench_cfg5_r64x.json') if l.startswith('{')][-1]); print('bench cfg5 r64x', d['value'], d['ms_per_step'])"
OMB_R64X_8K=0 timeout 300 python bench.py --config cfg5 --steps 5 --no-cpu-baseline --no-secondary --e2e-steps 1 > $O/r2q_bench_cfg5_fast8k.json 2> $O/r2q_bench_cfg5_fast8k.err; python -c "import json; d=json.loads([l for l in open('$O/r2q_bench_cfg5_fast8k.json') if l.startswith('{')][-1]); print('bench cfg5 fast8k', d['value'], d['ms_per_step'])"
timeout 300 compute-sanitizer --tool racecheck python -c "
import sys; sys.path.insert(0,'.')
from openmeters_b200 import _capi as capi, batch, synth
from openmeters_b200._lib import api as lib_api
from openmeters_b200.processors import SpectrogramConfig
api=lib_api(); api.set_device(0)
cfg=SpectrogramConfig(sample_rate=96000.0, fft_size=8192, hop_size=2048, window=capi.WINDOW_BLACKMAN_HARRIS, use_reassignment=True)
S=16384+10*2048
lanes=synth.cfg5_lanes(2,S)[:, :S]
plan=batch.StftPlan(cfg, kernel=capi.KERNEL_FAST, api=api); print('gen', plan.kernel_generation)
p,c=plan.execute_host(lanes); print(c.min())
" > $O/r2q_racecheck_8k.log 2>&1; tail -3 $O/r2q_racecheck_8k.log
