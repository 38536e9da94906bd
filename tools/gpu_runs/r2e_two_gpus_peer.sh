#!/bin/bash
# Round-2 GPU run E (2 GPUs): scatter-inclusive with the three transports (nccl / peer_dma / peer_direct), cfg2 + cfg5 + cfg4.
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
show() { python - "$1" <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1], 'value %.4g'%d['value'], 'ms %.3f'%d['ms_per_step'])
si=d.get('scatter_inclusive')
if si:
    print('  best', si['mode'], '%.4g'%si['value'], 'checksum ok', si['checksum_equals_resident'])
    for m,v in si['modes'].items(): print('   ', m, '%.4g'%v['value'], 'ms %.3f'%v['ms_per_step'], 'alone', v.get('transfer_alone_ms'), 'egress GB/s %.0f'%v['rank0_egress_gbs'])
print('  e2e %.4g'%d['e2e']['value'], 'e2e_image', d.get('e2e_image',{}).get('value'))
PY
}
timeout 600 $TR bench.py --gpus 2 > $O/r2e_bench_n2_cfg2.json 2> $O/r2e_bench_n2_cfg2.err || tail -20 $O/r2e_bench_n2_cfg2.err
show $O/r2e_bench_n2_cfg2.json
timeout 600 $TR bench.py --gpus 2 --config cfg5 --steps 5 --cfg5-lanes 512 > $O/r2e_bench_n2_cfg5.json 2> $O/r2e_bench_n2_cfg5.err || tail -20 $O/r2e_bench_n2_cfg5.err
show $O/r2e_bench_n2_cfg5.json
timeout 600 $TR bench.py --gpus 2 --config cfg4 --steps 5 > $O/r2e_bench_n2_cfg4.json 2> $O/r2e_bench_n2_cfg4.err || tail -20 $O/r2e_bench_n2_cfg4.err
show $O/r2e_bench_n2_cfg4.json
( timeout 600 python -m pytest tests -m gpu -x -q -k "loud or bank" > $O/r2e_pytest_loud.log 2>&1; echo "pytest exit $?" >> $O/r2e_pytest_loud.log )
tail -3 $O/r2e_pytest_loud.log
timeout 300 python tools/bench_configs.py --only loudbank > $O/r2e_loudbank.json 2> $O/r2e_loudbank.err; cat $O/r2e_loudbank.json
OMB_LOUDNESS_STREAM_SEQ=1 timeout 300 python tools/bench_configs.py --only loudbank > $O/r2e_loudbank_seq.json 2> $O/r2e_loudbank_seq.err; cat $O/r2e_loudbank_seq.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2e_smoke.log 2>&1; tail -2 $O/r2e_smoke.log
