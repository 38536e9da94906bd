#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
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
nvidia-smi topo -m > $O/r2m2_topo.txt 2>&1
timeout 500 $TR bench.py --gpus 8 --no-cpu-baseline > $O/r2m2_bench_n8_cfg2.json 2> $O/r2m2_bench_n8_cfg2.err || tail -20 $O/r2m2_bench_n8_cfg2.err
show $O/r2m2_bench_n8_cfg2.json
