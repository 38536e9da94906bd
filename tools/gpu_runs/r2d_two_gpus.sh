#!/bin/bash
# Round-2 GPU run D (2 GPUs): bench.py under torchrun with the scatter ingest, cfg2 / cfg4 / cfg5.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/r2d_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 > $O/r2d_bench_n2_cfg2.json 2> $O/r2d_bench_n2_cfg2.err
tail -3 $O/r2d_bench_n2_cfg2.err
python -c "import json; d=json.load(open('$O/r2d_bench_n2_cfg2.json')); print('cfg2 N=2 value', d['value'], 'scatter', d.get('scatter_inclusive'), 'e2e', d['e2e']['value'], 'e2e_image', d.get('e2e_image',{}).get('value'))"
timeout 600 $TR bench.py --gpus 2 --config cfg4 --steps 5 > $O/r2d_bench_n2_cfg4.json 2> $O/r2d_bench_n2_cfg4.err
tail -3 $O/r2d_bench_n2_cfg4.err
python -c "import json; d=json.load(open('$O/r2d_bench_n2_cfg4.json')); print('cfg4 N=2 value', d['value'], 'scatter', d.get('scatter_inclusive'), 'e2e', d['e2e']['value'])"
timeout 600 $TR bench.py --gpus 2 --config cfg5 --steps 5 --cfg5-lanes 512 > $O/r2d_bench_n2_cfg5.json 2> $O/r2d_bench_n2_cfg5.err
tail -3 $O/r2d_bench_n2_cfg5.err
python -c "import json; d=json.load(open('$O/r2d_bench_n2_cfg5.json')); print('cfg5 N=2 value', d['value'], 'scatter', d.get('scatter_inclusive'), 'e2e', d['e2e']['value'])"
timeout 300 python bench.py --config cfg4 --steps 5 > $O/r2d_bench_n1_cfg4.json 2> $O/r2d_bench_n1_cfg4.err
python -c "import json; d=json.load(open('$O/r2d_bench_n1_cfg4.json')); print('cfg4 N=1 value', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])"
