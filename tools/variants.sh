#!/bin/bash
# Measurement helper (GPU box): times the specialised STFT kernels under their tuning knobs.
# usage: tools/variants.sh "ENV1=a,ENV2=b ENV1=c ..." out.txt     (each word = one run; commas separate env assignments)
out=${2:-gpurun_out/variants.txt}
: > $out
for v in $1; do
  envs=$(echo $v | tr ',' ' ')
  r=$(env $envs timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['config']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])")
  echo "$v $r" | tee -a $out
done
