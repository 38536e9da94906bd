#!/bin/bash
# Measurement helper (GPU box): times the specialised STFT kernel under its tuning knobs.
# usage: tools/variants.sh "MINB:TW MINB:TW ..." out.txt
out=${2:-gpurun_out/variants.txt}
: > $out
for v in $1; do
  minb=${v%%:*}; tw=${v##*:}
  r=$(OMB_FAST_MINB=$minb OMB_FAST_TW=$tw timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['clocks'])")
  echo "minb=$minb tw=$tw $r" | tee -a $out
done
