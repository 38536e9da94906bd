#!/bin/bash
# Round-2 GPU run P (1 GPU): stft_r64x.cu (N = 16384 reassigned on chip): parity, sanitizers, throughput next to the generic tier
set -u
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "16384 or settings_grid" > $O/r2p_pytest.log 2>&1; echo "exit $?" >> $O/r2p_pytest.log ); tail -4 $O/r2p_pytest.log
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitizer_cases.py --only new > $O/r2p_${tool}.log 2>&1; tail -2 $O/r2p_${tool}.log
done
timeout 300 python tools/bench_grid.py --first 14 > $O/r2p_grid.json 2> $O/r2p_grid.err
OMB_NO_R64X=1 timeout 300 python tools/bench_grid.py --first 14 > $O/r2p_grid_no_r64x.json 2> $O/r2p_grid_no_r64x.err
python - <<'PY'
import json
for n in ("r2p_grid","r2p_grid_no_r64x"):
    try:
        rows=json.load(open(f"gpurun_out/{n}.json"))["settings_grid"]
        print(n, [(r["fft_size"], r["hop"], r["tier"], "%.4g" % r["frames_per_s"]) for r in rows][12:14])
    except Exception as e:
        print(n, "ERR", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_reassigned_r64x -s 2 -c 1 -f -o $O/r2p_r64x python tools/bench_grid.py --first 13 > $O/r2p_ncu.log 2>&1; tail -2 $O/r2p_ncu.log
