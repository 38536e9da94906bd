#!/bin/bash
# Round-2 GPU run O (1 GPU): compute-sanitizer (memcheck, racecheck, synccheck) over the round-2 kernels: generation 3 with the TMEM
# park and with the global park, generation 2 with the TMA ring (shipped) and with its tables in tensor memory; then the settings grid
# at N = 4096 for both generations (which kernel serves which hop)
set -u
O=gpurun_out
mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python tools/sanitizer_cases.py --only new > $O/r2o_${tool}_default.log 2>&1; tail -2 $O/r2o_${tool}_default.log
  OMB_FAST_KERNEL=3 OMB_R64_PARK=global timeout 400 compute-sanitizer --tool $tool python tools/sanitizer_cases.py --only new > $O/r2o_${tool}_r64_global.log 2>&1; tail -1 $O/r2o_${tool}_r64_global.log
  OMB_FAST_KERNEL=2 OMB_FAST2_BULK=7 timeout 400 compute-sanitizer --tool $tool python tools/sanitizer_cases.py --only new > $O/r2o_${tool}_gen2_tmemtab.log 2>&1; tail -1 $O/r2o_${tool}_gen2_tmemtab.log
done
OMB_FAST_KERNEL=2 timeout 200 python tools/bench_grid.py --first 8 > $O/r2o_grid_gen2.json 2> $O/r2o_grid_gen2.err
OMB_FAST_KERNEL=3 timeout 200 python tools/bench_grid.py --first 8 > $O/r2o_grid_gen3.json 2> $O/r2o_grid_gen3.err
timeout 200 python tools/bench_grid.py --first 8 > $O/r2o_grid_default.json 2> $O/r2o_grid_default.err
python - <<'PY'
import json
for n in ("gen2","gen3","default"):
    try:
        d=json.load(open(f"gpurun_out/r2o_grid_{n}.json"))
        rows=d["settings_grid"]
        print(n, [(r["fft_size"], r["hop"], r["tier"], "%.4g" % r["frames_per_s"]) for r in rows][2:8])
    except Exception as e:
        print(n, "ERR", e)
PY
