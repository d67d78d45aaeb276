#!/bin/bash
# GPU box: parity tests, phase trace of one CG iteration, a bench line (no CPU baseline).  Scratch output in gpurun_out/.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/trace_phases.py --out gpurun_out/trace_new.json > gpurun_out/trace_new.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/trace_new.log"))["cg_persistent_last_iteration"]
for k in ("iter_start","tiles_done","barrier1","shared_done","barrier2","xr_done","barrier3","durations_median_us","detail_us"): print(k, d[k])
PY
python bench.py --no-cpu-baseline --steps 100 "$@" 2>&1 | tail -1 > gpurun_out/bench_new.json
python -c "
import json; d=json.load(open('gpurun_out/bench_new.json')); print('it/s', d['value'], 'ms/step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'cg ms', d['roofline']['avg_launch_ms'], 'e2e', d['e2e']['value'])"
