#!/bin/bash
# usage: tools/sweep_env.sh VAR "v1 v2 ..." [bench args]   -> one bench line per value in gpurun_out/sweep_VAR.log
var=$1; vals=$2; shift 2
mkdir -p gpurun_out
: > gpurun_out/sweep_$var.log
for v in $vals; do
  echo "== $var=$v" >> gpurun_out/sweep_$var.log
  env $var=$v python bench.py --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import sys, json
l = sys.stdin.read().strip()
try:
    d = json.loads(l); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step')} | {'cg_kernel_ms': d['roofline'].get('avg_launch_ms'), 'frac': d['roofline']['frac'], 'e2e': d['e2e']['value']}))
except Exception as e:
    print('ERR', l[:300])
" >> gpurun_out/sweep_$var.log
done
cat gpurun_out/sweep_$var.log
