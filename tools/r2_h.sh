#!/bin/bash
# C5 on one GPU: fused (streamed) against the multi-kernel loop and the fused tail; C2 f64 the same
mkdir -p gpurun_out
run() { # name, workload, dtype, env...
  n=$1; w=$2; dt=$3; shift 3
  env "$@" timeout 600 python bench.py --workload $w --dtype $dt --no-configs --no-cpu-variants --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/h_$n.json 2> gpurun_out/h_$n.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/h_$n.json").read().strip().splitlines()[-1])
    print("$n", "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), d["roofline"]["kernel"][:40], d["roofline"].get("kernel_ms"))
except Exception as e:
    print("$n", "ERR", e, open("gpurun_out/h_$n.err").read()[-400:])
PY
}
run c5_fused C5 f32 A=1
run c5_multi C5 f32 SOFAB200_CG_PERSISTENT=0 SOFAB200_FUSED_TAIL=0
run c5_tail C5 f32 SOFAB200_CG_PERSISTENT=0
run c2d_fused C2 f64 A=1
run c2d_multi C2 f64 SOFAB200_CG_PERSISTENT=0 SOFAB200_FUSED_TAIL=0
run c2d_tail C2 f64 SOFAB200_CG_PERSISTENT=0
