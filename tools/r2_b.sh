#!/bin/bash
mkdir -p gpurun_out
SOFAB200_FUSED_GATHER_WARPS=0 timeout 300 python tools/trace_phases.py > gpurun_out/b_trace_gw0.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/b_trace_gw0.log",):
    try:
        d = json.load(open(f)); print(f); print(json.dumps(d.get("cg_fused_iteration_10_marks_us"))); print(json.dumps(d.get("cg_fused_iteration_10_durations_us"), indent=0)); print(d.get("cg_fused_kernel_us"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-600:])
PY
