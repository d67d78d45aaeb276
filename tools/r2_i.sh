#!/bin/bash
mkdir -p gpurun_out
SOFAB200_FUSED_CACHED_KB=0 timeout 300 python tools/trace_phases.py > gpurun_out/i_trace_c2_streamed.log 2>&1
timeout 300 python tools/trace_phases.py --dtype f64 > gpurun_out/i_trace_c2_f64.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/i_trace_c2_streamed.log", "gpurun_out/i_trace_c2_f64.log"):
    try:
        d = json.load(open(f)); print(f); print(json.dumps(d.get("cg_fused_iteration_10_durations_us"), indent=0)); print(d.get("cg_fused_kernel_us"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-600:])
PY
