#!/bin/bash
# quick check of a kernel change: parity subset, bench without the side configs, phase trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --no-configs --no-cpu-variants --steps 100 --warmup 5 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/g_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel ms", d["roofline"]["avg_launch_ms"], "e2e", d["e2e"]["value"], d["roofline"].get("fused_kernel"))
PY
bash tools/r2_b.sh
