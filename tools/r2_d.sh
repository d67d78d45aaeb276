#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 > gpurun_out/d_pytest.log
cat gpurun_out/d_pytest.log
( time python bench.py --steps 100 --warmup 3 ) > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -3 gpurun_out/d_bench.err
python - <<PY
import json
l=open("gpurun_out/d_bench.json").read().strip().splitlines()
try:
    d=json.loads(l[0])
    print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "create_s", d["create_s"], "launches", d["gpu_launches"])
    print("sustained", d["sustained"]["value"], d["sustained"]["seconds"], d["sustained"]["clocks"])
    print("fused", d["roofline"]["fused_kernel"], "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_source"])
    for k,c in d.get("configs",{}).items():
        print(k, {x: c.get(x) for x in ("value","ms_per_step","create_s","error")}, c.get("roofline",{}).get("frac"), c.get("roofline",{}).get("kernel","")[:30])
    print("cpu", d.get("cpu_baseline"))
except Exception as e:
    print("ERR", e, l[-2:] if l else "")
PY
( time python bench.py --impl reference --steps 5 --warmup 1 ) 2>&1 | tail -4 | cut -c1-1500
