#!/bin/bash
# 8-GPU box: the contract bench at N = 1, 2, 4, 8 back to back (what the driver does at round end); lines in gpurun_out/scale_N.json
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 100 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_1.json
for n in 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 100 --warmup 3 2>gpurun_out/scale_$n.err | grep '^{' | tail -1 > gpurun_out/scale_$n.json
done
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        d = json.load(open(f"gpurun_out/scale_{n}.json")); print(n, round(d["value"], 1), round(d["ms_per_step"], 4), d["e2e"]["value"])
    except Exception as e:
        print(n, "ERR", e)
PY
