#!/bin/bash
# N-GPU box, short: weak C2 x N and strong C5 / N bench lines (parity check inside each), logs in gpurun_out/
N=${1:-8}
mkdir -p gpurun_out
run() { # name args...
  name=$1; shift
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@" ) 2> gpurun_out/j_${name}_$N.err | grep '^{' | tail -1 > gpurun_out/j_${name}_$N.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/j_${name}_$N.json"))
    print("${name}", $N, "value", round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), "parity", d["parity_check"]["max_abs_dx"] if d.get("parity_check") else None)
except Exception as e:
    print("${name}", $N, "ERR", e); print(open("gpurun_out/j_${name}_$N.err").read()[-1500:])
PY
}
run weak --steps 100 --warmup 3
run c5 --workload C5 --steps 20 --warmup 3
