#!/bin/bash
# 8-GPU box: multi-GPU parity (worlds 4 and 8), weak scaling of C2 and strong scaling of C5 at N = 4, 8.  Logs in gpurun_out/ (copied to profiles/).
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/test_gpu_parallel.py -q -k "(4 or 8) and (peer or nccl) and not peer_v1" ) > gpurun_out/f_pytest_multi.log 2>&1
tail -6 gpurun_out/f_pytest_multi.log
run() { # name N args...
  name=$1; N=$2; shift 2
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus $N "$@" ) 2> gpurun_out/f_${name}_$N.err | grep '^{' | tail -1 > gpurun_out/f_${name}_$N.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/f_${name}_$N.json"))
    print("${name}", $N, "value", round(d["value"],1), d["unit"], "true", round(d["true_cg_iters_per_s"],1), "ms/step", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "create_s", round(d["create_s"],1), "frac", round(d["roofline"]["frac"],3), "parity", d["parity_check"]["max_abs_dx"] if d.get("parity_check") else None, "cg_ms", d["roofline"]["kernel_ms_rank0"].get("cg_persistent"))
except Exception as e:
    print("${name}", $N, "ERR", e); print(open("gpurun_out/f_${name}_$N.err").read()[-1500:])
PY
}
run weak 8 --steps 100 --warmup 3
run weak 4 --steps 100 --warmup 3
run c5 8 --workload C5 --steps 20 --warmup 3
run c5 4 --workload C5 --steps 20 --warmup 3
run c5rcb 8 --workload C5 --partition rcb --steps 20 --warmup 3
