#!/bin/bash
# multi-GPU: parity tests, then the contract bench at N GPUs with the fused and the first-generation kernel
N=${1:-2}
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_parallel.py -x -q ) > gpurun_out/c_pytest_$N.log 2>&1
tail -8 gpurun_out/c_pytest_$N.log
for cfg in "SOFAB200_CG_FUSED=1" "SOFAB200_CG_FUSED=0"; do
  for n in 2 4 8; do
    if [ $n -le $N ]; then
      env $cfg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 100 --warmup 3 2>gpurun_out/c_scale_$n.err | grep '^{' | tail -1 > gpurun_out/c_scale_${n}_$cfg.json
      python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/c_scale_${n}_$cfg.json')); print('$cfg', $n, round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))
except Exception as e: print('$cfg', $n, 'ERR', e, open('gpurun_out/c_scale_$n.err').read()[-800:])
"
    fi
  done
done
