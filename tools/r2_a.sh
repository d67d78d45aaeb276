#!/bin/bash
# round 2, first GPU call: parity of the fused CG kernel, sanitizer, A/B timing against the first-generation kernel, phase traces
mkdir -p gpurun_out; rm -f gpurun_out/a_bench.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/a_pytest.log
cat gpurun_out/a_pytest.log
line() { python -c "
import sys, json
l = sys.stdin.read().strip().splitlines()
try:
    d = json.loads(l[-1]); print(json.dumps({k: d[k] for k in ('value', 'ms_per_step')} | {'cg_kernel_ms': d['roofline'].get('avg_launch_ms'), 'frac': d['roofline']['frac'], 'e2e': d['e2e']['value'], 'kernel': d['roofline']['kernel'][:40]}))
except Exception as e:
    print('ERR', l[-3:])
"; }
for cfg in "SOFAB200_FUSED_GATHER_WARPS=0" "SOFAB200_FUSED_GATHER_WARPS=4" "SOFAB200_FUSED_GATHER_WARPS=2" "SOFAB200_FUSED_GATHER_WARPS=4 SOFAB200_FUSED_GATHER_SHARE=100" "SOFAB200_CG_FUSED=0" "SOFAB200_FUSED_CACHED_KB=0"; do
  echo "== $cfg" | tee -a gpurun_out/a_bench.log
  env $cfg timeout 300 python bench.py --no-cpu-baseline --steps 50 2>&1 | line | tee -a gpurun_out/a_bench.log
done
echo "== f64" | tee -a gpurun_out/a_bench.log
timeout 300 python bench.py --no-cpu-baseline --steps 30 --dtype f64 2>&1 | line | tee -a gpurun_out/a_bench.log
echo "== C5" | tee -a gpurun_out/a_bench.log
timeout 600 python bench.py --no-cpu-baseline --steps 10 --workload C5 2>&1 | line | tee -a gpurun_out/a_bench.log
SOFAB200_FUSED_GATHER_WARPS=4 timeout 300 python tools/trace_phases.py > gpurun_out/a_trace_gw4.log 2>&1
SOFAB200_FUSED_GATHER_WARPS=0 timeout 300 python tools/trace_phases.py > gpurun_out/a_trace_gw0.log 2>&1
python - <<PY
import json
for f in ("gpurun_out/a_trace_gw4.log", "gpurun_out/a_trace_gw0.log"):
    try:
        d = json.load(open(f)); print(f); print(json.dumps(d.get("cg_fused_iteration_10_marks_us"))); print(json.dumps(d.get("cg_fused_iteration_10_durations_us"), indent=0)); print(d.get("cg_fused_kernel_us"))
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-600:])
PY
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -x -q -k "cg_solve_matches_oracle and fused and not fused_tail" ) > gpurun_out/a_racecheck.log 2>&1
tail -5 gpurun_out/a_racecheck.log
( timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -k "cg_solve_matches_oracle and fused and not fused_tail" ) > gpurun_out/a_memcheck.log 2>&1
tail -5 gpurun_out/a_memcheck.log
