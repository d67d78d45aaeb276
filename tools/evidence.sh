#!/bin/bash
# GPU box: everything the round's evidence needs, in one call.  Scratch output in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 > gpurun_out/ev_pytest.log
python __graft_entry__.py smoke > gpurun_out/ev_smoke.log 2>&1
python bench.py --steps 100 --warmup 3 2>gpurun_out/ev_bench.err | tail -1 > gpurun_out/ev_bench_line.json
python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/ev_ref.err | tail -1 > gpurun_out/ev_ref_line.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ev_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tet_cg_persistent -c 1 -f -o gpurun_out/ev_persist \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ev_ncu_full.log 2>&1
python tools/trace_phases.py --out gpurun_out/ev_trace.json > gpurun_out/ev_trace.log 2>&1
cat gpurun_out/ev_pytest.log; tail -3 gpurun_out/ev_smoke.log; cat gpurun_out/ev_bench_line.json; cat gpurun_out/ev_ref_line.json
