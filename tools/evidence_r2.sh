#!/bin/bash
# GPU box: everything the round's evidence needs, in one call.  Scratch output in gpurun_out/ (copied to profiles/ by tools/collect_r2.sh).
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 > gpurun_out/ev2_pytest.log
python __graft_entry__.py smoke > gpurun_out/ev2_smoke.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ev2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --sustain-s 0 > gpurun_out/ev2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_cg_kernel --launch-skip 3 -c 1 -f -o gpurun_out/ev2_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --sustain-s 0 > gpurun_out/ev2_ncu_full.log 2>&1
# the DRAM traffic of that capture, stamped with the source hash, BEFORE the bench line is taken (bench.py reports roofline.traffic only for a matching stamp;
# the stamped file is also written to gpurun_out/ so that it comes back)
python tools/update_traffic.py gpurun_out/ev2_fused.ncu-rep C2_f32_cg_fused_bytes > gpurun_out/ev2_traffic.log 2>&1; cp profiles/traffic.json gpurun_out/ev2_traffic.json
python bench.py --steps 100 --warmup 3 2>gpurun_out/ev2_bench.err | tail -1 > gpurun_out/ev2_bench_line.json
python bench.py --impl reference --steps 20 --warmup 3 2>gpurun_out/ev2_ref.err | tail -1 > gpurun_out/ev2_ref_line.json
python tools/trace_phases.py > gpurun_out/ev2_trace.json 2> gpurun_out/ev2_trace.err
( compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_tet.py -x -q -k "cg_solve_matches_oracle or euler_implicit_cg_steps" ) > gpurun_out/ev2_racecheck.log 2>&1
( compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_hexa.py tests/test_gpu_fast_tet.py -x -q -k "cg_solve_matches_oracle or hexa_steps or hexa_add or update_stiffness or add_force_and_add_dforce or euler_implicit_cg_steps or small_tiles" ) > gpurun_out/ev2_memcheck.log 2>&1
./tools/micro/lat > gpurun_out/ev2_micro.txt 2>&1; ./tools/micro/div >> gpurun_out/ev2_micro.txt 2>&1
cat gpurun_out/ev2_pytest.log; tail -3 gpurun_out/ev2_smoke.log; cut -c1-400 gpurun_out/ev2_bench_line.json; cut -c1-300 gpurun_out/ev2_ref_line.json; tail -3 gpurun_out/ev2_racecheck.log; tail -2 gpurun_out/ev2_memcheck.log
