#!/usr/bin/env python
"""Copy the evidence of tools/evidence_r2.sh from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, launch list + summary, ncu summary of
the CG kernel with the stall samples per code region, phase trace, traffic stamp, sanitizer and micro-benchmark logs."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)

shutil.copy(os.path.join(G, "ev2_bench_line.json"), os.path.join(P, "r02_bench_line.json"))
shutil.copy(os.path.join(G, "ev2_ref_line.json"), os.path.join(P, "r02_reference_arm_line.json"))
shutil.copy(os.path.join(G, "ev2_trace.json"), os.path.join(P, "r02_trace_cg_iteration.json"))
shutil.copy(os.path.join(G, "ev2_micro.txt"), os.path.join(P, "r02_micro_latencies.txt"))
with open(os.path.join(P, "r02_sanitizer.log"), "w") as f:
    f.write("# compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_tet.py -k 'cg_solve_matches_oracle or euler_implicit_cg_steps'  (fused, fused all warps, fused streamed, v1, tail, multi-kernel; the edge pass of FastTetrahedralCorotationalForceField fused and multi-kernel; Vec3f + Vec3d)\n")
    f.write("".join(open(os.path.join(G, "ev2_racecheck.log")).readlines()[-6:]))
    f.write("\n# compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_hexa.py tests/test_gpu_fast_tet.py -k 'cg_solve_matches_oracle or hexa_steps or hexa_add or update_stiffness or add_force_and_add_dforce or euler_implicit_cg_steps or small_tiles'\n")
    f.write("".join(open(os.path.join(G, "ev2_memcheck.log")).readlines()[-5:]))
    f.write("\n# python -m pytest tests -m gpu (1 GPU)\n" + open(os.path.join(G, "ev2_pytest.log")).read())
# launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(G, "ev2_launches_raw.csv")) if l.startswith('"'))]
hdr = rows[0]
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
shutil.copy(os.path.join(G, "ev2_launches_raw.csv"), os.path.join(P, "r02_launches_raw.csv"))
agg = {}
for r in rows[1:]:
    k = r[iK].split("(")[0].replace("void ", "")[:90]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[iV].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, "r02_launches_summary.csv"), "w") as f:
    f.write("kernel,launches,total_ns,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{v[0]},{v[1]:.0f},{v[1] / tot:.4f}\n")
# ncu summary + per-region stall samples
rep = os.path.join(G, "ev2_fused.ncu-rep")
summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_ncu.py"), rep], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h = srows[1]; data = srows[2:]
iS, iSrc = h.index("# Samples"), h.index("Source")
total = sum(int(r[iS]) for r in data)
lines = ["", "# warp-state samples per code region (regions end at the listed barrier / release; all warps are sampled, waiting ones included)"]
acc = last = 0
for n, r in enumerate(data):
    acc += int(r[iS])
    s = r[iSrc].strip()
    if any(k in s for k in ("BAR.SYNC", "REDG", "EXIT")) and acc - last > 0.003 * total:
        lines.append(f"  instr {n:5d}  {s[:48]:50s} {acc - last:7d} samples ({100.0 * (acc - last) / total:5.1f} %)")
        last = acc
stall = {}
for r in data:
    for i, name in enumerate(h):
        if name.startswith("stall_") and "Not Issued" not in name:
            try:
                stall[name] = stall.get(name, 0) + int(r[i])
            except ValueError:
                pass
lines.append("# warp-state samples by reason, whole kernel: " + ", ".join(f"{k[6:]} {100.0 * v / total:.1f} %" for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:9]))
open(os.path.join(P, "r02_ncu_cg_fused.txt"), "w").write(summ + "\n".join(lines) + "\n")
subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "update_traffic.py"), rep, "C2_f32_cg_fused_bytes"], stdout=subprocess.DEVNULL)
d = json.load(open(os.path.join(P, "r02_bench_line.json")))
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "configs", {k: round(v.get("value", 0)) for k, v in d["configs"].items()})
print(open(os.path.join(P, "r02_launches_summary.csv")).read()[:900])
print(open(os.path.join(P, "r02_ncu_cg_fused.txt")).read()[-2500:])
