#!/usr/bin/env python
"""Timing of the hexahedral path on config C3 (grid 65x65x121 -> 491 520 hexahedra, HexahedronFEMForceField
method=polar, Vec3f): steps/s and CG iterations/s of the device-resident step.  Not the bench line (bench.py is the tetra C2)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sofa_b200 as sb  # noqa: E402
from sofa_b200 import topology as T  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", default="65,65,121")
ap.add_argument("--method", default="polar")
ap.add_argument("--steps", type=int, default=20)
args = ap.parse_args()
n = tuple(int(v) for v in args.n.split(","))
pos, hexas = T.regular_grid(n, (0, 0, 0), (8, 8, 15))
fixed = T.box_roi(pos, (-1, -1, -1, 9, 9, 1e-6))
ctx = sb.Context(0)
mo = sb.MechanicalObject(ctx, "B200Vec3f", position=pos)
ff = sb.HexahedronFEMForceField(mo, hexas, youngModulus=1000.0, poissonRatio=0.3, method=args.method)
mass = sb.DiagonalMass(mo, hexas, massDensity=1.0)
node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=0.01, gravity=(0.0, -9.0, 0.0), rayleighStiffness=0.1, rayleighMass=0.1,
                     iterations=25, tolerance=1e-9, threshold=1e-9)
for _ in range(4):
    node.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    node.step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
it = min(node.last_solve()["iterations"], 25)
H, N = hexas.shape[0], pos.shape[0]
alg = H * (32 + 309 * 4) + N * 34 * 4           # bytes per CG iteration, Vec3f
print(json.dumps({"workload": f"C3 grid {n}: {H} hexahedra, {N} nodes, method={args.method}, Vec3f", "ms_per_step": ms, "steps_per_s": 1e3 / ms,
                  "cg_iters_per_s": it * 1e3 / ms, "algorithmic_GBps_whole_step_as_cg": alg * it / (ms * 1e-3) / 1e9, "layout": ff.stats()}))
