#!/usr/bin/env python
"""Kernel-level microbenchmark used while tuning: times GraphScatteredMatrix::apply (A*p = element pass + boundary gather)
and the CG vector kernels on the C2 workload with the context's per-class CUDA-event profiler.  Not a bench line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sofa_b200 as sb  # noqa: E402
from bench import SCENE, build_mesh  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--dtype", default="f32")
ap.add_argument("--reps", type=int, default=40)
args = ap.parse_args()
dtype, template = (np.float32, "B200Vec3f") if args.dtype == "f32" else (np.float64, "B200Vec3d")
pos, tets, fixed = build_mesh(args.workload)
ctx = sb.Context(0)
mo = sb.MechanicalObject(ctx, template, position=pos)
ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=SCENE["young"], poissonRatio=SCENE["poisson"], method="large", tileElems=args.tile)
mass = sb.DiagonalMass(mo, tets, massDensity=SCENE["density"])
node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=0.01, gravity=SCENE["gravity"], rayleighStiffness=0.1, rayleighMass=0.1,
                     iterations=25, tolerance=1e-9, threshold=1e-9)
rng = np.random.default_rng(0)
p = torch.from_numpy(rng.standard_normal(pos.shape).astype(dtype)).to(ctx.device)
q = mo.new_vector(); f = mo.new_vector()
node.computeForce(f, mo.x)
for _ in range(5):
    node.apply(q, p, 1.001, -0.01, -0.0011)
torch.cuda.synchronize()
ctx.profile_begin()
for _ in range(args.reps):
    node.apply(q, p, 1.001, -0.01, -0.0011)
prof = ctx.profile_end()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    node.apply(q, p, 1.001, -0.01, -0.0011)
e1.record(); torch.cuda.synchronize()
out = {k: round(1000 * v["ms"] / max(v["launches"], 1), 2) for k, v in prof.items() if v["launches"]}
out["apply_us_back_to_back"] = round(1000 * e0.elapsed_time(e1) / args.reps, 2)
out["layout"] = ff.stats()
print(json.dumps(out))
