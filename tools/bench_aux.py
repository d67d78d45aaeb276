#!/usr/bin/env python
"""Timings of the per-operation kernels added late in round 1, at C2 size (983 040 tetrahedra, 175 329 nodes, Vec3f): MeshMatrixMass::addMDx,
getRotations, computeVonMisesStress, addForce with the plasticity branch.  CUDA events on the library's stream, 50 calls after 5 warm-ups.
Diagnostics for profiles/README.md; not a bench line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sofa_b200 as sb  # noqa: E402
from bench import SCENE, build_mesh  # noqa: E402

pos, tets, fixed = build_mesh("C2")
ctx = sb.Context(0)
mo = sb.MechanicalObject(ctx, "B200Vec3f", position=pos)
N, T = pos.shape[0], tets.shape[0]
x = torch.from_numpy((pos + 0.01 * np.random.default_rng(0).standard_normal(pos.shape)).astype(np.float32)).to(ctx.device)


def timed(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3   # us


out = {}
mm = sb.MeshMatrixMass(mo, tets, massDensity=1.0)
E = mm.edges.shape[0]
res, dx = mo.new_vector(), x.clone()
us = timed(lambda: mm.addMDx(res, dx, 0.5))
alg = N * (12 + 12 + 12 + 4) + 2 * E * 8          # res in/out, dx, vertex mass + one 8-byte record per half-edge (neighbour dx reads hit L2)
out["MeshMatrixMass.addMDx"] = {"us": round(us, 2), "edges": int(E), "algorithmic_MB": round(alg / 1e6, 2), "GB_s": round(alg / us / 1e3, 1)}
ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=SCENE["young"], poissonRatio=SCENE["poisson"], method="large", computeVonMisesStress=1)
f = mo.new_vector()
us_plain = timed(lambda: ff.addForce(f, x))
out["addForce(large)"] = {"us": round(us_plain, 2)}
R = ff.getRotations()
us = timed(lambda: ff.getRotations(R))
alg = 4 * T * (36 + 36 + 8) + N * 36               # per incident element: rotation (3 quads) + R0 (36 B) + two indices; 9 Reals out per node
out["getRotations"] = {"us": round(us, 2), "algorithmic_MB": round(alg / 1e6, 2), "GB_s": round(alg / us / 1e3, 1)}
us = timed(lambda: ff.computeVonMisesStress(x))
alg = T * (8 + 48 + 48 + 48 + 8 + 4 + 48) + 4 * T * 4 + N * 4   # lnode, X0, shape functions, rotation write, lambda/mu, out + per-node mean
out["computeVonMisesStress(1)"] = {"us": round(us, 2), "algorithmic_MB": round(alg / 1e6, 2), "GB_s": round(alg / us / 1e3, 1)}
del ff
ffp = sb.TetrahedronFEMForceField(mo, tets, youngModulus=SCENE["young"], poissonRatio=SCENE["poisson"], method="large", plasticMaxThreshold=0.5, plasticYieldThreshold=1e-4,
                                  plasticCreep=0.9)
us = timed(lambda: ffp.addForce(f, x))
out["addForce(large, plasticity)"] = {"us": round(us, 2), "extra_vs_plain_us": round(us - us_plain, 2), "extra_bytes_per_tet": 64}
print(json.dumps(out, indent=1))
