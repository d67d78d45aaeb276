#!/usr/bin/env python
"""The reference's own published benchmark scenes for this path (applications/plugins/SofaCUDA/scenes/benchmarks/
TetrahedronFEMForceField_beam10x10x40_gpu.scn and _beam16x16x76_gpu.scn; numbers in SofaCUDA/doc/SofaCUDA_benchmarks.csv, see
BASELINE.md), rebuilt from their Data on the device-resident solver node: RegularGridTopology + Hexa2TetraTopologicalMapping,
DiagonalMass totalMass=50, BoxROI + FixedProjectiveConstraint, TetrahedronFEMForceField method=large, PlaneForceField floor
(16x16x76 only), EulerImplicitSolver rayleigh 0.1/0.1 + CGLinearSolver.  Protocol of the sheet: 1000 steps, steps per second.
The collision pipeline of the scenes is empty (no collision model on the beam) and is not reproduced.  Not the bench line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sofa_b200 as sb  # noqa: E402
from sofa_b200 import topology as T  # noqa: E402

SCENES = {
    "hexa_beam10x10x40": dict(n=(40, 10, 10), mn=(0, 6, -2), mx=(16, 10, 2), dt=0.01, iters=20, young=2000.0, plane=None, hexa=True,
                              published="GPU CudaVec3f 226.0 (RTX 2070) / 143.1 (GTX 1060) steps/s; CPU 26.49 / 35.10 (csv:7-10)"),
    "hexa_beam16x16x76": dict(n=(76, 16, 16), mn=(0, 6, -2), mx=(19, 10, 2), dt=0.04, iters=10, young=1000.0, plane=dict(normal=(0, 1, 0), d=2.0, stiffness=10000.0), hexa=True,
                              published="GPU CudaVec3f 158.0 (RTX 2070) / 51.9 (GTX 1060) steps/s; CPU 9.18 / 9.94 (csv:11-14)"),
    "beam10x10x40": dict(n=(40, 10, 10), mn=(0, 6, -2), mx=(16, 10, 2), dt=0.01, iters=20, young=2000.0, plane=None,
                         published="GPU CudaVec3f 357.3 (RTX 2070) / 261.6 (GTX 1060) steps/s; CPU 29.79 / 30.98 (csv:15-18)"),
    "beam16x16x76": dict(n=(76, 16, 16), mn=(0, 6, -2), mx=(19, 10, 2), dt=0.04, iters=10, young=1000.0, plane=dict(normal=(0, 1, 0), d=2.0, stiffness=10000.0),
                         published="GPU CudaVec3f 385.3 (RTX 2070) / 291.3 (GTX 1060) steps/s; CPU 9.18 / 8.73 (csv:19-22)"),
}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
for name, c in SCENES.items():
    pos, hexas = T.regular_grid(c["n"], c["mn"], c["mx"])
    hexa = c.get("hexa", False)
    tets = hexas if hexa else T.hexas_to_tetras(hexas, c["n"], "mapping")      # Hexa2TetraTopologicalMapping, swapping off
    fixed = T.box_roi(pos, (-0.1, 5, -3, 0.1, 11, 3))
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f", position=pos)
    FF = sb.HexahedronFEMForceField if hexa else sb.TetrahedronFEMForceField
    ff = FF(mo, tets, youngModulus=c["young"], poissonRatio=0.3, method="large")
    mass = sb.DiagonalMass(mo, tets, totalMass=50.0)
    plane = sb.PlaneForceField(mo, **c["plane"]) if c["plane"] else None
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), plane=plane, dt=c["dt"], gravity=(0.0, -9.0, 0.0), rayleighStiffness=0.1,
                         rayleighMass=0.1, iterations=c["iters"], tolerance=1e-6, threshold=1e-6)
    for _ in range(5):
        node.step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        node.step()
    torch.cuda.synchronize()
    dev_s = time.perf_counter() - t0
    xh = mo.x.detach().cpu().pin_memory(); vh = mo.v.detach().cpu().pin_memory()
    for _ in range(3):
        node.step_host(xh, vh)
    t0 = time.perf_counter()
    for _ in range(steps):
        node.step_host(xh, vh)
    host_s = time.perf_counter() - t0
    info = node.last_solve()
    x = mo.x.cpu().numpy()
    print(json.dumps({"scene": name, "elements": int(tets.shape[0]), "nodes": int(pos.shape[0]), "steps": steps, "steps_per_s_device_resident": steps / dev_s,
                      "steps_per_s_host_buffers": steps / host_s, "cg_iterations_last_step": info["iterations"],
                      "contacts": int(node.get_plane_contacts().sum()) if plane else 0, "min_y": float(x[:, 1].min()), "finite": bool(np.isfinite(x).all()),
                      "published_other_hardware": c["published"]}))
