#!/usr/bin/env python
"""Generates the committed golden fixtures (run in the build container, where /root/reference exists):

  liver_mesh.npz          config C4 input: /root/reference/share/mesh/liver.msh (Gmsh v1, 181 nodes, 596 tetrahedra) read with the
                          product's reader sofa_b200.topology.read_gmsh_v1 -- the GPU box has no /root/reference
  c1_large_{f32,f64}.npz  config C1 (examples/Component/SolidMechanics/FEM/TetrahedronFEMForceField.scn, one beam, method=large)
  c4_polar_{f32,f64}.npz  config C4 (Demos/liver.scn with TetrahedronFEMForceField method=polar, fixed 3 39 64, no collision)

Each trajectory fixture holds, for 3 consecutive EulerImplicit steps of the ORACLE (oracle/sofa_oracle.hpp, reference
summation order), the state before the step (x, v) and the step's force vector f, right-hand side b, CG solution dx and
CG iteration count.  tests/test_gpu_golden.py replays them on the device."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402
from gpu_common import CONFIGS, mesh  # noqa: E402
from sofa_b200 import topology as T  # noqa: E402

C4 = dict(young=3000.0, poisson=0.3, density=1.0, gravity=(0.0, -9.81, 0.0), dt=0.02, rK=0.1, rM=0.1, iterations=25, tolerance=1e-9, threshold=1e-9,
          fixed=np.array([3, 39, 64], np.uint32))


def trajectory(s, steps=3):
    out = {}
    for k in range(steps):
        out[f"x{k}"] = s.get("x").copy(); out[f"v{k}"] = s.get("v").copy()
        out[f"it{k}"] = np.array(s.step())
        out[f"f{k}"] = s.get("f").copy(); out[f"b{k}"] = s.get("b").copy(); out[f"dx{k}"] = s.get("sol").copy()
    out["x_end"] = s.get("x").copy(); out["v_end"] = s.get("v").copy()
    return out


def main():
    pos, tets, hexas = T.read_gmsh_v1("/root/reference/share/mesh/liver.msh")
    assert pos.shape == (181, 3) and tets.shape == (596, 4)
    np.savez_compressed(os.path.join(HERE, "liver_mesh.npz"), positions=pos, tetrahedra=tets)
    for name, dtype in (("f32", np.float32), ("f64", np.float64)):
        c, p1, _, t1, fixed1 = mesh("C1")
        s = O.OracleScene(dtype, p1)
        s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
        s.set_mass_density(c["density"], t1); s.set_tets(t1, "large", c["young"], c["poisson"]); s.set_fixed(fixed1)
        np.savez_compressed(os.path.join(HERE, f"c1_large_{name}.npz"), **trajectory(s))
        s = O.OracleScene(dtype, pos)
        s.set_params(gravity=C4["gravity"], dt=C4["dt"], rayleighStiffness=C4["rK"], rayleighMass=C4["rM"], iterations=C4["iterations"], tolerance=C4["tolerance"], threshold=C4["threshold"])
        s.set_mass_density(C4["density"], tets); s.set_tets(tets, "polar", C4["young"], C4["poisson"]); s.set_fixed(C4["fixed"])
        np.savez_compressed(os.path.join(HERE, f"c4_polar_{name}.npz"), **trajectory(s))
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
