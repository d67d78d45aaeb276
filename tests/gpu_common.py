"""Shared builders for the GPU parity tests: the same scene on the device (through the C ABI) and in the oracle."""
import numpy as np

import oracle_lib as O

CONFIGS = {
    # C1: examples/Component/SolidMechanics/FEM/TetrahedronFEMForceField.scn (one beam)
    "C1": dict(n=(5, 5, 20), mn=(-5, -5, 0), mx=(5, 5, 40), tess="mapping_swapping", young=1000.0, poisson=0.4, density=0.2,
               gravity=(0.0, -9.0, 0.0), dt=0.01, rK=0.1, rM=0.1, iterations=25, tolerance=1e-9, threshold=1e-9, box=(-6, -6, -1, 50, 6, 0.1)),
    # reference test scene: BaseTetrahedronFEMForceField_test.h:112-168
    "GRID_TEST": dict(n=(4, 10, 4), mn=(0, 0, 20), mx=(10, 40, 30), tess="mapping", young=600.0, poisson=0.3, density=1.0,
                      gravity=(0.0, 10.0, 0.0), dt=0.01, rK=0.0, rM=0.0, iterations=20, tolerance=1e-5, threshold=1e-6, box=(-1, -1, 0, 10, 1, 50)),
    # C2: ~1M-tet cantilever
    "C2": dict(n=(33, 33, 161), mn=(0, 0, 0), mx=(4, 4, 20), tess="mapping_swapping", young=1000.0, poisson=0.3, density=1.0,
               gravity=(0.0, -9.0, 0.0), dt=0.01, rK=0.1, rM=0.1, iterations=25, tolerance=1e-9, threshold=1e-9, box=(-1, -1, -1, 5, 5, 1e-6)),
    "C2_SMALL": dict(n=(9, 9, 41), mn=(0, 0, 0), mx=(4, 4, 20), tess="mapping_swapping", young=1000.0, poisson=0.3, density=1.0,
                     gravity=(0.0, -9.0, 0.0), dt=0.01, rK=0.1, rM=0.1, iterations=25, tolerance=1e-9, threshold=1e-9, box=(-1, -1, -1, 5, 5, 1e-6)),
}
ORC_TESS = {"mapping": 0, "mapping_swapping": 1, "forcefield": 2}


def mesh(cfg):
    from sofa_b200 import topology as T
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    pos, hexas = T.regular_grid(c["n"], c["mn"], c["mx"])
    tets = T.hexas_to_tetras(hexas, c["n"], c["tess"])
    fixed = T.box_roi(pos, c["box"])
    return c, pos, hexas, tets, fixed


def oracle_scene(cfg, dtype, method="large"):
    c, pos, hexas, tets, fixed = mesh(cfg)
    s = O.OracleScene(dtype, pos)
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"],
                 tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_mass_density(c["density"], tets)
    s.set_tets(tets, method, c["young"], c["poisson"])
    s.set_fixed(fixed)
    return s


def gpu_scene(cfg, dtype, method="large", tile_elems=0, ctx=None):
    import sofa_b200 as sb
    c, pos, hexas, tets, fixed = mesh(cfg)
    template = "B200Vec3f" if np.dtype(dtype) == np.float32 else "B200Vec3d"
    ctx = ctx or sb.Context(0)
    mo = sb.MechanicalObject(ctx, template, position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method, tileElems=tile_elems)
    mass = sb.DiagonalMass(mo, tets, massDensity=c["density"])
    fix = sb.FixedProjectiveConstraint(mo, fixed)
    node = sb.SolverNode(mo, ff, mass, fix, dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                         iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    return dict(ctx=ctx, mo=mo, ff=ff, mass=mass, fix=fix, node=node, cfg=c, pos=pos, tets=tets, fixed=fixed)


def dev(mo, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, mo.ndtype)).to(mo.ctx.device)


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
