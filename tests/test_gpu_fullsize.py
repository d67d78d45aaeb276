"""BASELINE.json config C2 at full size (33x33x161 grid, 983 040 tetrahedra): bit-exact against the oracle for the
element passes, plus size-independent properties of the operator."""
import numpy as np
import pytest

from gpu_common import dev, gpu_scene, oracle_scene, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    g = gpu_scene("C2", np.float32)
    s = oracle_scene("C2", np.float32)
    return g, s


def test_c2_sizes(c2):
    g, _ = c2
    assert g["tets"].shape[0] == 983040 and g["pos"].shape[0] == 175329 and len(g["fixed"]) == 1089
    st = g["ff"].stats()
    assert st["interior_nodes"] + st["shared_nodes"] == 175329


def test_c2_force_and_dforce_bit_exact(c2):
    g, s = c2
    rng = np.random.default_rng(0)
    z = g["pos"][:, 2:3]
    x = (g["pos"] + np.hstack([0.02 * np.sin(z / 3.0), 0.05 * (z / 20.0) ** 2, 0 * z]) + 1e-3 * rng.standard_normal(g["pos"].shape)).astype(np.float32)
    zero = np.zeros_like(x)
    f_d = dev(g["mo"], zero); g["ff"].addForce(f_d, dev(g["mo"], x))
    f_ref = s.fem_add_force(zero, x)
    assert f_d.cpu().numpy().tobytes() == f_ref.tobytes()
    dx = (1e-3 * rng.standard_normal(x.shape)).astype(np.float32)
    df_d = dev(g["mo"], zero); g["ff"].addDForce(df_d, dev(g["mo"], dx), -0.0011)
    assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(zero, dx, -0.0011).tobytes()


@pytest.mark.parametrize("dtype,method", [(np.float32, "polar"), (np.float32, "svd"), (np.float64, "large"), (np.float64, "polar")])
def test_c2_other_methods_and_vec3d_bit_exact_at_full_size(dtype, method):
    """The same at full size for the rotation methods and the precision the headline does not use (Vec3d is SOFA's default build): addForce (with the
    cached rotations) and addDForce bit-exact on 983 040 tetrahedra -- this also runs the tile counts / slot budgets big meshes get (Vec3d: 3 tiles per SM)."""
    g = gpu_scene("C2", dtype, method)
    s = oracle_scene("C2", dtype, method)
    rng = np.random.default_rng(2)
    z = g["pos"][:, 2:3]
    x = (g["pos"] + np.hstack([0.02 * np.sin(z / 3.0), 0.05 * (z / 20.0) ** 2, 0 * z]) + 1e-3 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    zero = np.zeros_like(x)
    f_d = dev(g["mo"], zero); g["ff"].addForce(f_d, dev(g["mo"], x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(zero, x).tobytes()
    assert g["ff"].get("rotations").tobytes() == s.get("tet.rotations").tobytes()
    dx = (1e-3 * rng.standard_normal(x.shape)).astype(dtype)
    df_d = dev(g["mo"], zero); g["ff"].addDForce(df_d, dev(g["mo"], dx), -0.0011)
    assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(zero, dx, -0.0011).tobytes()
    # one whole step from the rest state: f and b bit-identical, the CG kernel (streamed tiles in Vec3d) agrees with the oracle's double-dot CG
    s.set_dot_double(True)
    g["node"].step()
    it = g["node"].last_solve()["iterations"]
    it_ref = s.step()
    assert abs(it - it_ref) <= 1
    assert g["node"].get("f").tobytes() == s.get("f").tobytes()
    assert g["node"].get("b").tobytes() == s.get("b").tobytes()
    assert rel_err(g["node"].get("dx"), s.get("sol")) <= (1e-4 if dtype == np.float32 else 1e-9)


def test_c3_hexahedra_at_full_size():
    """BASELINE.json config 3 at full size: 65x65x121 grid, 491 520 hexahedra, method=polar, Vec3f -- addForce / addDForce bit-exact, one whole step with f and b
    bit-identical and the same CG iteration count."""
    import sofa_b200 as sb
    from sofa_b200 import topology as T
    import oracle_lib as O
    pos, hexas = T.regular_grid((65, 65, 121), (0, 0, 0), (8, 8, 15))
    fixed = T.box_roi(pos, (-1, -1, -1, 9, 9, 1e-6))
    assert hexas.shape[0] == 491520
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f", position=pos)
    ff = sb.HexahedronFEMForceField(mo, hexas, youngModulus=1000.0, poissonRatio=0.3, method="polar")
    node = sb.SolverNode(mo, ff, sb.DiagonalMass(mo, hexas, massDensity=1.0), sb.FixedProjectiveConstraint(mo, fixed), dt=0.01, gravity=(0.0, -9.0, 0.0),
                         rayleighStiffness=0.1, rayleighMass=0.1, iterations=25, tolerance=1e-9, threshold=1e-9)
    s = O.OracleScene(np.float32, pos)
    s.set_params(gravity=(0.0, -9.0, 0.0), dt=0.01, rayleighStiffness=0.1, rayleighMass=0.1, iterations=25, tolerance=1e-9, threshold=1e-9)
    s.set_mass_density(1.0, hexas); s.set_hexas(hexas, "polar", 1000.0, 0.3); s.set_fixed(fixed)
    s.set_threads(8)      # (ParallelHexahedronFEMForceField-style addDForce: the same per-vertex summation order as the sequential class)
    rng = np.random.default_rng(4)
    z = pos[:, 2:3]
    x = (pos + np.hstack([0.02 * np.sin(z / 3.0), 0.05 * (z / 15.0) ** 2, 0 * z]) + 1e-3 * rng.standard_normal(pos.shape)).astype(np.float32)
    zero = np.zeros_like(x)
    f_d = dev(mo, zero); ff.addForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(zero, x).tobytes()
    dx = (1e-3 * rng.standard_normal(x.shape)).astype(np.float32)
    df_d = dev(mo, zero); ff.addDForce(df_d, dev(mo, dx), -0.0011)
    assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(zero, dx, -0.0011).tobytes()
    s.set_dot_double(True)     # (the reference's serial float vDot over 1.5 M entries alone moves dx by 8e-3 here: measured)
    node.step()
    it, it_ref = node.last_solve()["iterations"], s.step()
    assert abs(it - it_ref) <= 1
    assert node.get("f").tobytes() == s.get("f").tobytes()
    assert node.get("b").tobytes() == s.get("b").tobytes()
    assert rel_err(node.get("dx"), s.get("sol")) <= 1e-4


def test_c2_operator_properties(c2):
    g, _ = c2
    mo, node = g["mo"], g["node"]
    rng = np.random.default_rng(1)
    p = dev(mo, rng.standard_normal((mo.size, 3))); q = dev(mo, rng.standard_normal((mo.size, 3)))
    p[g["fixed"].astype(np.int64)] = 0; q[g["fixed"].astype(np.int64)] = 0
    Ap, Aq = mo.new_vector(), mo.new_vector()
    m, b, k = 1.001, -0.01, -0.0011
    node.apply(Ap, p, m, b, k); node.apply(Aq, q, m, b, k)
    # symmetry and positive definiteness of the projected system, fixed rows zero, idempotent projection, determinism
    pAq, qAp, pAp = mo.vDot(p, Aq), mo.vDot(q, Ap), mo.vDot(p, Ap)
    assert abs(pAq - qAp) <= 1e-4 * abs(pAp)
    assert pAp > 0
    assert not Ap.cpu().numpy()[g["fixed"]].any()
    Ap2 = mo.new_vector(); node.apply(Ap2, p, m, b, k)
    assert Ap2.cpu().numpy().tobytes() == Ap.cpu().numpy().tobytes()
    # linearity in p (float32: to rounding)
    pq = mo.new_vector(); mo.vOp(pq, p, q, 2.0)
    Apq = mo.new_vector(); node.apply(Apq, pq, m, b, k)
    lin = mo.new_vector(); mo.vOp(lin, Ap, Aq, 2.0)
    assert rel_err(Apq.cpu().numpy(), lin.cpu().numpy()) <= 1e-5


def test_c2_steps_match_oracle(c2):
    """Whole EulerImplicit steps at full size, each from the oracle's own state: f and b bit-identical, iteration count equal,
    CG solution within the Vec3f bound of test_euler_implicit_step_parity_from_same_state."""
    import torch
    g, s = c2
    node, mo = g["node"], g["mo"]
    for step in range(2):
        mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step()
        assert abs(it - it_ref) <= 1
        assert node.get("f").tobytes() == s.get("f").tobytes()
        assert node.get("b").tobytes() == s.get("b").tobytes()
        assert rel_err(node.get("dx"), s.get("sol")) <= 5e-3   # serial float vDot over 526k entries in the reference
    assert np.abs(mo.x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= 1e-5
    # ... and the device CG itself against the oracle with its dot products accumulated in double (the only difference left is the order of the
    # double sums and the one-step rho prediction of the fused kernel): two more steps, each from that oracle's own state
    s.set_dot_double(True)
    for step in range(2):
        mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))
        node.step()
        it_ref = s.step()
        assert abs(node.last_solve()["iterations"] - it_ref) <= 1
        assert rel_err(node.get("dx"), s.get("sol")) <= 2e-5, rel_err(node.get("dx"), s.get("sol"))


def test_c2_per_node_outputs_and_mesh_mass_properties():
    """Full size (983 040 tetrahedra), size-independent properties of the round's late additions: under a rigid motion getRotations gives
    that rotation at every node and computeVonMisesStress gives zero (both strain measures); MeshMatrixMass applied to a constant field
    conserves the total mass (sum of M*1 = rho * volume) and equals the lumped matrix row by row."""
    import sofa_b200 as sb
    import gpu_common
    c, pos, hexas, tets, fixed = gpu_common.mesh("C2")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3d", position=pos)
    a = 0.3
    Q = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    x = pos @ Q.T + np.array([0.5, -1.0, 2.0])
    for how in (1, 2):
        ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method="polar", computeVonMisesStress=how)
        f = mo.new_vector()
        ff.addForce(f, dev(mo, x))
        assert float(f.abs().max()) < 1e-8                                    # no elastic force under a rigid motion
        R = ff.getRotations().cpu().numpy()
        assert np.abs(R - Q).max() < 1e-10
        pe, pn = ff.computeVonMisesStress(dev(mo, x))
        assert float(pe.max()) < 1e-7 and float(pn.max()) < 1e-7
        del ff
    mm = sb.MeshMatrixMass(mo, tets, massDensity=2.5)
    lumped = sb.MeshMatrixMass(mo, tets, massDensity=2.5, lumping=True)
    ones = dev(mo, np.ones_like(pos))
    r1 = mo.new_vector(); mm.addMDx(r1, ones, 1.0)
    r2 = mo.new_vector(); lumped.addMDx(r2, ones, 1.0)
    total = 2.5 * 4.0 * 4.0 * 20.0                                             # density * volume of the beam
    assert abs(float(r1[:, 0].sum()) - total) < 1e-8 * total
    assert float((r1 - r2).abs().max()) < 1e-12                                # the lumped matrix is the row sum of the sparse one (2.5 x vertex mass)


@pytest.mark.parametrize("dtype,method", [(np.float32, "qr"), (np.float64, "polar")])
def test_c2_fast_tetrahedral_corotational_bit_exact_at_full_size(dtype, method):
    """FastTetrahedralCorotationalForceField on C2's mesh (983 040 tetrahedra, 1 180 896 edges): addForce, the per-edge matrices and addDForce over
    the edges bit-exact against the oracle's restatement, then one whole step (the fused CG kernel with the edge pass)."""
    import sofa_b200 as sb
    import oracle_lib as O
    from gpu_common import mesh
    c, pos, hexas, tets, fixed = mesh("C2")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.FastTetrahedralCorotationalForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method)
    s = O.OracleScene(dtype, pos)
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_mass_density(c["density"], tets); s.set_fast_tets(tets, method, c["young"], c["poisson"]); s.set_fixed(fixed)
    assert ff.get("n_edges") == 1180896 and ff.get("edges").astype(np.int64).tobytes() == s.get("fast.edges").tobytes()
    rng = np.random.default_rng(3)
    z = pos[:, 2:3]
    x = (pos + np.hstack([0.02 * np.sin(z / 3.0), 0.05 * (z / 20.0) ** 2, 0 * z]) + 1e-3 * rng.standard_normal(pos.shape)).astype(dtype)
    zero = np.zeros_like(x)
    f_d = dev(mo, zero); ff.addForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(zero, x).tobytes()
    assert ff.get("rotations").tobytes() == s.get("fast.rotations").tobytes()
    dx = (1e-3 * rng.standard_normal(x.shape)).astype(dtype)
    df_d = dev(mo, zero); ff.addDForce(df_d, dev(mo, dx), -0.0011)
    assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(zero, dx, -0.0011).tobytes()
    assert ff.get("edgeInfo").tobytes() == s.get("fast.edgeInfo").tobytes()
    mass = sb.DiagonalMass(mo, tets, massDensity=c["density"])
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                         iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_dot_double(True)
    node.step()
    it_ref = s.step()
    assert node.get("b").tobytes() == s.get("b").tobytes()
    assert abs(node.last_solve()["iterations"] - it_ref) <= 1
    assert rel_err(node.get("dx"), s.get("sol")) <= (1e-9 if dtype == np.float64 else 1e-4)
    assert bool(node.fused_info()["fused_enabled"])
