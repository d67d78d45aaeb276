"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bar (BASELINE.json north_star): force vectors <= 1e-5 relative, CG iteration count within +-1.
Because the kernels repeat the reference's operations in the reference's order with contraction off,
addForce / addDForce / A*p / the right-hand side are in fact required to be BIT-IDENTICAL here; only quantities
downstream of a dot product (the reference accumulates vDot serially in Real) carry a tolerance."""
import numpy as np
import pytest

import oracle_lib as O
import gpu_common
from gpu_common import dev, gpu_scene, oracle_scene, rel_err

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64]


@pytest.mark.parametrize("dtype", DTYPES)
def test_vop_cases_bit_exact(dtype):
    """MechanicalObjectVOp_test.cpp:770-979 cases."""
    g = gpu_scene("C1", dtype)
    mo = g["mo"]
    rng = np.random.default_rng(0)
    a = rng.standard_normal((mo.size, 3)).astype(dtype); b = rng.standard_normal((mo.size, 3)).astype(dtype); r0 = rng.standard_normal((mo.size, 3)).astype(dtype)
    k = 0.37
    cases = [("r=0", None, None), ("r*=k", None, "r"), ("r=b*k", None, "b"), ("r=a", "a", None), ("r+=b*k", "r", "b"), ("r=a+r*k", "a", "r"), ("r=a+b*k", "a", "b")]
    for name, ka, kb in cases:
        for kk in (k, 1.0):
            r_ref = r0.copy()
            pick = lambda key, r: None if key is None else (r if key == "r" else (a if key == "a" else b))
            O.vop(dtype, r_ref, pick(ka, r_ref), pick(kb, r_ref), kk)
            r_d, a_d, b_d = dev(mo, r0), dev(mo, a), dev(mo, b)
            pd = lambda key: None if key is None else (r_d if key == "r" else (a_d if key == "a" else b_d))
            mo.vOp(r_d, pd(ka), pd(kb), kk)
            assert r_d.cpu().numpy().tobytes() == r_ref.tobytes(), (name, kk)
    d = mo.vDot(dev(mo, a), dev(mo, b))
    exact = float(np.sum(a.astype(np.float64) * b.astype(np.float64)))
    assert abs(d - exact) <= 1e-12 * abs(exact) + 1e-12
    assert abs(d - O.vdot(dtype, a, b)) <= (1e-12 if dtype == np.float64 else 2e-4) * max(1.0, abs(exact))
    # integration fast path
    v_d, x_d, a_d = dev(mo, r0), dev(mo, a), dev(mo, b)
    mo.vMultiOp_integrate(v_d, x_d, a_d, 1.0, 0.01)
    v_ref = r0 + b; x_ref = a + v_ref * dtype(0.01)
    assert v_d.cpu().numpy().tobytes() == v_ref.tobytes() and x_d.cpu().numpy().tobytes() == x_ref.tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["small", "large", "polar", "svd"])
def test_add_force_add_dforce_bit_exact(dtype, method):
    g = gpu_scene("C1", dtype, method)
    s = oracle_scene("C1", dtype, method)
    mo, ff = g["mo"], g["ff"]
    rng = np.random.default_rng(1)
    x = (g["pos"] + 0.3 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    f0 = rng.standard_normal(x.shape).astype(dtype)
    f_d = dev(mo, f0)
    ff.addForce(f_d, dev(mo, x))
    f_ref = s.fem_add_force(f0, x)
    assert f_d.cpu().numpy().tobytes() == f_ref.tobytes()
    assert rel_err(f_d.cpu().numpy(), f_ref) <= 1e-5
    if method != "small":
        assert ff.get("rotations").tobytes() == s.get("tet.rotations").tobytes()
        assert ff.get("initialRotations").tobytes() == s.get("tet.initialRotations").tobytes()
    assert ff.get("strainDisplacements").tobytes() == s.get("tet.J").tobytes()
    assert ff.get("materialsStiffnesses").tobytes() == s.get("tet.K").tobytes()
    dx = rng.standard_normal(x.shape).astype(dtype)
    for kf in (1.0, -0.0011, 0.11):
        df_d = dev(mo, f0)
        ff.addDForce(df_d, dev(mo, dx), kf)
        assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, kf).tobytes(), kf


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["small", "large", "polar", "svd"])
@pytest.mark.parametrize("plastic", [(3.0, 1.0, 0.5), (4.0, 6.0, 0.9)], ids=["creep_everywhere", "yield_gate"])
def test_plasticity_branch_bit_exact(dtype, method, plastic):
    """computeForce with plasticMaxThreshold > 0 (TetrahedronFEMForceField.inl:357-371): the per-element plastic strain evolves over
    successive addForce calls; forces and strains must equal the oracle's bit for bit, addDForce is untouched, reset() clears."""
    import sofa_b200 as sb
    c, pos, hexas, tets, fixed = gpu_common.mesh("C1")
    s = oracle_scene("C1", dtype, method)
    s.set_plastic(*plastic)
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method, plasticMaxThreshold=plastic[0],
                                     plasticYieldThreshold=plastic[1], plasticCreep=plastic[2])
    rng = np.random.default_rng(11)
    f0 = rng.standard_normal(pos.shape).astype(dtype)
    for it in range(3):
        x = (pos + (0.3 - 0.1 * it) * rng.standard_normal(pos.shape)).astype(dtype)
        f_d = dev(mo, f0)
        ff.addForce(f_d, dev(mo, x))
        assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes(), it
        assert ff.get("plasticStrains").tobytes() == s.get("tet.plasticStrains").tobytes(), it
    n = np.linalg.norm(ff.get("plasticStrains").astype(np.float64), axis=1)
    assert (n > 0).any() and n.max() <= plastic[0] * (1 + 1e-5)
    dx = rng.standard_normal(pos.shape).astype(dtype)
    df_d = dev(mo, f0)
    ff.addDForce(df_d, dev(mo, dx), 0.11)
    assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, 0.11).tobytes()
    # a permanent set remains once the load is gone: at the rest shape the force is not zero any more
    f_d = dev(mo, np.zeros_like(f0))
    ff.addForce(f_d, dev(mo, pos.astype(dtype)))
    f_ref = s.fem_add_force(np.zeros_like(f0), pos.astype(dtype))
    assert f_d.cpu().numpy().tobytes() == f_ref.tobytes() and np.abs(f_ref).max() > 0
    ff.reset(); s.tet_reset()
    assert np.abs(ff.get("plasticStrains")).max() == 0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["small", "large", "polar"])
@pytest.mark.parametrize("update", [False, True], ids=["fixedJ", "updateStiffnessMatrix"])
@pytest.mark.parametrize("meshname", ["C1", "liver"])
def test_tetrahedral_corotational_fem_force_field(dtype, method, update, meshname):
    """TetrahedralCorotationalFEMForceField (what Demos/liver.scn uses; the reference's own tests hold it to the SAME golden vectors as
    TetrahedronFEMForceField, TetrahedralCorotationalFEMForceField_test.cpp:30-82): addForce / addDForce bit-identical to the oracle, also
    with updateStiffnessMatrix (for `large` the class rewrites all three copies of a cofactor, .inl:920-937); without it the results equal
    TetrahedronFEMForceField's bit for bit."""
    import os
    import sofa_b200 as sb
    if meshname == "C1":
        c, pos, hexas, tets, fixed = gpu_common.mesh("C1")
        young, poisson, amp = c["young"], c["poisson"], 0.2
    else:
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "liver_mesh.npz"))
        pos, tets, young, poisson, amp = z["positions"], z["tetrahedra"], 3000.0, 0.3, 0.05      # Demos/liver.scn:35
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedralCorotationalFEMForceField(mo, tets, youngModulus=young, poissonRatio=poisson, method=method, updateStiffnessMatrix=update)
    twin = sb.TetrahedronFEMForceField(mo, tets, youngModulus=young, poissonRatio=poisson, method=method)
    s = O.OracleScene(dtype, pos); s.set_tets(tets, method, young, poisson)
    s.set_tetrahedral_corotational(True); s.set_update_stiffness_matrix(update)
    rng = np.random.default_rng(41)
    f0 = rng.standard_normal(pos.shape).astype(dtype)
    for it in range(3):
        x = (pos + amp * rng.standard_normal(pos.shape)).astype(dtype)
        f_d = dev(mo, f0); ff.addForce(f_d, dev(mo, x))
        assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes(), it
        dx = rng.standard_normal(pos.shape).astype(dtype)
        df_d = dev(mo, f0); ff.addDForce(df_d, dev(mo, dx), -0.0011)
        assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, -0.0011).tobytes(), it
        if not update:
            t_d = dev(mo, f0); twin.addForce(t_d, dev(mo, x))
            assert t_d.cpu().numpy().tobytes() == f_d.cpu().numpy().tobytes()
        if method != "small":   # the class's own getRotation (.inl:779-820): rotation * initialTransformation averaged, Gram-Schmidt
            f_d = dev(mo, f0); ff.addForce(f_d, dev(mo, x)); s.fem_add_force(f0, x)      # (addDForce does not touch the rotations; same state on both sides)
            assert ff.getRotations().cpu().numpy().tobytes() == s.tet_get_rotations().tobytes(), it
    if method == "small":
        with pytest.raises(sb.Sofab200Error):
            ff.getRotations()
    with pytest.raises((ValueError, sb.Sofab200Error)):
        sb.TetrahedralCorotationalFEMForceField(mo, tets, method="svd")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["polar", "svd", "small", "large"])
def test_update_stiffness_matrix(dtype, method):
    """Data updateStiffnessMatrix (TetrahedronFEMForceField.inl:1063-1067,1174-1177): polar / svd recompute the strain-displacement terms
    from the deformed element in every addForce, and addDForce then uses them; `small` ignores the flag; `large` rewrites nine single entries
    of J, all in its normal-strain columns (:908-922), so the copies of those cofactors in the shear columns keep their initial values."""
    import sofa_b200 as sb
    c, pos, hexas, tets, fixed = gpu_common.mesh("C1")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method, updateStiffnessMatrix=True)
    s = oracle_scene("C1", dtype, method)
    s.set_update_stiffness_matrix(True)
    rng = np.random.default_rng(31)
    f0 = rng.standard_normal(pos.shape).astype(dtype)
    J0 = ff.get("strainDisplacements").copy()
    for it in range(3):
        x = (pos + 0.2 * rng.standard_normal(pos.shape)).astype(dtype)
        f_d = dev(mo, f0)
        ff.addForce(f_d, dev(mo, x))
        assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes(), it
        dx = rng.standard_normal(pos.shape).astype(dtype)
        df_d = dev(mo, f0)
        ff.addDForce(df_d, dev(mo, dx), 0.11)
        assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, 0.11).tobytes(), it
    if method != "small":
        assert np.abs(s.get("tet.J") - J0).max() > 0      # the oracle's J did change
    if method == "large":
        assert np.abs(s.get("tet.J") - s.get("tet.Jsh")).max() > 0      # ... and only its normal-strain columns


@pytest.mark.parametrize("dtype", DTYPES)
def test_update_stiffness_matrix_large_steps(dtype):
    """EulerImplicit + CG steps with method large and updateStiffnessMatrix (two cofactor sets per element: served by the multi-kernel CG loop)."""
    import sofa_b200 as sb
    c, pos, hexas, tets, fixed = gpu_common.mesh("C1")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method="large", updateStiffnessMatrix=True)
    mass = sb.DiagonalMass(mo, tets, massDensity=c["density"])
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                         iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s = oracle_scene("C1", dtype, "large")
    s.set_update_stiffness_matrix(True)
    s.set_dot_double(True)
    ref = oracle_scene("C1", dtype, "large"); ref.set_dot_double(True)      # without the flag: must differ
    for it in range(4):
        n_it = node.step(); s_it = s.step(); ref.step()
        assert abs(node.last_solve()["iterations"] - s_it) <= 1
        assert node.get("b").tobytes() == s.get("b").tobytes() or it > 0      # (first step: same state on both sides => bit-exact right-hand side)
        # (the flag makes J, hence the operator, non-symmetric: CG amplifies the last-bit differences between the two dot-product orders about 20x per step --
        #  measured 7e-13, 2e-9, 2e-8, 8e-8 in Vec3d; Vec3f with the oracle's dots in double comes out bit-identical)
        assert np.abs(mo.x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= (1e-6 if dtype == np.float64 else 2e-4)
    assert np.abs(s.get("x") - ref.get("x")).max() > 1e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["small", "large", "polar", "svd"])
@pytest.mark.parametrize("how", [1, 2])
def test_von_mises_stress_bit_exact(dtype, method, how):
    """computeVonMisesStress (TetrahedronFEMForceField.inl:2196-2372): per element and per node, both strain measures."""
    import sofa_b200 as sb
    c, pos, hexas, tets, fixed = gpu_common.mesh("C1")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method, computeVonMisesStress=how)
    rng = np.random.default_rng(21)
    x = (pos + 0.2 * rng.standard_normal(pos.shape)).astype(dtype)
    if method == "small" and how == 1:
        with pytest.raises(sb.Sofab200Error):
            ff.computeVonMisesStress(dev(mo, x))
        return
    s = oracle_scene("C1", dtype, method)
    pe_ref, pn_ref = s.tet_von_mises(x, how)
    pe, pn = ff.computeVonMisesStress(dev(mo, x))
    assert pe.cpu().numpy().tobytes() == pe_ref.tobytes()
    assert pn.cpu().numpy().tobytes() == pn_ref.tobytes()
    assert pe_ref.max() > 0
    if how == 1:
        assert ff.get("rotations").tobytes() == s.get("tet.rotations").tobytes()     # method 1 rewrites rotations[e], as the reference does
    # no stress in the rest configuration
    pe0, _ = ff.computeVonMisesStress(dev(mo, pos.astype(dtype)))
    assert float(pe0.max()) <= (1e-2 if dtype == np.float32 else 1e-9)
    off = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method)
    with pytest.raises(sb.Sofab200Error):
        off.computeVonMisesStress(dev(mo, x))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["small", "large", "polar", "svd"])
def test_get_rotations_per_node_bit_exact(dtype, method):
    """getRotations(VecReal&) (TetrahedronFEMForceField.inl:781-833): mean of rotations[t] * R0(t) around each node + polar."""
    g = gpu_scene("C1", dtype, method)
    s = oracle_scene("C1", dtype, method)
    mo, ff = g["mo"], g["ff"]
    rng = np.random.default_rng(5)
    for amp in (0.0, 0.3):
        x = (g["pos"] + amp * rng.standard_normal(g["pos"].shape)).astype(dtype)
        f0 = np.zeros_like(x)
        ff.addForce(dev(mo, f0), dev(mo, x))
        s.fem_add_force(f0, x)
        R_d = ff.getRotations().cpu().numpy()
        R_ref = s.tet_get_rotations()
        assert R_d.tobytes() == R_ref.tobytes(), (method, amp)
        if method != "small":
            eye = np.einsum("nij,nkj->nik", R_d.astype(np.float64), R_d.astype(np.float64))
            assert np.abs(eye - np.eye(3)).max() < (1e-5 if dtype == np.float32 else 1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [dict(), dict(maxForce=0.05), dict(bilateral=True, normal=(0.3, 1.0, -0.2), d=0.7)], ids=["default", "maxForce", "bilateral"])
def test_plane_force_field_bit_exact(dtype, case):
    """PlaneForceField addForce / addDForce (per-operation level; in every SofaCUDA FEM benchmark scene), bit-identical to the
    restatement of PlaneForceField.inl:139-226, including the contact set."""
    import sofa_b200 as sb
    from gpu_common import mesh
    c, pos, hexas, tets, fixed = mesh("C1")
    rng = np.random.default_rng(5)
    x = (pos + 0.3 * rng.standard_normal(pos.shape)).astype(dtype)
    x[:, 1] -= 4.0                                   # part of the beam below the plane y = d
    v = rng.standard_normal(pos.shape).astype(dtype)
    kw = dict(normal=(0.0, 2.0, 0.0), d=-1.0, stiffness=500.0, damping=5.0, maxForce=0.0, bilateral=False); kw.update(case)
    prm = list(kw["normal"]) + [kw["d"], kw["stiffness"], kw["damping"], kw["maxForce"], float(kw["bilateral"])]
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    pf = sb.PlaneForceField(mo, rayleighStiffness=0.1, **kw)
    f0 = rng.standard_normal(pos.shape).astype(dtype)
    f_d = dev(mo, f0); pf.addForce(f_d, dev(mo, x), dev(mo, v))
    f_ref, c_ref = O.plane_add_force(dtype, prm, f0, x, v)
    assert 0 < c_ref.sum() <= pos.shape[0]
    assert pf.contacts.cpu().numpy().tobytes() == c_ref.tobytes()
    assert f_d.cpu().numpy().tobytes() == f_ref.tobytes()
    dx = (1e-2 * rng.standard_normal(pos.shape)).astype(dtype)
    df_d = dev(mo, f0); pf.addDForce(df_d, dev(mo, dx), kFactor=-0.0011, bFactor=-0.01)
    assert df_d.cpu().numpy().tobytes() == O.plane_add_dforce(dtype, prm, f0, dx, c_ref, -0.0011 + -0.01 * 0.1).tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_solver_node_with_plane_force_field(dtype):
    """The SofaCUDA benchmark scenes' node: mass, TetrahedronFEMForceField, FixedConstraint, PlaneForceField.  The plane's addForce /
    addDForce are fused into the epilogue of the element passes: f, b and A*p bit-identical, steps as in the plain test."""
    import sofa_b200 as sb
    import torch
    from gpu_common import mesh
    c, pos, hexas, tets, fixed = mesh("C1")
    kw = dict(normal=(0.0, 1.0, 0.0), d=-4.0, stiffness=200.0, damping=1.0)       # the beam (y in [-5, 5]) dips below y = -4
    prm = list(kw["normal"]) + [kw["d"], kw["stiffness"], kw["damping"], 0.0, 0.0]
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method="large")
    mass = sb.DiagonalMass(mo, tets, massDensity=c["density"])
    plane = sb.PlaneForceField(mo, rayleighStiffness=0.05, **kw)
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), plane=plane, dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"],
                         rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s = oracle_scene("C1", dtype, "large")
    s.set_plane(prm, 0.05)
    rng = np.random.default_rng(8)
    p = rng.standard_normal(pos.shape).astype(dtype)
    for step in range(5):
        mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step()
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1
        assert rel_err(node.get("dx"), s.get("sol")) <= (1e-8 if dtype == np.float64 else 2e-4), step
        q_d = mo.new_vector(); node.apply(q_d, dev(mo, p), 1.001, -0.01, -0.0011)      # contacts of this step's addForce
        assert q_d.cpu().numpy().tobytes() == s.apply(p, 1.001, -0.01, -0.0011).tobytes(), step
    assert int(plane_contacts_of(node).sum()) > 0


def plane_contacts_of(node):
    return node.get_plane_contacts()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("how", ["vertexMass", "totalMass"])
def test_uniform_mass_parity(dtype, how):
    """UniformMass instead of DiagonalMass ("UniformMass equivalents"): its addMDx multiplies the MassType by the
    factor first (UniformMass.inl:414-419), its addForce adds one precomputed weight vector (:484-496).  Per-op entry points and the
    fused solver node: f, b and A*p bit-identical, steps as in the DiagonalMass test."""
    import sofa_b200 as sb
    from gpu_common import mesh
    c, pos, hexas, tets, fixed = mesh("C1")
    kw = dict(vertexMass=0.37) if how == "vertexMass" else dict(totalMass=123.4)
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method="large")
    mass = sb.UniformMass(mo, **kw)
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                         iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s = O.OracleScene(dtype, pos)
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_uniform_mass(**kw); s.set_tets(tets, "large", c["young"], c["poisson"]); s.set_fixed(fixed)
    assert mass.vertexMass_host.tobytes() == s.get("vertexMass").tobytes()
    rng = np.random.default_rng(3)
    # per-op entry points against the fused operator with the stiffness switched off: A = m M
    p = rng.standard_normal(pos.shape).astype(dtype)
    for fac in (1.0, 1.001, -0.3):
        r0 = rng.standard_normal(pos.shape).astype(dtype)
        r_d = dev(mo, r0); mass.addMDx(r_d, dev(mo, p), fac)
        m = np.asarray(mass.vertexMass_value, dtype)
        if fac != 1.0:
            m = m * dtype(fac)
        assert r_d.cpu().numpy().tobytes() == (r0 + p * m).astype(dtype).tobytes()
    g = g0 = rng.standard_normal(pos.shape).astype(dtype)
    g_d = dev(mo, g0); mass.addForce(g_d, c["gravity"])
    mg = np.array([dtype(v) * dtype(mass.vertexMass_value) for v in c["gravity"]], dtype)
    assert g_d.cpu().numpy().tobytes() == (g0 + mg).astype(dtype).tobytes()
    for (mf, bf, kf) in ((1.001, -0.01, -0.0011), (1.0, 0.0, -0.01)):
        q_d = mo.new_vector(); node.apply(q_d, dev(mo, p), mf, bf, kf)
        assert q_d.cpu().numpy().tobytes() == s.apply(p, mf, bf, kf).tobytes(), (mf, bf, kf)
    for step in range(4):
        import torch
        mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step()
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1
        assert rel_err(node.get("dx"), s.get("sol")) <= (1e-8 if dtype == np.float64 else 2e-4), step


@pytest.mark.parametrize("dtype", DTYPES)
def test_per_element_material_data_bit_exact(dtype):
    """youngModulus / poissonRatio given per element (getYoungModulusInElement, BaseLinearElasticityFEMForceField.inl:125-139) and
    localStiffnessFactor (TetrahedronFEMForceField.inl:261): the material values are computed on the host in the reference's arithmetic."""
    import sofa_b200 as sb
    from gpu_common import mesh
    c, pos, hexas, tets, fixed = mesh("C1")
    rng = np.random.default_rng(11)
    young = rng.uniform(500.0, 5000.0, tets.shape[0]); poisson = rng.uniform(0.1, 0.45, tets.shape[0])
    lsf = np.array([0.5, 1.0, 2.0, 1.5])
    x = (pos + 0.05 * rng.standard_normal(pos.shape)).astype(dtype)
    dx = (1e-3 * rng.standard_normal(pos.shape)).astype(dtype)
    for kw_dev, kw_ref in ((dict(youngModulus=young, poissonRatio=poisson), dict(young=young, poisson=poisson)),
                           (dict(youngModulus=1000.0, poissonRatio=0.4, localStiffnessFactor=lsf), dict(young=1000.0, poisson=0.4, local_stiffness_factor=lsf))):
        ctx = sb.Context(0)
        mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
        ff = sb.TetrahedronFEMForceField(mo, tets, method="polar", **kw_dev)
        s = O.OracleScene(dtype, pos); s.set_tets(tets, "polar", **kw_ref)
        f0 = rng.standard_normal(pos.shape).astype(dtype)
        f_d = dev(mo, f0); ff.addForce(f_d, dev(mo, x))
        assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes()
        df_d = dev(mo, f0); ff.addDForce(df_d, dev(mo, dx), -0.37)
        assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, -0.37).tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_svd_inverted_elements(dtype):
    g = gpu_scene("C1", dtype, "svd")
    s = oracle_scene("C1", dtype, "svd")
    rng = np.random.default_rng(2)
    x = g["pos"].copy(); x[::5] += 2.0 * rng.standard_normal(x[::5].shape)   # crushed and inverted elements
    x = x.astype(dtype)
    z = np.zeros_like(x)
    f_d = dev(g["mo"], z)
    g["ff"].addForce(f_d, dev(g["mo"], x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(z, x).tobytes()
    assert g["ff"].get("rotations").tobytes() == s.get("tet.rotations").tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tile", [256, 1024, 2048])
def test_result_is_independent_of_the_tiling(dtype, tile):
    g = gpu_scene("C2_SMALL", dtype, "large", tile_elems=tile)
    s = oracle_scene("C2_SMALL", dtype, "large")
    rng = np.random.default_rng(3)
    x = (g["pos"] + 0.02 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    z = np.zeros_like(x)
    f_d = dev(g["mo"], z); g["ff"].addForce(f_d, dev(g["mo"], x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(z, x).tobytes()
    dx = rng.standard_normal(x.shape).astype(dtype)
    df_d = dev(g["mo"], z); g["ff"].addDForce(df_d, dev(g["mo"], dx), -0.0011)
    assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(z, dx, -0.0011).tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_mass_and_constraint_ops(dtype):
    g = gpu_scene("C1", dtype)
    mo, mass, fix = g["mo"], g["mass"], g["fix"]
    rng = np.random.default_rng(4)
    res = rng.standard_normal((mo.size, 3)).astype(dtype); dx = rng.standard_normal((mo.size, 3)).astype(dtype)
    m = mass.vertexMass_host
    for factor in (1.0, 1.001, -0.1):
        r_d = dev(mo, res); mass.addMDx(r_d, dev(mo, dx), factor)
        ref = res + (dx * m[:, None]) * dtype(factor) if factor != 1.0 else res + dx * m[:, None]
        assert r_d.cpu().numpy().tobytes() == ref.astype(dtype).tobytes()
    f_d = dev(mo, res); mass.addForce(f_d, (0.0, -9.0, 0.5))
    grav = np.array([0.0, -9.0, 0.5], dtype)
    assert f_d.cpu().numpy().tobytes() == (res + grav[None, :] * m[:, None]).astype(dtype).tobytes()
    r_d = dev(mo, res); fix.projectResponse(r_d)
    ref = res.copy(); ref[g["fixed"]] = 0
    assert r_d.cpu().numpy().tobytes() == ref.tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["large", "polar"])
def test_node_compute_force_and_apply_bit_exact(dtype, method):
    g = gpu_scene("C1", dtype, method)
    s = oracle_scene("C1", dtype, method)
    mo, node = g["mo"], g["node"]
    rng = np.random.default_rng(5)
    x = (g["pos"] + 0.2 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    s.set_x(x)
    f_d = mo.new_vector(); node.computeForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.compute_force().tobytes()
    p = rng.standard_normal(x.shape).astype(dtype)
    for (m, b, k) in ((1.001, -0.01, -0.0011), (1.0, 0.0, -0.01), (0.0, 0.0, 0.11)):
        q_d = mo.new_vector(); node.apply(q_d, dev(mo, p), m, b, k)
        assert q_d.cpu().numpy().tobytes() == s.apply(p, m, b, k).tobytes(), (m, b, k)


# the three device implementations of CGLinearSolver::solve (node.cu: Node::cg_solve): ONE persistent cooperative kernel; element
# pass + cooperative tail kernel per iteration; four plain kernels per iteration (the latter two serve meshes the first cannot hold)
# "fused" (default, cg_fused.cuh): ONE reduction per iteration -- alpha from the measured rho, beta from the one-step prediction
# rho' = rho - 2 alpha r.q + alpha^2 q.q -- so its scalars are not bit-equal to a classical CG's; in Vec3f it stays closer to the
# double-dot oracle than the reference's own serial-float dots do (tests/test_cg_fused_recurrence.py measures both).
CG_PATHS = {"fused": {}, "fused_all_warps": {"SOFAB200_FUSED_GATHER_WARPS": "0"}, "fused_streamed": {"SOFAB200_FUSED_CACHED_KB": "0"},
            "persistent_v1": {"SOFAB200_CG_FUSED": "0"}, "fused_tail": {"SOFAB200_CG_PERSISTENT": "0"},
            "multi_kernel": {"SOFAB200_CG_PERSISTENT": "0", "SOFAB200_FUSED_TAIL": "0"}}


@pytest.mark.parametrize("path", list(CG_PATHS))
@pytest.mark.parametrize("dtype", DTYPES)
def test_cg_solve_matches_oracle(dtype, path, monkeypatch):
    for k, v in CG_PATHS[path].items():
        monkeypatch.setenv(k, v)          # read when the solver node is created
    g = gpu_scene("C1", dtype)
    s = oracle_scene("C1", dtype)
    mo, node = g["mo"], g["node"]
    rng = np.random.default_rng(6)
    x = (g["pos"] + 0.05 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    s.fem_add_force(np.zeros_like(x), x); z = mo.new_vector(); g["ff"].addForce(z, dev(mo, x))   # cache rotations at x on both sides
    b = rng.standard_normal(x.shape).astype(dtype); b[g["fixed"]] = 0
    m, bf, k = 1.001, -0.01, -0.0011
    for iters, tol in ((25, 1e-9), (100, 1e-4), (1, 1e-9), (2, 1e-9)):   # (1 and 2: the solve ends inside its first iterations)
        node.set_params(iterations=iters, tolerance=tol, threshold=1e-9)
        sol_d = mo.new_vector()
        it = node.cg_solve(sol_d, dev(mo, b), m, bf, k)
        info = node.last_solve()
        for dd in (False, True):   # reference-order dots, then double-accumulated dots (see oracle Scene::dotDouble)
            s.set_dot_double(dd)
            s.set_params(iterations=iters, tolerance=tol, threshold=1e-9)
            sol_ref, it_ref = s.cg(b, m, bf, k)
            assert abs(it - it_ref) <= 1, (it, it_ref, dd)
            ge_ref = s.graph("Error")
            nmin = min(len(ge_ref), len(info["graph_error"]), 26 if not dd else 10 ** 6)
            rtol = 1e-7 if (dd or dtype == np.float64) else 2e-3
            sol_tol = 1e-8 if (dd or dtype == np.float64) else 2e-3
            if path.startswith("fused") and path != "fused_tail" and dtype == np.float32 and dd:
                # (Vec3f: the predicted rho enters beta at ~1e-7 relative; near the Vec3f floor the residual histories of ANY two float CGs part)
                rtol, sol_tol, nmin = 5e-6, 5e-6, min(nmin, 26)
            assert np.allclose(info["graph_error"][:nmin], ge_ref[:nmin], rtol=rtol, atol=1e-14), dd
            if it == it_ref:
                assert rel_err(sol_d.cpu().numpy(), sol_ref) <= sol_tol
                assert info["end_condition"] == s.end_condition
    # b == 0: the reference returns at once with x = 0 (CGLinearSolver.inl:141-152)
    node.set_params(iterations=25, tolerance=1e-9, threshold=1e-9)
    sol_d = mo.new_vector(); sol_d.fill_(7.0)
    it = node.cg_solve(sol_d, dev(mo, np.zeros_like(b)), m, bf, k)
    assert it == 0 and float(sol_d.abs().max()) == 0.0


def _sync_state(g, s):
    """Copy the oracle's (x, v) into the device MechanicalObject."""
    import torch
    mo = g["mo"]
    mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["large", "polar", "svd", "small"])
def test_euler_implicit_step_parity_from_same_state(dtype, method):
    """One EulerImplicitSolver::solve from the reference's own state, 10 consecutive steps of its trajectory:
       * force vector f and right-hand side b: BIT-IDENTICAL (north_star bar: <= 1e-5 relative);
       * CG iteration count within +-1 (here: equal);
       * CG solution dx: <= 1e-8 relative (Vec3d), <= 2e-4 (Vec3f).  dx is downstream of vDot, which the reference
         sums serially in Real (MechanicalObject.inl:2333-2356) and the device sums in double in a fixed tree; the
         bounds are 100x / 10x the oracle's own sensitivity to that summation order (measured: 4e-10 / 2.5e-5)."""
    g = gpu_scene("C1", dtype, method)
    s = oracle_scene("C1", dtype, method)
    node = g["node"]
    for step in range(10):
        _sync_state(g, s)
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step()
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1, (step, it, it_ref)
        assert rel_err(node.get("dx"), s.get("sol")) <= (1e-8 if dtype == np.float64 else 2e-4), step
        assert np.abs(g["mo"].x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= (1e-11 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("opts", [dict(warmStart=1), dict(trapezoidalScheme=1), dict(firstOrder=1), dict(vdamping=0.7), dict(warmStart=1, trapezoidalScheme=1, vdamping=0.3)],
                         ids=["warmStart", "trapezoidal", "firstOrder", "vdamping", "combined"])
def test_solver_options_step_parity_from_same_state(dtype, opts):
    """The Data of EulerImplicitSolver (trapezoidalScheme, firstOrder, vdamping: EulerImplicitSolver.cpp:117-123,169-171,285-300) and of
    CGLinearSolver (warmStart: CGLinearSolver.inl:118-128): f and b bit-identical, iteration count +-1, and the CG solution within
    10x the reference's OWN sensitivity to the summation order of its dot products on that system (a second oracle, same state, dots
    summed in double in reverse order, is the yardstick; the first-order system is the most sensitive one)."""
    g = gpu_scene("C1", dtype, "large")
    s = oracle_scene("C1", dtype, "large")
    s2 = oracle_scene("C1", dtype, "large"); s2.set_dot_double(True, reverse=True)
    node = g["node"]
    node.set_params(**opts)
    o = dict(opts)
    if "trapezoidalScheme" in o:
        o["trapezoidal"] = o.pop("trapezoidalScheme")
    s.set_params(**o); s2.set_params(**o)
    for step in range(6):
        _sync_state(g, s)
        s2.set_x(s.get("x")); s2.set_v(s.get("v"))
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step(); s2.step()
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1, (step, it, it_ref)
        yard = rel_err(s2.get("sol"), s.get("sol"))
        assert rel_err(node.get("dx"), s.get("sol")) <= max(10 * yard, 1e-8 if dtype == np.float64 else 2e-4), (step, yard)
        xyard = np.abs(s2.get("x") - s.get("x")).max()
        assert np.abs(g["mo"].x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= max(10 * xyard, 1e-11 if dtype == np.float64 else 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_free_running_trajectory_stays_within_the_references_own_sensitivity(dtype):
    """Free-running 10 steps.  A 25-iteration truncated CG on a stiff system amplifies last-bit differences of the dot
    products from step to step, in the reference itself: the oracle run with its dots summed in reverse order drifts
    from the oracle by `yard`.  The device trajectory must stay within 20x that yardstick (and below 1e-3 of the
    beam length in any case).  The yardstick is ONE sample of a chaotic amplification (the only other summation order the
    oracle offers), so the factor is an order of magnitude, not a fit: the device's own deviation moves between 3x and 10.2x
    of it when nothing but the assignment of shared nodes to CTAs (i.e. the grouping of the dot-product partials) changes."""
    g = gpu_scene("C1", dtype, "large")
    s = oracle_scene("C1", dtype, "large")
    s_rev = oracle_scene("C1", dtype, "large"); s_rev.set_dot_double(True, reverse=True)
    for step in range(10):
        g["node"].step(); s.step(); s_rev.step()
    yard = np.abs(s.get("x") - s_rev.get("x")).max()
    dev_err = np.abs(g["mo"].x.cpu().numpy().astype(np.float64) - s.get("x")).max()
    assert dev_err <= 20 * yard + 1e-12, (dev_err, yard)
    assert dev_err <= 1e-3 * 40.0


def test_reference_golden_beam_on_gpu():
    """BaseTetrahedronFEMForceField_test.h:379-431: 4x10x4 beam, 100 steps, position[159] (EXPECT_NEAR 1e-4)."""
    for dtype, tol in ((np.float64, 1e-4), (np.float32, 2e-3)):
        g = gpu_scene("GRID_TEST", dtype)
        for _ in range(100):
            g["node"].step()
        x = g["mo"].x.cpu().numpy()
        assert np.abs(x[159] - np.array([9.99985, 45.0487, 30.0011])).max() <= tol, (dtype, x[159])
        if dtype == np.float64:
            rot = g["ff"].get("rotations")[100]
            exp = np.array([[-1, 8.01488e-06, 0.000541687], [-0.000320764, -0.814541, -0.580106], [0.000436576, -0.580106, 0.814541]])
            assert np.abs(rot - exp).max() <= 1e-4


@pytest.mark.parametrize("dtype", DTYPES)
def test_run_to_run_bitwise_reproducible(dtype):
    outs = []
    for _ in range(2):
        g = gpu_scene("C2_SMALL", dtype)
        for _ in range(3):
            g["node"].step()
        outs.append((g["mo"].x.cpu().numpy().tobytes(), g["mo"].v.cpu().numpy().tobytes(), g["node"].last_solve()["iterations"]))
    assert outs[0] == outs[1]


def test_step_host_equals_device_step():
    g1 = gpu_scene("C1", np.float32); g2 = gpu_scene("C1", np.float32)
    x = g1["pos"].astype(np.float32).copy(); v = np.zeros_like(x)
    for _ in range(3):
        g1["node"].step()
        g2["node"].step_host(x, v)
    assert g1["mo"].x.cpu().numpy().tobytes() == x.tobytes() and g1["mo"].v.cpu().numpy().tobytes() == v.tobytes()


def test_empty_and_degenerate_inputs():
    import sofa_b200 as sb
    ctx = sb.Context(0)
    # a node that belongs to no element, and a single element
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [5, 5, 5]], np.float64)
    mo = sb.MechanicalObject(ctx, "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, np.array([[2, 3, 1, 0]], np.uint32), 1000.0, 0.3, "large")
    s = O.OracleScene(np.float64, pos); s.set_tets(np.array([[2, 3, 1, 0]], np.uint32), "large", 1000.0, 0.3)
    x = pos + 0.1 * np.random.default_rng(8).standard_normal(pos.shape)
    f0 = np.ones_like(x)
    f_d = dev(mo, f0); ff.addForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes()
    # golden single-tetra values of the reference test (BaseTetrahedronFEMForceField_test.h:290-315)
    assert np.abs(ff.get("materialsStiffnesses")[0] - [224.359, 96.1538, 64.1026]).max() < 1e-3
    # out-of-range index is rejected, not silently computed
    with pytest.raises(sb.Sofab200Error):
        sb.TetrahedronFEMForceField(mo, np.array([[0, 1, 2, 7]], np.uint32), 1000.0, 0.3, "large")
    # no elements at all: addForce leaves f untouched
    ff0 = sb.TetrahedronFEMForceField(mo, np.zeros((0, 4), np.uint32), 1000.0, 0.3, "large")
    f_d = dev(mo, f0); ff0.addForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == f0.tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_per_node_outputs_with_isolated_nodes_and_no_elements(dtype):
    """getRotations / computeVonMisesStress / MeshMatrixMass on degenerate topologies: nodes that belong to no element (the reference
    gives them element _rotationIdx[node] = 0, a zero nodal stress and a zero mass row), one single element, and no element at all."""
    import torch
    import sofa_b200 as sb
    ctx = sb.Context(0)
    template = "B200Vec3f" if dtype == np.float32 else "B200Vec3d"
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [5, 5, 5], [1, 1, 1], [-3, 0, 2]], np.float64)
    tets = np.array([[2, 3, 1, 0], [1, 2, 3, 5]], np.uint32)                      # nodes 4 and 6 are isolated
    mo = sb.MechanicalObject(ctx, template, position=pos)
    x = (pos + 0.1 * np.random.default_rng(8).standard_normal(pos.shape)).astype(dtype)
    for method in ("large", "polar"):
        ff = sb.TetrahedronFEMForceField(mo, tets, 1000.0, 0.3, method, computeVonMisesStress=1)
        s = O.OracleScene(dtype, pos); s.set_tets(tets, method, 1000.0, 0.3)
        f0 = np.zeros_like(x)
        ff.addForce(dev(mo, f0), dev(mo, x)); s.fem_add_force(f0, x)
        assert ff.getRotations().cpu().numpy().tobytes() == s.tet_get_rotations().tobytes()
        pe, pn = ff.computeVonMisesStress(dev(mo, x))
        pe_ref, pn_ref = s.tet_von_mises(x, 1)
        assert pe.cpu().numpy().tobytes() == pe_ref.tobytes() and pn.cpu().numpy().tobytes() == pn_ref.tobytes()
        assert pn_ref[4] == 0 and pn_ref[6] == 0
    mm = sb.MeshMatrixMass(mo, tets, massDensity=2.0)
    ref = O.OracleMeshMatrixMass(dtype, pos, tets, 2.0)
    assert mm.vertexMass_host[4] == 0 and mm.edges.shape[0] == 9
    r0 = np.ones_like(x); r = dev(mo, r0)
    mm.addMDx(r, dev(mo, x), 0.7)
    assert r.cpu().numpy().tobytes() == ref.addMDx(r0, x, 0.7).tobytes()
    # no element at all
    none = np.zeros((0, 4), np.uint32)
    ff0 = sb.TetrahedronFEMForceField(mo, none, 1000.0, 0.3, "large", computeVonMisesStress=2)
    R0 = ff0.getRotations().cpu().numpy()
    assert np.array_equal(R0, np.broadcast_to(np.eye(3, dtype=dtype), R0.shape))
    pe, pn = ff0.computeVonMisesStress(dev(mo, x))
    assert pe.numel() == 0 and float(pn.abs().max()) == 0
    mm0 = sb.MeshMatrixMass(mo, none, massDensity=2.0)
    r = dev(mo, r0); mm0.addMDx(r, dev(mo, x), 0.7)
    assert r.cpu().numpy().tobytes() == r0.astype(dtype).tobytes()
    mo0 = sb.MechanicalObject(ctx, template, position=np.zeros((0, 3)))
    mme = sb.MeshMatrixMass(mo0, none)
    e = torch.zeros((0, 3), dtype=mo0.tdtype, device=ctx.device)
    mme.addMDx(e, e.clone(), 1.0); mme.addForce(e, (0, -9, 0))


@pytest.mark.parametrize("dtype", DTYPES)
def test_liver_scene_as_written(dtype):
    """examples/Demos/liver.scn as it is written (collision and visual sub-nodes left out): MeshGmshLoader mesh/liver.msh (181 nodes,
    596 tetrahedra; fixture tests/golden/liver_mesh.npz), DiagonalMass massDensity=1, TetrahedralCorotationalFEMForceField method=large
    E=3000 nu=0.3, FixedProjectiveConstraint 3 39 64, EulerImplicit rayleigh 0.1 / 0.1, CG 25 / 1e-9 / 1e-9, dt 0.02, gravity -9.81.
    Ten steps from the oracle's own state: f and b bit-identical, CG iteration counts equal, dx within the vDot bound (oracle dots in double)."""
    import os
    import torch
    import sofa_b200 as sb
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "liver_mesh.npz"))
    pos, tets = z["positions"], z["tetrahedra"]
    fixed = np.array([3, 39, 64], np.uint32)
    prm = dict(dt=0.02, gravity=(0.0, -9.81, 0.0), rayleighStiffness=0.1, rayleighMass=0.1, iterations=25, tolerance=1e-9, threshold=1e-9)
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedralCorotationalFEMForceField(mo, tets, youngModulus=3000.0, poissonRatio=0.3, method="large")
    node = sb.SolverNode(mo, ff, sb.DiagonalMass(mo, tets, massDensity=1.0), sb.FixedProjectiveConstraint(mo, fixed), **prm)
    s = O.OracleScene(dtype, pos)
    s.set_params(**prm)
    s.set_mass_density(1.0, tets); s.set_tets(tets, "large", 3000.0, 0.3); s.set_tetrahedral_corotational(True); s.set_fixed(fixed)
    s.set_dot_double(True)   # vDot summed in double like the device: on this irregular mesh 25 CG iterations amplify the Real-vs-double
                             # summation order of the reference's serial vDot to 2e-3 in Vec3f, which is not what this test is about
    # ... and the yardstick for it: the reference semantics proper (serial vDot in Real), stepped from the same states.  The device's CG (one
    # reduction per iteration, cg_fused.cuh) must stay at least as close to the double-dot oracle as the reference's own CG does.
    s2 = O.OracleScene(dtype, pos)
    s2.set_params(**prm)
    s2.set_mass_density(1.0, tets); s2.set_tets(tets, "large", 3000.0, 0.3); s2.set_tetrahedral_corotational(True); s2.set_fixed(fixed)
    for step in range(10):
        mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))
        s2.set_x(s.get("x")); s2.set_v(s.get("v"))
        node.step()
        it, it_ref = node.last_solve()["iterations"], s.step()
        s2.step()
        yard = rel_err(s2.get("sol"), s.get("sol"))
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1, (step, it, it_ref)
        # (the yardstick is ONE sample of a chaotic amplification over 25 truncated CG iterations: a factor, not a fit)
        assert rel_err(node.get("dx"), s.get("sol")) <= (1e-8 if dtype == np.float64 else max(2e-4, 3 * yard)), (step, yard)
    assert np.abs(s.get("x") - pos).max() > 1e-3      # the organ did move
