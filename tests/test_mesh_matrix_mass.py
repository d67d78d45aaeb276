"""MeshMatrixMass on tetrahedra: the oracle restatement against the reference's own KAT
(Sofa/Component/Mass/tests/MeshMatrixMass_test.cpp:762-797: one 2x2x2 cube cut into tetrahedra, massDensity 1 -> totalMass 8,
vertexMass[0] = 2/3), the host-side mirror against the oracle bit for bit, and the device kernels against the oracle bit for bit."""
import numpy as np
import pytest

import oracle_lib as O

DTYPES = [np.float32, np.float64]


def _cube():
    pos, hexas = O.regular_grid((2, 2, 2), (0, 0, 0), (2, 2, 2))
    return pos, O.hexas_to_tetras((2, 2, 2), 0)


def _beam():
    pos, hexas = O.regular_grid((4, 5, 7), (0, 0, 0), (1.1, 1.3, 2.9))
    return pos, O.hexas_to_tetras((4, 5, 7), 1)


@pytest.mark.parametrize("dtype", DTYPES)
def test_reference_kat_mass_density_tetra(dtype):
    pos, tets = _cube()
    m = O.OracleMeshMatrixMass(dtype, pos, tets, 1.0)
    tol = 4 * np.finfo(dtype).eps  # EXPECT_FLOATINGPOINT_EQ = 4 ulp
    assert abs(m.totalMass - 8.0) <= 8 * tol
    assert abs(m.vertexMass[0] - 2.0 / 3.0) <= tol
    # the lumped matrix carries the same total: sum(vertexMass) * 2.5
    assert abs(float(m.vertexMass.astype(np.float64).sum()) * 2.5 - 8.0) <= 8 * tol
    # and the sparse one: sum(vertexMass) + 2 sum(edgeMass)
    assert abs(float(m.vertexMass.astype(np.float64).sum()) + 2 * float(m.edgeMass.astype(np.float64).sum()) - 8.0) <= 8 * tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_host_mirror_equals_oracle_bitwise(dtype):
    from sofa_b200 import topology as T
    for pos, tets in (_cube(), _beam()):
        m = O.OracleMeshMatrixMass(dtype, pos, tets, 1.7)
        vm, edges, em, coeff = T.mesh_matrix_mass(pos, tets, dtype, 1.7)
        assert coeff == 2.5
        assert edges.tobytes() == m.edges.tobytes()
        assert vm.tobytes() == m.vertexMass.tobytes() and em.tobytes() == m.edgeMass.tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_add_mdx_is_the_matrix_product(dtype):
    pos, tets = _beam()
    m = O.OracleMeshMatrixMass(dtype, pos, tets, 1.3)
    n = pos.shape[0]
    M = np.diag(m.vertexMass.astype(np.float64))
    for (a, b), w in zip(m.edges, m.edgeMass.astype(np.float64)):
        M[a, b] += w; M[b, a] += w
    rng = np.random.default_rng(0)
    dx = rng.standard_normal((n, 3)); r0 = rng.standard_normal((n, 3))
    got = m.addMDx(r0, dx, -0.37).astype(np.float64)
    want = r0.astype(dtype).astype(np.float64) + (-0.37) * (M @ dx.astype(dtype).astype(np.float64))
    assert np.abs(got - want).max() <= (1e-5 if dtype == np.float32 else 1e-13)
    a, ok = m.accFromF(r0)
    assert ok == 0  # the reference refuses accFromF on the sparse matrix
    lumped = O.OracleMeshMatrixMass(dtype, pos, tets, 1.3, lumping=True)
    a, ok = lumped.accFromF(r0)
    assert ok == 1 and np.allclose(a, r0.astype(dtype) / (lumped.vertexMass * dtype(2.5))[:, None])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("lumping", [False, True])
@pytest.mark.parametrize("meshname", ["beam", "liver"])
def test_device_mesh_matrix_mass_bit_exact(dtype, lumping, meshname):
    import os
    import torch
    import sofa_b200 as sb
    if meshname == "beam":
        pos, tets = _beam()
    else:
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "liver_mesh.npz"))
        pos, tets = z["positions"], z["tetrahedra"]
    ref = O.OracleMeshMatrixMass(dtype, pos, tets, 1.3, lumping=lumping)
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    mm = sb.MeshMatrixMass(mo, tets, massDensity=1.3, lumping=lumping)
    assert mm.vertexMass_host.tobytes() == ref.vertexMass.tobytes()
    rng = np.random.default_rng(3)
    n = pos.shape[0]
    dx = rng.standard_normal((n, 3)).astype(dtype); r0 = rng.standard_normal((n, 3)).astype(dtype)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    for factor in (1.0, -0.37, 1.001):
        r = dev(r0)
        mm.addMDx(r, dev(dx), factor)
        assert r.cpu().numpy().tobytes() == ref.addMDx(r0, dx, factor).tobytes(), factor
    f = dev(r0)
    mm.addForce(f, (0.0, -9.81, 0.3))
    assert f.cpu().numpy().tobytes() == ref.addForce(r0, (0.0, -9.81, 0.3)).tobytes()
    a = dev(np.zeros_like(r0))
    if lumping:
        mm.accFromF(a, dev(r0))
        assert a.cpu().numpy().tobytes() == ref.accFromF(r0)[0].tobytes()
    else:
        with pytest.raises(sb.Sofab200Error):
            mm.accFromF(a, dev(r0))
    # run-to-run reproducible
    r1 = dev(r0); mm.addMDx(r1, dev(dx), 0.5); r2 = dev(r0); mm.addMDx(r2, dev(dx), 0.5)
    assert torch.equal(r1, r2)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("lumping", [False, True])
@pytest.mark.parametrize("cg_path", ["fused_tail", "multi_kernel"])
def test_solver_node_with_mesh_matrix_mass(dtype, lumping, cg_path, monkeypatch):
    """The device-resident solver node with a MeshMatrixMass as its mass component (sofab200_node_set_mesh_mass): force vector, A*p and
    right-hand side BIT-IDENTICAL to the oracle scene with the same mass, CG iteration counts equal, solution within the vDot bounds."""
    import sofa_b200 as sb
    import gpu_common
    from gpu_common import dev, rel_err
    if cg_path == "multi_kernel":
        monkeypatch.setenv("SOFAB200_FUSED_TAIL", "0")
    c, pos, hexas, tets, fixed = gpu_common.mesh("C1")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.TetrahedronFEMForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method="large")
    mass = sb.MeshMatrixMass(mo, tets, massDensity=c["density"], lumping=lumping)
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                         iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s = O.OracleScene(dtype, pos)
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"],
                 threshold=c["threshold"])
    s.set_tets(tets, "large", c["young"], c["poisson"]); s.set_mesh_mass(tets, c["density"], lumping); s.set_fixed(fixed)
    rng = np.random.default_rng(5)
    x = (pos + 0.2 * rng.standard_normal(pos.shape)).astype(dtype)
    s.set_x(x)
    f_d = mo.new_vector(); node.computeForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.compute_force().tobytes()
    p = rng.standard_normal(x.shape).astype(dtype)
    for (m, b, k) in ((1.001, -0.01, -0.0011), (1.0, 0.0, -0.01), (0.0, 0.0, 0.11)):
        q_d = mo.new_vector(); node.apply(q_d, dev(mo, p), m, b, k)
        assert q_d.cpu().numpy().tobytes() == s.apply(p, m, b, k).tobytes(), (m, b, k)
    # a mass of another size or real type is refused
    other = sb.MechanicalObject(ctx, "B200Vec3d" if dtype == np.float32 else "B200Vec3f", position=pos)
    with pytest.raises(sb.Sofab200Error):
        check_node = sb.SolverNode(mo, ff, sb.MeshMatrixMass(other, tets), None)
    s.set_x(pos.astype(dtype))
    for step in range(4):
        import torch
        mo.x.copy_(torch.from_numpy(s.get("x").astype(dtype))); mo.v.copy_(torch.from_numpy(s.get("v").astype(dtype)))
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step()
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1, (step, it, it_ref)
        assert rel_err(node.get("dx"), s.get("sol")) <= (1e-8 if dtype == np.float64 else 2e-4), step
