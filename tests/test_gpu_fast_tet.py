"""FastTetrahedralCorotationalForceField<B200Vec3Types> (sofa_b200/csrc/fast_fem.cu) against the oracle's restatement of the class
(oracle/sofa_oracle.hpp FastTetFEM, pinned on the reference's golden vectors in tests/test_oracle_golden.py): init values, addForce and the
per-edge addDForce bit for bit on a grid beam and on the liver mesh, the edge matrices, a caller-supplied edge list, and EulerImplicit + CG steps."""
import numpy as np
import pytest
import torch

import gpu_common
import oracle_lib as O
from gpu_common import dev

pytestmark = pytest.mark.gpu
DTYPES = [np.float64, np.float32]
METHODS = ["qr", "polar", "polar2", "none"]


def _mesh(name):
    if name == "liver":      # share/mesh/liver.msh as Demos/liver.scn loads it (fixture tests/golden/liver_mesh.npz), E and nu of that scene
        import os
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", "liver_mesh.npz"))
        return dict(young=3000.0, poisson=0.3), z["positions"], z["tetrahedra"], np.array([3, 39, 64])
    c, pos, hexas, tets, fixed = gpu_common.mesh(name)
    return c, pos, tets, fixed


def _pair(name, dtype, method, edges=None, tile=0):
    import sofa_b200 as sb
    c, pos, tets, fixed = _mesh(name)
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    ff = sb.FastTetrahedralCorotationalForceField(mo, tets, youngModulus=c["young"], poissonRatio=c["poisson"], method=method, edges=edges, tileElems=tile)
    s = O.OracleScene(dtype, pos)
    s.set_fast_tets(tets, method, c["young"], c["poisson"], edges=edges)
    return c, pos, tets, fixed, mo, ff, s


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", METHODS)
def test_init_values_bit_exact(dtype, method):
    c, pos, tets, fixed, mo, ff, s = _pair("C1", dtype, method)
    assert ff.get("edges").astype(np.int64).tobytes() == s.get("fast.edges").tobytes()
    for mine, theirs in (("shapeVectors", "fast.shapeVectors"), ("linearDfDx", "fast.linearDfDx"), ("linearDfDxDiag", "fast.linearDfDxDiag"),
                         ("restRotations", "fast.restRotations"), ("restEdgeVectors", "fast.restEdgeVectors"), ("edgeOrientations", "fast.edgeOrientations")):
        assert ff.get(mine).tobytes() == s.get(theirs).tobytes(), mine


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("name", ["C1", "liver"])
def test_add_force_and_add_dforce_bit_exact(dtype, method, name):
    """addForce ([FTC].inl:296-399), the edge matrices (:414-450) and addDForce over the edges (:455-466), three rounds with new positions
    (the matrices are re-assembled after every addForce, and re-used by the second addDForce of a round)."""
    c, pos, tets, fixed, mo, ff, s = _pair(name, dtype, method)
    rng = np.random.default_rng(41)
    scale = 0.2 if name == "C1" else 0.05
    f0 = rng.standard_normal(pos.shape).astype(dtype)
    for it in range(3):
        x = (pos + scale * rng.standard_normal(pos.shape)).astype(dtype)
        f_d = dev(mo, f0)
        ff.addForce(f_d, dev(mo, x))
        assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes(), it
        assert ff.get("rotations").tobytes() == s.get("fast.rotations").tobytes()
        for k in (0.11, -1.0):
            dx = rng.standard_normal(pos.shape).astype(dtype)
            df_d = dev(mo, f0)
            ff.addDForce(df_d, dev(mo, dx), k)
            assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, k).tobytes(), (it, k)
        assert ff.get("edgeInfo").tobytes() == s.get("fast.edgeInfo").tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
def test_small_tiles_and_given_edge_list(dtype):
    """Tiling must not change a bit (tileElems=64: many tiles, most nodes on the staging path); an edge list supplied by the topology
    (here: the container's edges reversed, every edge flipped) changes the numbering and the orientations, and still matches the oracle."""
    c, pos, tets, fixed, mo, ff, s = _pair("C1", dtype, "qr", tile=64)
    edges = ff.get("edges")[::-1, ::-1].copy()
    c, pos, tets, fixed, mo2, ff2, s2 = _pair("C1", dtype, "qr", edges=edges)
    assert (ff2.get("edgeOrientations") != np.asarray(ff.get("edgeOrientations"))).any()
    rng = np.random.default_rng(43)
    x = (pos + 0.2 * rng.standard_normal(pos.shape)).astype(dtype)
    dx = rng.standard_normal(pos.shape).astype(dtype)
    z = np.zeros_like(x)
    for m, f, o in ((mo, ff, s), (mo2, ff2, s2)):
        f_d = dev(m, z); f.addForce(f_d, dev(m, x))
        assert f_d.cpu().numpy().tobytes() == o.fem_add_force(z, x).tobytes()
        df_d = dev(m, z); f.addDForce(df_d, dev(m, dx), 1.0)
        assert df_d.cpu().numpy().tobytes() == o.fem_add_dforce(z, dx, 1.0).tobytes()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["qr", "polar"])
@pytest.mark.parametrize("path", ["fused", "multi_kernel"])
def test_euler_implicit_cg_steps(dtype, method, path, monkeypatch):
    """The class inside EulerImplicitSolver + CGLinearSolver + DiagonalMass + FixedProjectiveConstraint, with the CG loop in the fused persistent
    kernel (A*p over the edges, EdgePass) and as separate kernels: right-hand side of the first step bit for bit, iteration counts within one,
    positions within the Vec3f / Vec3d bands of the other force fields' step tests."""
    import sofa_b200 as sb
    monkeypatch.setenv("SOFAB200_FAST_FUSED", "1" if path == "fused" else "0")
    c, pos, tets, fixed, mo, ff, s = _pair("C1", dtype, method)
    mass = sb.DiagonalMass(mo, tets, massDensity=c["density"])
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                         iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_mass_density(c["density"], tets); s.set_fixed(fixed); s.set_dot_double(True)
    for it in range(5):
        node.step(); s_it = s.step()
        assert abs(node.last_solve()["iterations"] - s_it) <= 1
        if it == 0:
            assert node.get("b").tobytes() == s.get("b").tobytes()
        # (the two sides sum the CG dot products in different orders; measured 4e-9 after 5 steps in Vec3d)
        assert np.abs(mo.x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= (1e-7 if dtype == np.float64 else 2e-4)
    assert np.abs(s.get("x") - pos).max() > 1e-3      # the beam did move
    assert bool(node.fused_info()["fused_enabled"]) == (path == "fused")


def test_refusals():
    import sofa_b200 as sb
    c, pos, tets, fixed = _mesh("C1")
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3f", position=pos)
    with pytest.raises(ValueError):
        sb.FastTetrahedralCorotationalForceField(mo, tets, method="svd")
    ff = sb.FastTetrahedralCorotationalForceField(mo, tets)
    with pytest.raises(sb.Sofab200Error):
        ff.getRotations()
    with pytest.raises(sb.Sofab200Error):
        sb.FastTetrahedralCorotationalForceField(mo, tets, edges=np.array([[0, 1]], np.uint32))     # not every edge of the tetrahedra
