"""HexahedronFEMForceField on the device against the oracle (BASELINE config C3 family: grid beam of hexahedra, method=polar)."""
import numpy as np
import pytest

import oracle_lib as O
from gpu_common import dev, rel_err

pytestmark = pytest.mark.gpu
DTYPES = [np.float32, np.float64]


def _scene(dtype, method, n=(5, 5, 13), mx=(1.0, 1.0, 3.0), perturb_rest=0.0, seed=0):
    import sofa_b200 as sb
    from sofa_b200 import topology as T
    pos, hexas = T.regular_grid(n, (0, 0, 0), mx)
    rng = np.random.default_rng(seed)
    if perturb_rest:
        pos = pos + perturb_rest * rng.standard_normal(pos.shape)   # non-parallelepiped elements: every K_e differs
    fixed = T.box_roi(pos, (-1, -1, -1, 2, 2, 1e-3 + perturb_rest * 4))
    ctx = sb.Context(0)
    template = "B200Vec3f" if np.dtype(dtype) == np.float32 else "B200Vec3d"
    mo = sb.MechanicalObject(ctx, template, position=pos)
    ff = sb.HexahedronFEMForceField(mo, hexas, youngModulus=1000.0, poissonRatio=0.3, method=method)
    mass = sb.DiagonalMass(mo, hexas, massDensity=1.0)
    node = sb.SolverNode(mo, ff, mass, sb.FixedProjectiveConstraint(mo, fixed), dt=0.01, gravity=(0.0, -9.0, 0.0), rayleighStiffness=0.1,
                         rayleighMass=0.1, iterations=25, tolerance=1e-9, threshold=1e-9)
    s = O.OracleScene(dtype, pos)
    s.set_params(gravity=(0.0, -9.0, 0.0), dt=0.01, rayleighStiffness=0.1, rayleighMass=0.1, iterations=25, tolerance=1e-9, threshold=1e-9)
    s.set_mass_density(1.0, hexas); s.set_hexas(hexas, method, 1000.0, 0.3); s.set_fixed(fixed)
    return dict(mo=mo, ff=ff, node=node, mass=mass, pos=pos, hexas=hexas, fixed=fixed, rng=rng), s


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["large", "polar", "small"])
@pytest.mark.parametrize("perturb", [0.0, 0.01])
def test_hexa_add_force_add_dforce_bit_exact(dtype, method, perturb):
    g, s = _scene(dtype, method, perturb_rest=perturb)
    mo, ff, rng = g["mo"], g["ff"], g["rng"]
    assert ff.get("elementStiffnesses").tobytes() == s.get("hex.Ke").tobytes()
    assert ff.get("rotatedInitialElements").tobytes() == s.get("hex.X0").tobytes()
    st = ff.stats()
    if perturb == 0.0:
        assert st["unique_stiffness_matrices"] == 1     # bit-identical K_e are stored once
    else:
        assert st["unique_stiffness_matrices"] > 1
    x = (g["pos"] + 0.05 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    f0 = rng.standard_normal(x.shape).astype(dtype)
    f_d = dev(mo, f0); ff.addForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(f0, x).tobytes()
    if method != "small":
        assert ff.get("rotations").tobytes() == s.get("hex.rotations").tobytes()
    # getNodeRotation for every node (HexahedronFEMForceField.inl:946-974; the reference starts the mean from the identity)
    assert ff.getRotations().cpu().numpy().tobytes() == s.hex_get_rotations().tobytes()
    dx = rng.standard_normal(x.shape).astype(dtype)
    for kf in (1.0, -0.0011):
        df_d = dev(mo, f0); ff.addDForce(df_d, dev(mo, dx), kf)
        assert df_d.cpu().numpy().tobytes() == s.fem_add_dforce(f0, dx, kf).tobytes(), kf


def test_hexa_small_kat_on_gpu():
    """HexahedronFEMForceField_test.cpp:55-91: unit cube stretched to z=1.1, method small, E=10, nu=0 -> +-0.25 z forces."""
    import sofa_b200 as sb
    x0 = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], np.float64)
    ctx = sb.Context(0)
    mo = sb.MechanicalObject(ctx, "B200Vec3d", position=x0)
    ff = sb.HexahedronFEMForceField(mo, np.arange(8, dtype=np.uint32)[None, :], youngModulus=10.0, poissonRatio=0.0, method="small")
    x = x0.copy(); x[4:, 2] = 1.1
    f = mo.new_vector(); ff.addForce(f, dev(mo, x))
    exp = np.array([[0, 0, 0.25]] * 4 + [[0, 0, -0.25]] * 4)
    assert np.abs(f.cpu().numpy() - exp).max() < 1e-9


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("method", ["polar", "large"])
def test_hexa_step_parity_from_same_state(dtype, method):
    import torch
    g, s = _scene(dtype, method, n=(5, 5, 17), mx=(1.0, 1.0, 4.0))
    node, mo = g["node"], g["mo"]
    for step in range(6):
        mo.x.copy_(torch.from_numpy(s.get("x"))); mo.v.copy_(torch.from_numpy(s.get("v")))
        node.step()
        it = node.last_solve()["iterations"]
        it_ref = s.step()
        assert node.get("f").tobytes() == s.get("f").tobytes(), step
        assert node.get("b").tobytes() == s.get("b").tobytes(), step
        assert abs(it - it_ref) <= 1
        assert rel_err(node.get("dx"), s.get("sol")) <= (1e-8 if dtype == np.float64 else 2e-4), step


def test_hexa_c3_family_operator_properties():
    """A 33x33x61 grid beam of 61 440 hexahedra, method=polar (C3 is the 65x65x121 member of this family)."""
    g, s = _scene(np.float32, "polar", n=(33, 33, 61), mx=(4.0, 4.0, 7.5))
    mo, node, rng = g["mo"], g["node"], g["rng"]
    assert g["ff"].stats()["unique_stiffness_matrices"] == 1   # one unique K_e on an exactly representable grid
    x = (g["pos"] + 0.01 * rng.standard_normal(g["pos"].shape)).astype(np.float32)
    z = np.zeros_like(x)
    f_d = dev(mo, z); g["ff"].addForce(f_d, dev(mo, x))
    assert f_d.cpu().numpy().tobytes() == s.fem_add_force(z, x).tobytes()
    p = dev(mo, rng.standard_normal(x.shape)); q = dev(mo, rng.standard_normal(x.shape))
    p[g["fixed"].astype(np.int64)] = 0; q[g["fixed"].astype(np.int64)] = 0
    Ap, Aq = mo.new_vector(), mo.new_vector()
    node.apply(Ap, p, 1.001, -0.01, -0.0011); node.apply(Aq, q, 1.001, -0.01, -0.0011)
    assert Ap.cpu().numpy().tobytes() == s.apply(p.cpu().numpy(), 1.001, -0.01, -0.0011).tobytes()
    assert abs(mo.vDot(p, Aq) - mo.vDot(q, Ap)) <= 1e-4 * abs(mo.vDot(p, Ap))
    assert not Ap.cpu().numpy()[g["fixed"]].any()
