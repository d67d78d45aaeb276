"""Host coupling of the device-resident solver node: MechanicalObject's externalForce (accumulateForce, MechanicalObject.inl:1356-1375) against the
oracle, and the pipelined step (external forces up, positions down on a copy stream under the next step) against the synchronous one."""
import numpy as np
import pytest
import torch

from gpu_common import gpu_scene, oracle_scene

pytestmark = pytest.mark.gpu
DTYPES = [np.float64, np.float32]


@pytest.mark.parametrize("dtype", DTYPES)
def test_external_force_matches_oracle(dtype):
    g = gpu_scene("C1", dtype)
    s = oracle_scene("C1", dtype)
    s.set_dot_double(True)
    rng = np.random.default_rng(17)
    ext = (0.3 * rng.standard_normal(g["pos"].shape)).astype(dtype)
    ext[::3] = 0                      # rows equal to Deriv() are skipped by the reference
    ext[1, 0] = -0.0                  # a negative zero does not survive `f += ext` on a reset f
    g["node"].set_external_force(ext); s.set_external_force(ext)
    for it in range(3):
        g["node"].step(); s_it = s.step()
        assert g["node"].get("f").tobytes() == s.get("f").tobytes() or it > 0       # (first step: same state on both sides)
        assert g["node"].get("b").tobytes() == s.get("b").tobytes() or it > 0
        assert abs(g["node"].last_solve()["iterations"] - s_it) <= 1
        assert np.abs(g["mo"].x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= (1e-11 if dtype == np.float64 else 1e-5)
    # removing it again restores the plain path
    g["node"].set_external_force(None); s.set_external_force(None)
    g["node"].step(); s.step()
    assert np.abs(g["mo"].x.cpu().numpy().astype(np.float64) - s.get("x")).max() <= (1e-10 if dtype == np.float64 else 2e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_pipelined_step_equals_synchronous_step(dtype):
    """The same external forces step by step: the positions the pipelined path delivers for step k (one call later, alternating buffers) are
    bit-identical to those of the synchronous path."""
    a = gpu_scene("C2_SMALL", dtype)
    b = gpu_scene("C2_SMALL", dtype)
    rng = np.random.default_rng(19)
    K = 8
    exts = [(0.05 * rng.standard_normal(a["pos"].shape)).astype(dtype) for _ in range(K)]
    ref = []
    for k in range(K):
        a["node"].set_external_force(exts[k]); a["node"].step()
        ref.append(a["mo"].x.cpu().numpy().copy())
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    out = [torch.zeros(a["pos"].shape, dtype=tdt).pin_memory() for _ in range(2)]
    ext_pinned = torch.zeros(a["pos"].shape, dtype=tdt).pin_memory()
    got = []
    for k in range(K):
        ext_pinned.copy_(torch.from_numpy(exts[k]))
        b["node"].step_pipelined(ext_pinned, out[k & 1])
        if k > 0:
            got.append(out[(k - 1) & 1].numpy().copy())       # complete on return of the next call
    b["node"].flush()
    got.append(out[(K - 1) & 1].numpy().copy())
    for k in range(K):
        assert got[k].tobytes() == ref[k].tobytes(), k


@pytest.mark.parametrize("dtype", DTYPES)
def test_mo_accumulate_force_per_operation(dtype):
    """sofab200_mo_accumulate_force == MechanicalObject::accumulateForce: f[i] += ext[i] only for rows that differ from Deriv() (a row of zeros,
    signed or not, leaves f untouched -- so a -0 in f survives there, and `-0 + ext` rounds as the reference's `+=` does elsewhere)."""
    import sofa_b200 as sb
    from gpu_common import dev, mesh
    c, pos, hexas, tets, fixed = mesh("C1")
    mo = sb.MechanicalObject(sb.Context(0), "B200Vec3f" if dtype == np.float32 else "B200Vec3d", position=pos)
    rng = np.random.default_rng(23)
    f = rng.standard_normal(pos.shape).astype(dtype); f[5] = -0.0
    ext = rng.standard_normal(pos.shape).astype(dtype); ext[::2] = 0; ext[4, 1] = -0.0; ext[5] = -0.0
    ref = f.copy()
    for i in range(pos.shape[0]):
        if not (ext[i] == 0).all():
            ref[i] += ext[i]
    f_d = dev(mo, f)
    mo.accumulateForce(f_d, dev(mo, ext))
    assert f_d.cpu().numpy().tobytes() == ref.tobytes()
