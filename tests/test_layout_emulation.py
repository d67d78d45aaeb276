"""CPU emulation of the product's tile / gather layout and of its host+device element functions,
checked BIT-FOR-BIT against the oracle (no GPU needed).  The CUDA path itself is checked in test_gpu_*.py."""
import numpy as np
import pytest

import oracle_lib as O
from emu_lib import EmuTet


def _beam(n=(4, 4, 9), mode=1, seed=0, amp=0.03):
    pos, _ = O.regular_grid(n, (0, 0, 0), (1.0, 1.0, 2.5))
    tets = O.hexas_to_tetras(n, mode)
    rng = np.random.default_rng(seed)
    x = pos + amp * rng.standard_normal(pos.shape)
    return pos, tets, x, rng


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", ["small", "large", "polar", "svd"])
@pytest.mark.parametrize("tile_e", [256, 1024])
def test_add_force_and_dforce_bit_exact(dtype, method, tile_e):
    pos, tets, x, rng = _beam()
    s = O.OracleScene(dtype, pos); s.set_tets(tets, method, 1000.0, 0.3)
    e = EmuTet(dtype, pos, tets, method, 1000.0, 0.3, tile_e)
    f0 = rng.standard_normal(pos.shape).astype(dtype)
    xr = x.astype(dtype)
    f_ref = s.fem_add_force(f0, xr)
    f_emu = e.run(False, xr, init=f0)
    assert f_emu.tobytes() == f_ref.tobytes()
    if method != "small":
        assert e.rotations().tobytes() == s.get("tet.rotations").tobytes()
    dx = rng.standard_normal(pos.shape).astype(dtype)
    for kf in (1.0, -0.0011, 0.37):
        df_ref = s.fem_add_dforce(f0, dx, kf)
        df_emu = e.run(True, dx, kf=kf, init=f0)
        assert df_emu.tobytes() == df_ref.tobytes()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_svd_inverted_elements_bit_exact(dtype):
    pos, tets, x, rng = _beam(seed=3)
    x[::7] += 0.4 * rng.standard_normal(x[::7].shape)  # crush / invert a good share of the elements
    s = O.OracleScene(dtype, pos); s.set_tets(tets, "svd", 1000.0, 0.3)
    e = EmuTet(dtype, pos, tets, "svd", 1000.0, 0.3, 256)
    z = np.zeros(pos.shape, dtype)
    xr = x.astype(dtype)
    assert e.run(False, xr, init=z).tobytes() == s.fem_add_force(z, xr).tobytes()
    assert e.rotations().tobytes() == s.get("tet.rotations").tobytes()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mass_first", [1, 0])
def test_fused_apply_matches_graph_scattered_apply(dtype, mass_first):
    pos, tets, x, rng = _beam(n=(5, 4, 8))
    s = O.OracleScene(dtype, pos)
    s.set_params(massFirst=mass_first)
    s.set_mass_density(2.0, tets); s.set_tets(tets, "large", 800.0, 0.35)
    fixed_idx = O.box_roi(pos, (-1, -1, -1, 2, 2, 0.01)); s.set_fixed(fixed_idx)
    xr = x.astype(dtype)
    s.fem_add_force(np.zeros_like(xr), xr)
    e = EmuTet(dtype, pos, tets, "large", 800.0, 0.35, 256)
    e.run(False, xr, init=np.zeros_like(xr))
    mass = s.get("vertexMass")
    mask = np.zeros(pos.shape[0], np.uint8); mask[fixed_idx] = 1
    p = rng.standard_normal(pos.shape).astype(dtype)
    m, b, k = 1.001, -0.01, -0.0011
    q_ref = s.apply(p, m, b, k)
    kind = dict(pre_kind=2) if mass_first else dict(post_kind=2)
    q_emu, dot = e.run(True, p, kf=k, mass=mass, mass_factor=m, fixed=mask, dot=True, **kind)
    assert q_emu.tobytes() == q_ref.tobytes()
    assert abs(dot - float(np.sum(q_ref.astype(np.float64) * p))) <= 1e-12 * abs(dot) + 1e-30
    # computeForce: gravity + elastic force in one pass
    s.set_params(gravity=(0.0, -9.0, 0.5)); s.set_x(xr)
    f_ref = s.compute_force()
    kind = dict(pre_kind=1) if mass_first else dict(post_kind=1)
    f_emu = e.run(False, xr, mass=mass, gravity=(0.0, -9.0, 0.5), **kind)
    assert f_emu.tobytes() == f_ref.tobytes()


def test_layout_invariants_and_isolated_nodes():
    pos, tets, x, rng = _beam(n=(6, 6, 11))
    pos = np.vstack([pos, [[9, 9, 9], [8, 8, 8]]])  # two nodes that belong to no element
    e = EmuTet(np.float32, pos, tets, "large", 1000.0, 0.3, 256)
    st = e.stats()
    assert st["interior"] + st["shared"] == pos.shape[0]
    assert st["tiles"] == (tets.shape[0] + 255) // 256
    s = O.OracleScene(np.float32, pos); s.set_tets(tets, "large", 1000.0, 0.3)
    xr = np.vstack([x, pos[-2:]]).astype(np.float32)
    f0 = rng.standard_normal(pos.shape).astype(np.float32)
    assert e.run(False, xr, init=f0).tobytes() == s.fem_add_force(f0, xr).tobytes()


def test_slot_budget_demotes_interior_nodes_bit_exact(monkeypatch):
    """A tile whose interior corners exceed the shared-memory budget keeps only what fits; the rest goes through the
    staging path.  Same bits either way."""
    pos, tets, x, rng = _beam(n=(6, 6, 11))
    monkeypatch.setenv("SOFAB200_SMEM_KB", "16")
    e = EmuTet(np.float32, pos, tets, "large", 1000.0, 0.3, 1024)
    st = e.stats()
    assert st["smem"] <= 16 * 1024
    monkeypatch.delenv("SOFAB200_SMEM_KB")
    e_free = EmuTet(np.float32, pos, tets, "large", 1000.0, 0.3, 1024)
    assert e_free.stats()["staged"] < st["staged"]          # the budget pushed corners to the staging path
    s = O.OracleScene(np.float32, pos); s.set_tets(tets, "large", 1000.0, 0.3)
    f0 = rng.standard_normal(pos.shape).astype(np.float32)
    f_ref = s.fem_add_force(f0, x.astype(np.float32))
    assert e.run(False, x.astype(np.float32), init=f0).tobytes() == f_ref.tobytes()
    dx = (1e-3 * rng.standard_normal(pos.shape)).astype(np.float32)
    df0 = rng.standard_normal(pos.shape).astype(np.float32)
    assert e.run(True, dx, kf=-0.37, init=df0).tobytes() == s.fem_add_dforce(df0, dx, -0.37).tobytes()


def test_forced_shared_nodes_bit_exact():
    """Nodes flagged in sofab200_tetfem_desc::shared_nodes (the partition interface of a multi-GPU run) take the staging path even
    when all their elements sit in one tile; the sums stay bit-identical."""
    pos, tets, x, rng = _beam(n=(6, 6, 11))
    forced = np.arange(0, pos.shape[0], 3)
    e = EmuTet(np.float32, pos, tets, "large", 1000.0, 0.3, 512, shared_nodes=forced)
    e0 = EmuTet(np.float32, pos, tets, "large", 1000.0, 0.3, 512)
    assert e.stats()["shared"] > e0.stats()["shared"] and e.stats()["shared"] >= len(forced)
    s = O.OracleScene(np.float32, pos); s.set_tets(tets, "large", 1000.0, 0.3)
    f0 = rng.standard_normal(pos.shape).astype(np.float32)
    assert e.run(False, x.astype(np.float32), init=f0).tobytes() == s.fem_add_force(f0, x.astype(np.float32)).tobytes()
    dx = (1e-3 * rng.standard_normal(pos.shape)).astype(np.float32)
    assert e.run(True, dx, kf=-0.37, init=f0).tobytes() == s.fem_add_dforce(f0, dx, -0.37).tobytes()
