"""Multi-GPU host logic on CPU: world_size-2 (and 3) gloo process groups, the oracle standing in for the rank-local
device operators (test infrastructure).  Checks the partition, the halo plan, the ordered halo sum and the distributed
CG / EulerImplicit step against the single-domain oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from gpu_common import CONFIGS, mesh


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class OracleBackend:
    """Rank-local operators computed by the CPU oracle on the rank's sub-mesh (same interface as parallel.DeviceBackend)."""

    def __init__(self, rm, m_pass, fixed_local, youngModulus, poissonRatio, method, params, dtype=np.float64):
        self.s = O.OracleScene(dtype, rm.positions)
        self.s.set_params(gravity=params["gravity"], dt=params["dt"], rayleighStiffness=params["rK"], rayleighMass=params["rM"],
                          iterations=params["iterations"], tolerance=params["tolerance"], threshold=params["threshold"])
        self.s.set_vertex_mass(m_pass)
        self.s.set_tets(rm.elems, method, youngModulus, poissonRatio)
        self.s.set_fixed(fixed_local)
        self.device = torch.device("cpu")
        self.np_dtype = dtype
        self.x = torch.from_numpy(np.ascontiguousarray(rm.positions, dtype))
        self.v = torch.zeros_like(self.x)
        self.owned = torch.from_numpy(rm.owned.astype(np.float64))[:, None]
        self.rm = rm

    def new_vector(self):
        return torch.zeros_like(self.x)

    def vop(self, r, a=None, b=None, k=1.0):
        rn = r.numpy()
        an = None if a is None else (rn if a is r else a.numpy())
        bn = None if b is None else (rn if b is r else b.numpy())
        O.vop(self.np_dtype, rn, an, bn, k)

    def dot_owned(self, a, b):
        return torch.tensor([float((a.double() * b.double() * self.owned).sum())], dtype=torch.float64)

    def compute_force(self, f, x):
        self.s.set_x(x.numpy())
        f.copy_(torch.from_numpy(self.s.compute_force()))

    def add_mbkdx(self, out, d, m, b, k, init=None, scale=None, project=False):
        # df = init + (m M + b B + k K) d through the oracle's GraphScattered pieces
        q = self.s.apply(d.numpy(), m, b, k) if project and init is None and scale is None else None
        if q is None:
            base = np.zeros_like(d.numpy()) if init is None else init.numpy().copy()
            mass = self.s.get("vertexMass")
            if m != 0.0:
                base = base + (d.numpy() * mass[:, None]) * self.np_dtype(m)
            base = self.s.fem_add_dforce(base, d.numpy(), k)
            if scale is not None:
                base = base * self.np_dtype(scale)
            if project:
                base[self.fixed_idx()] = 0
            q = base
        out.copy_(torch.from_numpy(np.ascontiguousarray(q, self.np_dtype)))

    def fixed_idx(self):
        return np.nonzero(np.isin(self.rm.global_ids, self._fixed_global))[0]

    def integrate(self, v, x, a, h):
        v += a
        x += v * self.np_dtype(h)


def _worker(rank, world, port, cfg_name, q_out, partition="slab"):
    import sofa_b200.parallel as PAR
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        c, pos, hexas, tets, fixed = mesh(cfg_name)

        def factory(**kw):
            be = OracleBackend(**kw)
            be._fixed_global = fixed
            return be
        node = PAR.DistributedSolverNode(pos, tets, fixed, c["density"], c["young"], c["poisson"], "large", backend_factory=factory, template="B200Vec3d",
                                         partition=partition, dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"],
                                         tolerance=c["tolerance"], threshold=c["threshold"])
        rm = node.rm
        # (1) partition invariants
        owned_count = torch.tensor([int(rm.owned.sum())]); dist.all_reduce(owned_count)
        assert int(owned_count) == pos.shape[0]
        elems_count = torch.tensor([rm.elems.shape[0]]); dist.all_reduce(elems_count)
        assert int(elems_count) == tets.shape[0]
        # (2) distributed A*p against the single-domain oracle
        rng = np.random.default_rng(0)
        p_glob = rng.standard_normal(pos.shape)
        p_loc = torch.from_numpy(p_glob[rm.global_ids].copy())
        node.be.s.fem_add_force(np.zeros_like(rm.positions), rm.positions)          # rotations at rest on the sub-mesh
        q_loc = node.apply(node.be.new_vector(), p_loc, 1.001, -0.01, -0.0011)
        q_glob = node.gather_global(q_loc, pos.shape[0])
        # (3) a few distributed EulerImplicit steps
        its = [node.step() for _ in range(3)]
        x_glob = node.gather_global(node.be.x, pos.shape[0])
        if rank == 0:
            q_out.put(dict(q=q_glob, x=x_glob, its=its, n_if=len(rm.interface), max_sharers=rm.max_sharers))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,partition", [(2, "slab"), (3, "slab"), (3, "rcb"), (4, "rcb")])
def test_distributed_apply_and_steps_match_single_domain(world, partition):
    cfg = "C1"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, q, partition)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    c, pos, hexas, tets, fixed = mesh(cfg)
    s = O.OracleScene(np.float64, pos)
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_mass_density(c["density"], tets); s.set_tets(tets, "large", c["young"], c["poisson"]); s.set_fixed(fixed)
    rng = np.random.default_rng(0)
    p_glob = rng.standard_normal(pos.shape)
    q_ref = s.apply(p_glob, 1.001, -0.01, -0.0011)
    assert res["n_if"] > 0
    err = np.linalg.norm(res["q"] - q_ref) / np.linalg.norm(q_ref)
    assert err <= 1e-14, err            # identical up to the association of the interface sums
    its_ref = [s.step() for _ in range(3)]
    assert all(abs(a - b) <= 1 for a, b in zip(res["its"], its_ref)), (res["its"], its_ref)
    # interface sums are associated differently than in the sequential loop (1 ulp), amplified by the truncated CG
    assert np.abs(res["x"] - s.get("x")).max() <= 1e-7


def test_rank_mesh_and_halo_plan():
    import sofa_b200.parallel as PAR
    c, pos, hexas, tets, fixed = mesh("C2_SMALL")
    world = 4
    meshes = [PAR.RankMesh(pos, tets, r, world) for r in range(world)]
    assert sum(m.elems.shape[0] for m in meshes) == tets.shape[0]
    assert sum(int(m.owned.sum()) for m in meshes) == pos.shape[0]
    for r, m in enumerate(meshes):
        # z-slabs: at most two neighbours, symmetric interface lists in the same (global id) order
        assert set(m.neighbours) <= {r - 1, r + 1}
        for s, nb in m.neighbours.items():
            other = meshes[s].neighbours[r]
            assert np.array_equal(m.global_ids[nb["local"]], meshes[s].global_ids[other["local"]])
        assert np.array_equal(pos[m.global_ids], m.positions)
        assert np.array_equal(m.global_ids[m.elems.astype(np.int64)], tets[PAR.element_ranges(tets.shape[0], world)[r][0]:PAR.element_ranges(tets.shape[0], world)[r][1]])


def _liver_replicated(k):
    """The liver mesh (tests/golden/liver_mesh.npz) replicated k x k x k times with a random permutation of the element list: a mesh whose
    element numbering says nothing about space (copies are disjoint: the partitioner must keep each copy's elements together to do well)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "liver_mesh.npz"))
    pos0, tets0 = z["positions"].astype(np.float64), z["tetrahedra"].astype(np.int64)
    ext = pos0.max(0) - pos0.min(0)
    pos, tets = [], []
    for i in range(k):
        for j in range(k):
            for l in range(k):
                tets.append(tets0 + len(pos) * pos0.shape[0])
                pos.append(pos0 + np.array([i, j, l]) * ext * 0.97)      # (slightly overlapping boxes: centroids of neighbouring copies interleave)
    pos, tets = np.concatenate(pos), np.concatenate(tets)
    rng = np.random.default_rng(3)
    return pos, tets[rng.permutation(tets.shape[0])]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_rcb_partition_of_a_mesh_without_spatial_numbering(world):
    import sofa_b200.parallel as PAR
    pos, tets = _liver_replicated(4)          # 64 livers: 38 144 tetrahedra, shuffled
    part = PAR.partition_elements(pos, tets, world, "rcb")
    counts = np.bincount(part, minlength=world)
    assert counts.sum() == tets.shape[0] and counts.max() - counts.min() <= world          # balanced
    assert np.array_equal(part, PAR.partition_elements(pos, tets, world, "rcb"))            # deterministic
    n_rcb = PAR.interface_node_count(tets, part, world)
    n_slab = PAR.interface_node_count(tets, PAR.partition_elements(pos, tets, world, "slab"), world)
    # contiguous ranges of a shuffled list put every node on the interface; bisection keeps it to the cut surfaces
    assert n_slab > 0.5 * pos.shape[0]
    assert n_rcb < 0.1 * pos.shape[0], (n_rcb, n_slab, pos.shape[0])
    meshes = [PAR.RankMesh(pos, tets, r, world, "rcb") for r in range(world)]
    assert sum(m.elems.shape[0] for m in meshes) == tets.shape[0]
    assert sum(int(m.owned.sum()) for m in meshes) == np.unique(tets).shape[0]
    for r, m in enumerate(meshes):
        for s_, nb in m.neighbours.items():
            other = meshes[s_].neighbours[r]
            assert np.array_equal(m.global_ids[nb["local"]], meshes[s_].global_ids[other["local"]])     # symmetric interface lists, same order


def test_rcb_on_a_grid_beam_is_as_good_as_slabs():
    import sofa_b200.parallel as PAR
    c, pos, hexas, tets, fixed = mesh("C2_SMALL")
    for world in (2, 4):
        n_rcb = PAR.interface_node_count(tets, PAR.partition_elements(pos, tets, world, "rcb"), world)
        n_slab = PAR.interface_node_count(tets, PAR.partition_elements(pos, tets, world, "slab"), world)
        assert n_rcb <= 1.25 * n_slab, (world, n_rcb, n_slab)
