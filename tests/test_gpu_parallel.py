"""Multi-GPU path on real devices (needs >= 2 GPUs; `gpurun --gpus 2` / `--gpus 8`), against the single-domain oracle, in its forms:
"peer"  one persistent CG kernel per GPU (cg_fused.cuh) exchanging over NVLink peer memory (sofab200_node_set_peer); "peer_v1": the first-generation kernel,
"nccl"  the library's multi-kernel loop with NCCL send/recv + allreduce (sofab200_node_set_distributed only),
"torch" the host-driven loop of sofa_b200/parallel.py over torch.distributed (what tests/test_parallel_cpu.py runs on gloo)."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle_lib as O
from gpu_common import mesh

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, cfg_name, dtype_name, mode, q_out):
    import torch.distributed as dist
    import sofa_b200 as sb
    import sofa_b200.parallel as PAR
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["SOFAB200_PEER"] = "1" if mode.startswith("peer") else "0"
    os.environ["SOFAB200_CG_FUSED"] = "0" if mode == "peer_v1" else "1"      # peer_v1: the first-generation persistent kernel (two cross-GPU syncs per iteration)
    native = mode != "torch"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        c, pos, hexas, tets, fixed = mesh(cfg_name)
        ctx = sb.Context(rank)
        node = PAR.DistributedSolverNode(pos, tets, fixed, c["density"], c["young"], c["poisson"], "large", ctx=ctx,
                                         template="B200Vec3f" if dtype_name == "f32" else "B200Vec3d", native=native, dt=c["dt"], gravity=c["gravity"],
                                         rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
        rm = node.rm
        rng = np.random.default_rng(0)
        p_glob = rng.standard_normal(pos.shape)
        p_loc = torch.from_numpy(p_glob[rm.global_ids].astype(np.float32 if dtype_name == "f32" else np.float64)).to(ctx.device)
        node.be.compute_force(node.f, node.be.x)      # rotations at rest
        q_loc = node.apply(node.be.new_vector(), p_loc, 1.001, -0.01, -0.0011)
        q_glob = node.gather_global(q_loc, pos.shape[0])
        if native:      # MechanicalObject's externalForce on a partitioned node: every rank passes the rows of its nodes, the owner's copy of an interface node keeps them
            ext_glob = (0.05 * np.random.default_rng(1).standard_normal(pos.shape)).astype(np.float32 if dtype_name == "f32" else np.float64)
            node.be.node.set_external_force(ext_glob[rm.global_ids])
        its = [node.step() for _ in range(3)]
        if native:
            its = [node.be.node.last_solve()["iterations"]] * 3
        x_glob = node.gather_global(node.be.x, pos.shape[0])
        if rank == 0:
            q_out.put(dict(q=q_glob, x=x_glob, its=its, peer=bool(getattr(node.be, "peer", False))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["peer", "peer_v1", "nccl", "torch"])
@pytest.mark.parametrize("dtype_name", ["f64", "f32"])
def test_n_gpus_match_single_domain_oracle(dtype_name, mode, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if world > 2 and mode not in ("peer", "nccl"):
        pytest.skip("the larger worlds exercise the two library loops")
    import torch.multiprocessing as mp
    cfg = "C2_SMALL"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cfg, dtype_name, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res["peer"] == mode.startswith("peer")     # the mode under test is the one that ran
    dtype = np.float32 if dtype_name == "f32" else np.float64
    c, pos, hexas, tets, fixed = mesh(cfg)
    s = O.OracleScene(dtype, pos)
    s.set_params(gravity=c["gravity"], dt=c["dt"], rayleighStiffness=c["rK"], rayleighMass=c["rM"], iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    s.set_mass_density(c["density"], tets); s.set_tets(tets, "large", c["young"], c["poisson"]); s.set_fixed(fixed)
    s.set_dot_double(True)
    rng = np.random.default_rng(0)
    p_glob = rng.standard_normal(pos.shape).astype(dtype)
    q_ref = s.apply(p_glob, 1.001, -0.01, -0.0011)
    err = np.linalg.norm(res["q"] - q_ref) / np.linalg.norm(q_ref)
    assert err <= (1e-14 if dtype == np.float64 else 1e-6), err      # only the interface sums are associated differently
    if mode != "torch":
        s.set_external_force((0.05 * np.random.default_rng(1).standard_normal(pos.shape)).astype(dtype))
    its_ref = [s.step() for _ in range(3)]
    assert all(abs(a - b) <= 1 for a, b in zip(res["its"], its_ref))
    assert np.abs(res["x"] - s.get("x")).max() <= (1e-7 if dtype == np.float64 else 1e-4)
