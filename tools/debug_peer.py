#!/usr/bin/env python
"""2-GPU debugging aid: the distributed step with the peer-memory CG kernel against the NCCL loop (same process group)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import sofa_b200 as sb  # noqa: E402
import sofa_b200.parallel as PAR  # noqa: E402
from gpu_common import mesh  # noqa: E402

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2_SMALL"
dtn = sys.argv[2] if len(sys.argv) > 2 else "f32"
c, pos, hexas, tets, fixed = mesh(cfg)
out = {}
for peer in ("0", "1"):
    os.environ["SOFAB200_PEER"] = peer
    ctx = sb.Context(rank)
    node = PAR.DistributedSolverNode(pos, tets, fixed, c["density"], c["young"], c["poisson"], "large", ctx=ctx, template="B200Vec3f" if dtn == "f32" else "B200Vec3d",
                                     native=True, dt=c["dt"], gravity=c["gravity"], rayleighStiffness=c["rK"], rayleighMass=c["rM"],
                                     iterations=c["iterations"], tolerance=c["tolerance"], threshold=c["threshold"])
    res = []
    for s in range(3):
        node.step()
        torch.cuda.synchronize()
        info = node.be.node.last_solve()
        x = node.gather_global(node.be.x, pos.shape[0])
        res.append((info["iterations"], info.get("end_condition"), float(np.abs(x).max())))
        out[(peer, s, "den")] = info["graph_den"]; out[(peer, s, "err")] = info["graph_error"]
        out[(peer, s)] = x
    if rank == 0:
        print("peer", peer, "mode", getattr(node.be, "peer", None), res, flush=True)
if rank == 0:
    for s in range(3):
        print("step", s, "max |x_peer - x_nccl|", float(np.abs(out[("1", s)] - out[("0", s)]).max()), flush=True)
    np.set_printoptions(precision=6, linewidth=200)
    print("den nccl", out[("0", 0, "den")][:6]); print("den peer", out[("1", 0, "den")][:6])
    print("err nccl", out[("0", 0, "err")][:6]); print("err peer", out[("1", 0, "err")][:6])
dist.destroy_process_group()
