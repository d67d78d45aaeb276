"""Multi-GPU host layer: one process per GPU (torch.distributed), the mesh partitioned into owned + ghost nodes, a
per-iteration halo exchange of the nodal A*p partial sums and an allreduce of the CG scalars.

The reference has no distributed mode (no MPI/NCCL anywhere in the tree); this is the multi-GPU design
north_star asks for.  Partition (`partition_elements`): contiguous ranges of the topology's element list, one per rank (on a
RegularGridTopology beam these are z-slabs, so a rank has at most two neighbours), or recursive coordinate bisection of the element
centroids for meshes whose numbering says nothing about space.  A node touched by elements of several ranks is an
interface node; every sharing rank keeps it (ghost copy), the lowest sharing rank owns it.

Per A*p:  each rank runs the fused element pass on its own elements (mass term of an interface node added by its owner
only), then for every interface node the partial sums of all sharing ranks are added IN ASCENDING RANK ORDER on every
sharing rank (same operands, same order => every copy of the node holds the same bits; no atomics, run-to-run
reproducible).  Dot products run over owned nodes and are all-reduced.  Everything else of the CG iteration is local.

The exchange itself (`HaloPlan`, `halo_exchange_sum`) is backend-agnostic torch code: NCCL over NVLink on the GPUs, gloo
on CPU tensors in tests/test_parallel_cpu.py.
"""
import numpy as np
import torch
import torch.distributed as dist


# ---------------------------------------------------------------------------------------------------
# partition (host, numpy)
# ---------------------------------------------------------------------------------------------------
def element_ranges(n_elems, world):
    """Contiguous, balanced element ranges [lo, hi) per rank."""
    base, rem = divmod(n_elems, world)
    lo = [r * base + min(r, rem) for r in range(world)]
    return [(lo[r], lo[r] + base + (1 if r < rem else 0)) for r in range(world)]


def partition_rcb(positions, elems, world):
    """Recursive coordinate bisection of the element CENTROIDS (the partitioner for meshes whose element numbering says
    nothing about space): the current set is cut across its longest extent at the weighted median, the two halves get floor(w/2) and
    ceil(w/2) of the ranks and element counts in the same proportion, recursively.  Deterministic (stable argsort, ties by element index).
    Returns part[e] = rank of element e."""
    elems = np.asarray(elems, np.int64)
    cen = np.asarray(positions, np.float64)[elems].mean(axis=1)
    part = np.zeros(elems.shape[0], np.int32)

    def split(idx, r0, w):
        if w == 1 or len(idx) == 0:
            part[idx] = r0
            return
        c = cen[idx]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0))) if len(idx) else 0
        order = idx[np.argsort(c[:, axis], kind="stable")]
        wl = w // 2
        nl = (len(idx) * wl + w // 2) // w
        split(np.sort(order[:nl]), r0, wl)
        split(np.sort(order[nl:]), r0 + wl, w - wl)

    split(np.arange(elems.shape[0]), 0, world)
    return part


def partition_elements(positions, elems, world, method="slab"):
    """part[e] = rank of element e.  "slab": contiguous ranges of the topology's element list (z-slabs of a RegularGridTopology beam, whose
    hexahedra are numbered z-major: GridTopology.cpp:381-398); "rcb": recursive coordinate bisection, for any mesh."""
    n = np.asarray(elems).shape[0]
    if method in ("slab", "contiguous"):
        part = np.zeros(n, np.int32)
        for r, (a, b) in enumerate(element_ranges(n, world)):
            part[a:b] = r
        return part
    if method == "rcb":
        return partition_rcb(positions, elems, world)
    raise ValueError(f"unknown partition method {method!r}")


def interface_node_count(elems, part, world):
    """Number of nodes touched by elements of more than one part (the quantity a partitioner minimises)."""
    elems = np.asarray(elems, np.int64)
    n_nodes = int(elems.max()) + 1 if elems.size else 0
    cnt = np.zeros(n_nodes, np.int32)
    for r in range(world):
        ids = np.unique(elems[part == r])
        cnt[ids] += 1
    return int((cnt > 1).sum())


class RankMesh:
    """What one rank needs: its elements in local numbering, its nodes (global ids), ownership and the halo plan."""

    def __init__(self, positions, elems, rank, world, partition="slab"):
        elems = np.asarray(elems, np.int64)
        n_nodes = positions.shape[0]
        part = partition_elements(positions, elems, world, partition) if isinstance(partition, str) else np.asarray(partition, np.int32)
        self.rank, self.world, self.part = rank, world, part
        self.ranges = element_ranges(elems.shape[0], world) if isinstance(partition, str) and partition in ("slab", "contiguous") else None
        mine = elems[part == rank]                                       # (original relative order of the rank's elements)
        self.global_ids = np.unique(mine)                               # sorted global node ids of the local nodes
        g2l = np.full(n_nodes, -1, np.int64); g2l[self.global_ids] = np.arange(len(self.global_ids))
        self.elems = g2l[mine].astype(np.uint32)                         # local numbering, original relative order
        self.positions = positions[self.global_ids]
        # which ranks touch each of MY nodes
        touch = np.zeros((len(self.global_ids), world), bool)
        for r in range(world):
            ids = np.unique(elems[part == r])
            l = g2l[ids]; touch[l[l >= 0], r] = True
        self.sharers = touch
        nsh = touch.sum(1)
        first = touch.argmax(1)                                          # lowest sharing rank
        self.owned = (first == rank)
        self.interface = np.nonzero(nsh > 1)[0]                          # local ids, ascending (== ascending global id)
        self.max_sharers = int(nsh.max()) if len(nsh) else 1
        # position of every sharing rank in the node's ascending-rank list
        order = np.cumsum(touch, 1) - 1
        self.my_slot = order[self.interface, rank]
        self.neighbours = {}
        for s in range(world):
            if s == rank:
                continue
            rows = np.nonzero(touch[self.interface, s])[0]               # rows of the interface table shared with s
            if len(rows):
                self.neighbours[s] = dict(rows=rows, local=self.interface[rows], slot=order[self.interface[rows], s])
        self.n_local = len(self.global_ids)


class HaloPlan:
    """Device-side (or CPU-side) index tensors of a RankMesh for halo_exchange_sum."""

    def __init__(self, rm, device):
        self.rank, self.world = rm.rank, rm.world
        self.n_if, self.max_sharers = len(rm.interface), rm.max_sharers
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.int64)).to(device)
        self.if_idx = t(rm.interface)
        self.my_slot = t(rm.my_slot)
        self.nb = [(s, t(v["local"]), t(v["rows"]), t(v["slot"])) for s, v in sorted(rm.neighbours.items())]
        self.device = device


def halo_exchange_sum(q, plan, group=None):
    """Replace the rows of q ([n_local,3]) of the interface nodes by the sum over all sharing ranks of their partial
    values, added in ascending rank order (identical bits on every sharing rank)."""
    if plan.n_if == 0 or not plan.nb:
        return q
    parts = torch.zeros((plan.n_if, plan.max_sharers, 3), dtype=q.dtype, device=q.device)
    parts[torch.arange(plan.n_if, device=q.device), plan.my_slot] = q[plan.if_idx]
    sends = [q[loc].contiguous() for (_, loc, _, _) in plan.nb]
    recvs = [torch.empty_like(s) for s in sends]
    ops = []
    for (s, _, _, _), sb, rb in zip(plan.nb, sends, recvs):
        ops.append(dist.P2POp(dist.isend, sb, s, group))
        ops.append(dist.P2POp(dist.irecv, rb, s, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for (_, _, rows, slot), rb in zip(plan.nb, recvs):
        parts[rows, slot] = rb
    acc = parts[:, 0]
    for j in range(1, plan.max_sharers):          # ascending rank order; absent sharers contribute an exact +0
        acc = acc + parts[:, j]
    q[plan.if_idx] = acc
    return q


# ---------------------------------------------------------------------------------------------------
# distributed solver node: EulerImplicitSolver + CGLinearSolver over the partitioned mesh
# ---------------------------------------------------------------------------------------------------
class DeviceBackend:
    """The rank-local operators on the GPU (every method is one C-ABI call on the rank's context)."""

    def __init__(self, ctx, template, rm, m_pass, fixed_local, youngModulus, poissonRatio, method, params):
        from . import components as C
        self.ctx = ctx
        self.mo = C.MechanicalObject(ctx, template, position=rm.positions)
        # the partition-interface nodes take the staging path, so that one thread per node owns their halo exchange (peer mode)
        self.ff = C.TetrahedronFEMForceField(self.mo, rm.elems, youngModulus=youngModulus, poissonRatio=poissonRatio, method=method, sharedNodes=rm.interface)
        self.mass = C.DiagonalMass(self.mo, vertexMass=m_pass)
        self.fix = C.FixedProjectiveConstraint(self.mo, fixed_local)
        self.node = C.SolverNode(self.mo, self.ff, self.mass, self.fix, dt=params["dt"], gravity=params["gravity"], rayleighStiffness=params["rK"],
                                 rayleighMass=params["rM"], iterations=params["iterations"], tolerance=params["tolerance"], threshold=params["threshold"])
        self.device = ctx.device
        self.dtype = self.mo.tdtype
        self.x, self.v = self.mo.x, self.mo.v
        self.owned_mask = torch.from_numpy(rm.owned.astype(np.uint8)).to(self.device)
        self._scal = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.native = False

    def attach_native(self, rm, group=None):
        """Run the whole distributed step inside the library: NCCL send/recv + allreduce enqueued by libsofa_b200 itself,
        device-resident CG scalars, one CUDA graph per step (sofab200_node_set_distributed)."""
        from . import components as C
        self.comm = C.Communicator(self.ctx, group)
        self.node.set_distributed(self.comm, rm)
        self.native = True
        self.peer = False
        import os
        if os.environ.get("SOFAB200_PEER", "1") != "0" and rm.world <= 8:
            self.peer = self._attach_peer(rm, group)

    def _attach_peer(self, rm, group):
        """Peer-memory mode: every rank's mailbox mapped into every process with CUDA IPC (torch.distributed only ships the
        64-byte handles), then the CG loop is one persistent kernel per GPU that exchanges over NVLink by itself."""
        rows = [None] * rm.world
        dist.all_gather_object(rows, int(sum(len(v["rows"]) for v in rm.neighbours.values())), group=group)
        inbox_rows = max(rows)
        try:
            ptr, handle = self.ctx.peer_alloc(self.node.peer_bytes(inbox_rows))
        except Exception:
            handle, ptr = None, None
        handles = [None] * rm.world
        dist.all_gather_object(handles, handle, group=group)
        # first inbox row of each neighbour's block on this rank (neighbours in ascending rank order, like set_distributed)
        nbs = sorted(rm.neighbours.items())
        offs, acc = {}, 0
        for s, v in nbs:
            offs[s] = acc; acc += len(v["rows"])
        all_offs = [None] * rm.world
        dist.all_gather_object(all_offs, offs, group=group)
        ok = all(h is not None for h in handles)
        bases = []
        if ok:
            try:
                bases = [ptr if r == rm.rank else self.ctx.peer_open(handles[r]) for r in range(rm.world)]
            except Exception:
                ok = False
        flags = [None] * rm.world
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            return False
        remote_off = [all_offs[s][rm.rank] for s, _ in nbs]
        try:
            self.node.set_peer(rm.rank, rm.world, bases, inbox_rows, remote_off)
            ok = True
        except Exception as e:  # e.g. the partition does not fit the persistent kernel's assumptions
            ok = False
        dist.all_gather_object(flags, ok, group=group)
        if not all(flags):
            self.node.clear_peer()       # every rank must run the same loop
        dist.barrier(group=group)
        return all(flags)

    def new_vector(self):
        return self.mo.new_vector()

    def vop(self, r, a=None, b=None, k=1.0):
        self.mo.vOp(r, a, b, k)

    def dot_owned(self, a, b):
        self.mo.vDot_dev(a, b, self._scal, self.owned_mask)
        return self._scal

    def compute_force(self, f, x):
        self.node.computeForce(f, x)

    def add_mbkdx(self, out, d, m, b, k, init=None, scale=None, project=False):
        self.node.addMBKdx(out, d, m, b, k, init=init, scale=scale, project=project)

    def integrate(self, v, x, a, h):
        self.mo.vMultiOp_integrate(v, x, a, 1.0, h)


class DistributedSolverNode:
    """One solver node over a mesh partitioned across the ranks of `group`.  Same Data as SolverNode; the CG loop follows
    CGLinearSolver.inl:73-315 with its scalars all-reduced (two small allreduces and one halo exchange per iteration)."""

    def __init__(self, positions, tets, fixed_global, massDensity, youngModulus, poissonRatio, method="large", group=None, ctx=None,
                 template="B200Vec3f", backend_factory=None, native=True, partition="slab", dt=0.01, gravity=(0.0, -9.81, 0.0), rayleighStiffness=0.0, rayleighMass=0.0,
                 iterations=25, tolerance=1e-5, threshold=1e-5):
        from .topology import diagonal_mass
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rm = rm = RankMesh(positions, tets, self.rank, self.world, partition)
        self.p = dict(dt=dt, gravity=tuple(gravity), rK=rayleighStiffness, rM=rayleighMass, iterations=iterations, tolerance=tolerance, threshold=threshold)
        ndtype = np.float32 if template.endswith("f") else np.float64
        # global lumped masses (identical on every sharing rank); the fused element pass of a rank adds the mass term of an
        # interface node only where that rank owns it
        m_global = diagonal_mass(positions, tets, ndtype, mass_density=massDensity)
        m_pass = m_global[rm.global_ids].copy(); m_pass[~rm.owned] = 0
        fixed_mask = np.zeros(positions.shape[0], bool); fixed_mask[np.asarray(fixed_global, np.int64)] = True
        fixed_local = np.nonzero(fixed_mask[rm.global_ids])[0].astype(np.uint32)
        make = backend_factory or (lambda **kw: DeviceBackend(ctx, template, **kw))
        self.be = be = make(rm=rm, m_pass=m_pass, fixed_local=fixed_local, youngModulus=youngModulus, poissonRatio=poissonRatio, method=method, params=self.p)
        if native and backend_factory is None:
            be.attach_native(rm, group)
        self.plan = HaloPlan(rm, be.device)
        self.owner_scale = torch.from_numpy(rm.owned.astype(ndtype)).to(be.device)[:, None]
        self.f, self.b, self.dx = be.new_vector(), be.new_vector(), be.new_vector()
        self.pv, self.q, self.r = be.new_vector(), be.new_vector(), be.new_vector()
        self.time_step_count = 0
        self.last_iterations = 0
        self.cg_iterations_total = 0

    # -- distributed pieces ---------------------------------------------------------------------------
    def _dot(self, a, b):
        """vDot over owned nodes, all-reduced; returned as a host double (one synchronisation, like the reference's vDot)."""
        s = self.be.dot_owned(a, b)
        dist.all_reduce(s, group=self.group)
        return float(s.item())

    def apply(self, q, p, m, b, k):
        """q = project((m M + b B + k K) p) over the whole partitioned mesh."""
        if getattr(self.be, "native", False):
            self.be.node.apply(q, p, m, b, k)
            return q
        self.be.add_mbkdx(q, p, m, b, k, project=True)
        halo_exchange_sum(q, self.plan, self.group)
        return q

    def cg_solve(self, x, bvec, m, bfac, k):
        """CGLinearSolver::solve with distributed dots (CGLinearSolver.inl:94-272)."""
        be, P = self.be, self.p
        be.vop(x); be.vop(self.r, bvec)                                   # x = 0 ; r = b
        normb = np.sqrt(self._dot(bvec, bvec))
        nb_iter, rho_1 = 0, 0.0
        if normb != 0.0:
            nb_iter = 1
            while nb_iter <= P["iterations"]:
                rho = self._dot(self.r, self.r)
                err = np.sqrt(rho) / normb
                if err <= P["tolerance"] and not (nb_iter == 1 and self.time_step_count == 0):
                    break
                if nb_iter == 1:
                    be.vop(self.pv, self.r)                                # p = r
                else:
                    be.vop(self.pv, self.r, self.pv, rho / rho_1)          # p = r + p*beta
                self.apply(self.q, self.pv, m, bfac, k)
                den = self._dot(self.pv, self.q)
                if den == 0.0:
                    break
                if abs(den) <= P["threshold"] and not (nb_iter == 1 and self.time_step_count == 0):
                    break
                alpha = rho / den
                be.vop(x, x, self.pv, alpha)
                be.vop(self.r, self.r, self.q, -alpha)
                rho_1 = rho
                nb_iter += 1
        self.time_step_count += 1
        self.last_iterations = nb_iter
        self.cg_iterations_total += min(nb_iter, P["iterations"])
        return nb_iter

    def step(self):
        """EulerImplicitSolver::solve (EulerImplicitSolver.cpp:127-300) on the partitioned mesh."""
        be, P = self.be, self.p
        if getattr(be, "native", False):      # everything below, inside the library (no host round trip)
            be.node.step()
            self.time_step_count += 1
            self.cg_iterations_total += P["iterations"]     # upper bound; the exact count is last_solve()["iterations"]
            return None
        h = P["dt"]
        be.compute_force(self.f, be.x)                                     # gravity*m (owner only) + local element forces
        halo_exchange_sum(self.f, self.plan, self.group)
        # b = (f + (-rM M + (h + rK) K) v) * h, projected.  The start value f of an interface node enters on its owner only.
        f_init = self.f * self.owner_scale if self.plan.n_if else self.f
        be.add_mbkdx(self.b, be.v, -P["rM"], 0.0, h + P["rK"], init=f_init, scale=h, project=True)
        halo_exchange_sum(self.b, self.plan, self.group)
        it = self.cg_solve(self.dx, self.b, 1 + h * P["rM"], -h, -h * (h + P["rK"]))
        be.integrate(be.v, be.x, self.dx, h)
        return it

    def gather_global(self, local_vec, n_global):
        """Assemble a global [n_global,3] array on every rank from the owners' rows (tests / checkpoints)."""
        out = torch.zeros((n_global, 3), dtype=torch.float64)
        rows = torch.from_numpy(self.rm.global_ids[self.rm.owned])
        out[rows] = local_vec.detach().cpu().double()[torch.from_numpy(np.nonzero(self.rm.owned)[0])]
        if out.device != torch.device("cpu") or dist.get_backend(self.group) == "gloo":
            dist.all_reduce(out, group=self.group)
        else:
            t = out.to(self.be.device); dist.all_reduce(t, group=self.group); out = t.cpu()
        return out.numpy()
