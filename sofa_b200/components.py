"""Host-side mirror of the reference components on the hot path (names, Data fields and call
signatures follow the reference; every method forwards to one C-ABI entry point of libsofa_b200.so).

  MechanicalObject            Sofa/Component/StateContainer/src/sofa/component/statecontainer/MechanicalObject.inl
  TetrahedronFEMForceField    Sofa/Component/SolidMechanics/FEM/Elastic/src/.../TetrahedronFEMForceField.{h,inl}
  HexahedronFEMForceField     .../HexahedronFEMForceField.{h,inl}
  DiagonalMass                Sofa/Component/Mass/src/sofa/component/mass/DiagonalMass.inl
  FixedProjectiveConstraint   Sofa/Component/Constraint/Projective/src/.../FixedProjectiveConstraint.inl
  SolverNode                  EulerImplicitSolver.cpp:83-341 + CGLinearSolver.inl:73-315 + GraphScatteredTypes.cpp:33-46

Vectors are torch CUDA tensors of shape [N,3] (float32 for template "B200Vec3f", float64 for "B200Vec3d"):
PyTorch only owns the memory and the stream.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import F32, F64, check

TEMPLATES = {"B200Vec3f": (F32, torch.float32, np.float32), "B200Vec3d": (F64, torch.float64, np.float64),
             "Vec3f": (F32, torch.float32, np.float32), "Vec3d": (F64, torch.float64, np.float64), "Vec3": (F64, torch.float64, np.float64)}
TET_METHODS = {"small": 0, "large": 1, "polar": 2, "svd": 3}
HEX_METHODS = {"large": 0, "polar": 1, "small": 2}
_P = C.c_void_p


def _dptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _darr(a):
    a = np.ascontiguousarray(np.atleast_1d(np.asarray(a, np.float64)))
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class Context:
    """One device + one stream (sofab200_ctx).  Uses torch's current stream of that device."""

    def __init__(self, device=0, stream=None):
        self.L = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.Sofab200Error("no CUDA device: sofa_b200 has no CPU path")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.h = _P()
        # torch's default stream is the legacy NULL stream: hand the library the explicit cudaStreamLegacy handle (0x1) so that
        # it enqueues on the SAME stream as torch's copies instead of creating its own (a NULL handle means "create one").
        handle = self.stream.cuda_stream or 1
        check(self.L.sofab200_ctx_create(device, C.c_void_p(handle), C.byref(self.h)))

    def synchronize(self):
        check(self.L.sofab200_ctx_synchronize(self.h))

    @property
    def launch_count(self):
        return int(self.L.sofab200_ctx_launch_count(self.h))

    PROFILE_CLASSES = ("element_pass_dforce", "boundary_gather", "element_pass_force", "cg_vector_kernels", "cg_persistent")

    def profile_begin(self):
        check(self.L.sofab200_ctx_profile_begin(self.h))

    def profile_end(self):
        ms = (C.c_double * 5)(); cnt = (C.c_uint64 * 5)()
        check(self.L.sofab200_ctx_profile_end(self.h, ms, cnt))
        return {k: dict(ms=ms[i], launches=int(cnt[i])) for i, k in enumerate(self.PROFILE_CLASSES)}

    def peer_alloc(self, nbytes):
        """A zeroed mailbox in device memory + its 64-byte CUDA IPC handle (bytes) for the other ranks of the node."""
        ptr = _P(); h = (C.c_ubyte * 64)()
        check(self.L.sofab200_peer_alloc(self.h, int(nbytes), C.byref(ptr), h))
        return ptr.value, bytes(h)

    def peer_open(self, handle):
        ptr = _P(); h = (C.c_ubyte * 64)(*handle)
        check(self.L.sofab200_peer_open(self.h, h, C.byref(ptr)))
        return ptr.value

    def trace_begin(self):
        check(self.L.sofab200_ctx_trace_begin(self.h))

    def trace_end(self):
        """(element_pass[4096, 16], cg_tail[4096, 16]) uint64 %globaltimer ns per CTA and phase boundary; column 7 = SM id."""
        import numpy as np
        out = np.zeros(2 * 4096 * 16, dtype=np.uint64)
        check(self.L.sofab200_ctx_trace_end(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64)), out.size))
        return out[:4096 * 16].reshape(4096, 16), out[4096 * 16:].reshape(4096, 16)

    def __del__(self):
        try:
            self.L.sofab200_ctx_destroy(self.h)
        except Exception:
            pass


class Communicator:
    """An NCCL communicator of the library bound to a Context (sofab200_comm).  The unique id is created on rank 0 and
    shipped with torch.distributed, which is only the bootstrap channel."""

    def __init__(self, ctx, group=None):
        import torch.distributed as dist
        self.ctx = ctx
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            check(ctx.L.sofab200_comm_get_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            t = ident.to(ctx.device); dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group); ident = t.cpu()
        else:
            dist.broadcast(ident, src=0, group=group)
        raw = (C.c_ubyte * 128)(*ident.tolist())
        self.h = _P()
        check(ctx.L.sofab200_comm_create(ctx.h, self.world, self.rank, raw, C.byref(self.h)))

    def __del__(self):
        try:
            self.ctx.L.sofab200_comm_destroy(self.h)
        except Exception:
            pass


class MechanicalObject:
    """MechanicalObject<B200Vec3Types>: state vectors + vOp / vMultiOp / vDot."""

    def __init__(self, ctx, template="B200Vec3f", position=None, velocity=None, rest_position=None):
        self.ctx = ctx
        self.real, self.tdtype, self.ndtype = TEMPLATES[template]
        self.template = template
        pos = np.ascontiguousarray(position, self.ndtype).reshape(-1, 3)
        self.size = pos.shape[0]
        self.rest_position_host = pos.copy() if rest_position is None else np.ascontiguousarray(rest_position, self.ndtype).reshape(-1, 3)
        self.x = torch.from_numpy(pos).to(ctx.device)
        self.v = torch.zeros_like(self.x) if velocity is None else torch.from_numpy(np.ascontiguousarray(velocity, self.ndtype)).to(ctx.device)
        self.f = torch.zeros_like(self.x)
        self.dx = torch.zeros_like(self.x)

    def new_vector(self):
        return torch.zeros_like(self.x)

    def vOp(self, r, a=None, b=None, k=1.0):
        check(self.ctx.L.sofab200_mo_vop(self.ctx.h, self.real, self.size, _dptr(r), _dptr(a), _dptr(b), float(k)))
        return r

    def vDot(self, a, b):
        out = C.c_double()
        check(self.ctx.L.sofab200_mo_vdot(self.ctx.h, self.real, self.size, _dptr(a), _dptr(b), C.byref(out)))
        return out.value

    def vDot_dev(self, a, b, result, node_mask=None):
        """Masked dot product left in the 1-element float64 device tensor `result` (asynchronous)."""
        check(self.ctx.L.sofab200_mo_vdot_dev(self.ctx.h, self.real, self.size, _dptr(a), _dptr(b), _dptr(node_mask), _dptr(result)))

    def vMultiOp_integrate(self, v, x, a, f_v_a=1.0, f_x_v=1.0):
        check(self.ctx.L.sofab200_mo_vmultiop_integrate(self.ctx.h, self.real, self.size, _dptr(v), _dptr(x), _dptr(a), float(f_v_a), float(f_x_v)))

    def resetForce(self, f=None):
        return self.vOp(self.f if f is None else f)

    def accumulateForce(self, f, externalForce):
        """MechanicalObject::accumulateForce (MechanicalObject.inl:1356-1375): f[i] += externalForce[i] for the rows that differ from Deriv()."""
        check(self.ctx.L.sofab200_mo_accumulate_force(self.ctx.h, self.real, self.size, _dptr(f), _dptr(externalForce)))
        return f


class TetrahedronFEMForceField:
    """TetrahedronFEMForceField<B200Vec3Types>.  Data: youngModulus, poissonRatio, method, localStiffnessFactor,
    rayleighStiffness (BaseForceField), plasticMaxThreshold / plasticYieldThreshold / plasticCreep (defaults are the reference's
    (Real)0.0001f and (Real)0.9f, TetrahedronFEMForceField.inl:51-53)."""

    def __init__(self, mstate, tetrahedra, youngModulus=5000.0, poissonRatio=0.45, method="large", localStiffnessFactor=None,
                 rayleighStiffness=0.0, tileElems=0, sharedNodes=None, plasticMaxThreshold=0.0, plasticYieldThreshold=float(np.float32(0.0001)), plasticCreep=float(np.float32(0.9)), computeVonMisesStress=0, updateStiffnessMatrix=False):
        if method not in TET_METHODS:
            raise ValueError(f"method must be one of {list(TET_METHODS)}")
        self.mstate, self.ctx = mstate, mstate.ctx
        self.method, self.rayleighStiffness = method, float(rayleighStiffness)
        self.tetrahedra = np.ascontiguousarray(tetrahedra, np.uint32).reshape(-1, 4)
        y, yp = _darr(youngModulus); p, pp = _darr(poissonRatio)
        d = _lib.TetFemDesc(); d.method = TET_METHODS[method]
        d.n_young, d.young, d.n_poisson, d.poisson = len(y), yp, len(p), pp
        if localStiffnessFactor is not None:
            l, lp = _darr(localStiffnessFactor); d.n_local_stiffness, d.local_stiffness = len(l), lp
        d.tile_elems = int(tileElems)
        d.compute_von_mises = int(computeVonMisesStress)
        d.update_stiffness_matrix = int(bool(updateStiffnessMatrix))
        d.tetrahedral_corotational = int(getattr(self, "_tetrahedral_corotational", False))
        d.plastic_max_threshold, d.plastic_yield_threshold, d.plastic_creep = float(plasticMaxThreshold), float(plasticYieldThreshold), float(plasticCreep)
        if sharedNodes is not None:      # nodes that must take the staging path (partition interface of a multi-GPU run)
            self._shared = np.zeros(mstate.size, np.uint8); self._shared[np.asarray(sharedNodes, np.int64)] = 1
            d.shared_nodes = self._shared.ctypes.data_as(C.POINTER(C.c_ubyte))
        self.h = _P()
        rest = mstate.rest_position_host
        check(self.ctx.L.sofab200_tetfem_create(self.ctx.h, mstate.real, mstate.size, rest.ctypes.data_as(_P), self.tetrahedra.shape[0],
                                                self.tetrahedra.ctypes.data_as(_P), C.byref(d), C.byref(self.h)))

    def addForce(self, f, x, v=None):
        check(self.ctx.L.sofab200_tetfem_add_force(self.h, _dptr(f), _dptr(x)))

    def addDForce(self, df, dx, kFactor=1.0, bFactor=0.0):
        """kFactor/bFactor as in MechanicalParams; the effective factor is kFactor + bFactor*rayleighStiffness."""
        check(self.ctx.L.sofab200_tetfem_add_dforce(self.h, _dptr(df), _dptr(dx), float(kFactor) + float(bFactor) * self.rayleighStiffness))

    def get(self, what):
        T = self.tetrahedra.shape[0]
        shape = {"rotations": (T, 3, 3), "initialRotations": (T, 3, 3), "strainDisplacements": (T, 4, 3), "materialsStiffnesses": (T, 3),
                 "rotatedInitialElements": (T, 4, 3), "initialTransformation": (T, 3, 3), "plasticStrains": (T, 6)}[what]
        out = np.empty(shape, self.mstate.ndtype)
        check(self.ctx.L.sofab200_tetfem_get(self.h, what.encode(), out.ctypes.data_as(_P)))
        return out

    def computeVonMisesStress(self, x):
        """computeVonMisesStress() (TetrahedronFEMForceField.inl:2196-2416) at positions x -> (vonMisesPerElement, vonMisesPerNode) device tensors."""
        import torch
        pe = torch.empty(self.tetrahedra.shape[0], dtype=self.mstate.tdtype, device=self.ctx.device)
        pn = torch.empty(self.mstate.size, dtype=self.mstate.tdtype, device=self.ctx.device)
        check(self.ctx.L.sofab200_tetfem_compute_von_mises(self.h, _dptr(x), _dptr(pe), _dptr(pn)))
        return pe, pn

    def reset(self):
        """reset() (TetrahedronFEMForceField.inl:1380-1388): clears the plastic strains."""
        check(self.ctx.L.sofab200_tetfem_reset(self.h))

    def getRotations(self, vecR=None):
        """getRotations(VecReal&) (TetrahedronFEMForceField.inl:2033-2042): per-node rotations, n x 3 x 3 device tensor."""
        import torch
        if vecR is None:
            vecR = torch.empty((self.mstate.size, 3, 3), dtype=self.mstate.tdtype, device=self.ctx.device)
        check(self.ctx.L.sofab200_tetfem_get_rotations(self.h, _dptr(vecR)))
        return vecR

    def stats(self):
        out = (C.c_uint64 * 8)()
        check(self.ctx.L.sofab200_tetfem_stats(self.h, out))
        return dict(zip(["tiles", "tile_elems", "interior_nodes", "shared_nodes", "staged_corners", "smem_bytes", "max_valence", "n_elems"], list(out)))

    def __del__(self):
        try:
            self.ctx.L.sofab200_tetfem_destroy(self.h)
        except Exception:
            pass


class TetrahedralCorotationalFEMForceField(TetrahedronFEMForceField):
    """TetrahedralCorotationalFEMForceField<B200Vec3Types> (what Demos/liver.scn uses): methods small / large / polar, Data youngModulus,
    poissonRatio, localStiffnessFactor, updateStiffnessMatrix.  Its arithmetic is statement for statement TetrahedronFEMForceField's
    (TetrahedralCorotationalFEMForceField.inl:356-1175), so the same device object serves it (descriptor flag tetrahedral_corotational)."""
    _tetrahedral_corotational = True

    def __init__(self, mstate, tetrahedra, youngModulus=5000.0, poissonRatio=0.45, method="large", localStiffnessFactor=None, rayleighStiffness=0.0,
                 updateStiffnessMatrix=False, tileElems=0):
        if method == "svd":
            raise ValueError("TetrahedralCorotationalFEMForceField has no svd method")
        super().__init__(mstate, tetrahedra, youngModulus, poissonRatio, method, localStiffnessFactor, rayleighStiffness, tileElems,
                         updateStiffnessMatrix=updateStiffnessMatrix)


FAST_METHODS = {"qr": 1, "large": 1, "polar": 2, "polar2": 4, "none": 0, "linear": 0, "small": 0}      # d_method strings -> sofab200_tet_method


class FastTetrahedralCorotationalForceField(TetrahedronFEMForceField):
    """FastTetrahedralCorotationalForceField<B200Vec3Types> (FastTetrahedralCorotationalForceField.inl): Data youngModulus, poissonRatio, method
    ("qr" -- the default --, "polar", "polar2", "none" and their synonyms).  addDForce runs over the topology's edges; `edges` is the topology's
    own edge list when it has one, else the edges are numbered as TetrahedronSetTopologyContainer does.  Same handle type and calls as
    TetrahedronFEMForceField, so a SolverNode takes it as its force field."""

    def __init__(self, mstate, tetrahedra, youngModulus=5000.0, poissonRatio=0.45, method="qr", rayleighStiffness=0.0, edges=None, tileElems=0):
        if method not in FAST_METHODS:
            raise ValueError(f"method must be one of {list(FAST_METHODS)}")
        self.mstate, self.ctx = mstate, mstate.ctx
        self.method, self.rayleighStiffness = method, float(rayleighStiffness)
        self.tetrahedra = np.ascontiguousarray(tetrahedra, np.uint32).reshape(-1, 4)
        y, yp = _darr(youngModulus); p, pp = _darr(poissonRatio)
        d = _lib.TetFemDesc(); d.method = FAST_METHODS[method]
        d.n_young, d.young, d.n_poisson, d.poisson = len(y), yp, len(p), pp
        d.tile_elems = int(tileElems)
        d.fast_corotational = 1
        if edges is not None:
            self._edges = np.ascontiguousarray(edges, np.uint32).reshape(-1, 2)
            d.n_edges, d.edges = self._edges.shape[0], self._edges.ctypes.data_as(C.POINTER(C.c_uint32))
        self.h = _P()
        rest = mstate.rest_position_host
        check(self.ctx.L.sofab200_tetfem_create(self.ctx.h, mstate.real, mstate.size, rest.ctypes.data_as(_P), self.tetrahedra.shape[0],
                                                self.tetrahedra.ctypes.data_as(_P), C.byref(d), C.byref(self.h)))

    def get(self, what):
        T = self.tetrahedra.shape[0]
        if what in ("edges", "n_edges", "edgeInfo"):
            n = np.zeros(1, np.uint64)
            check(self.ctx.L.sofab200_tetfem_get(self.h, b"n_edges", n.ctypes.data_as(_P)))
            if what == "n_edges":
                return int(n[0])
            out = np.empty((int(n[0]), 2), np.uint32) if what == "edges" else np.empty((int(n[0]), 3, 3), self.mstate.ndtype)
        else:
            shape = {"rotations": (T, 3, 3), "restRotations": (T, 3, 3), "shapeVectors": (T, 4, 3), "linearDfDx": (T, 6, 3, 3), "linearDfDxDiag": (T, 4, 3, 3),
                     "restEdgeVectors": (T, 6, 3), "edgeOrientations": (T, 6)}[what]
            out = np.empty(shape, self.mstate.ndtype)
        check(self.ctx.L.sofab200_tetfem_get(self.h, what.encode(), out.ctypes.data_as(_P)))
        return out


class HexahedronFEMForceField:
    """HexahedronFEMForceField<B200Vec3Types>.  Data: youngModulus, poissonRatio, method (large | polar | small)."""

    def __init__(self, mstate, hexahedra, youngModulus=5000.0, poissonRatio=0.45, method="large", rayleighStiffness=0.0, tileElems=0):
        if method not in HEX_METHODS:
            raise ValueError(f"method must be one of {list(HEX_METHODS)}")
        self.mstate, self.ctx = mstate, mstate.ctx
        self.method, self.rayleighStiffness = method, float(rayleighStiffness)
        self.hexahedra = np.ascontiguousarray(hexahedra, np.uint32).reshape(-1, 8)
        y, yp = _darr(youngModulus); p, pp = _darr(poissonRatio)
        d = _lib.HexFemDesc(); d.method = HEX_METHODS[method]
        d.n_young, d.young, d.n_poisson, d.poisson = len(y), yp, len(p), pp
        d.tile_elems = int(tileElems)
        self.h = _P()
        rest = mstate.rest_position_host
        check(self.ctx.L.sofab200_hexfem_create(self.ctx.h, mstate.real, mstate.size, rest.ctypes.data_as(_P), self.hexahedra.shape[0],
                                                self.hexahedra.ctypes.data_as(_P), C.byref(d), C.byref(self.h)))

    def addForce(self, f, x, v=None):
        check(self.ctx.L.sofab200_hexfem_add_force(self.h, _dptr(f), _dptr(x)))

    def addDForce(self, df, dx, kFactor=1.0, bFactor=0.0):
        check(self.ctx.L.sofab200_hexfem_add_dforce(self.h, _dptr(df), _dptr(dx), float(kFactor) + float(bFactor) * self.rayleighStiffness))

    def get(self, what):
        H = self.hexahedra.shape[0]
        shape = {"rotations": (H, 3, 3), "elementStiffnesses": (H, 24, 24), "rotatedInitialElements": (H, 8, 3)}[what]
        out = np.empty(shape, self.mstate.ndtype)
        check(self.ctx.L.sofab200_hexfem_get(self.h, what.encode(), out.ctypes.data_as(_P)))
        return out

    def getRotations(self, vecR=None):
        """getNodeRotation for every node (HexahedronFEMForceField.inl:946-1023): n x 3 x 3 device tensor."""
        if vecR is None:
            vecR = torch.empty((self.mstate.size, 3, 3), dtype=self.mstate.tdtype, device=self.ctx.device)
        check(self.ctx.L.sofab200_hexfem_get_rotations(self.h, _dptr(vecR)))
        return vecR

    def stats(self):
        out = (C.c_uint64 * 8)()
        check(self.ctx.L.sofab200_hexfem_stats(self.h, out))
        return dict(zip(["tiles", "tile_elems", "interior_nodes", "shared_nodes", "staged_corners", "smem_bytes", "unique_stiffness_matrices", "n_elems"], list(out)))

    def __del__(self):
        try:
            self.ctx.L.sofab200_hexfem_destroy(self.h)
        except Exception:
            pass


class DiagonalMass:
    """DiagonalMass<B200Vec3Types>: Data vertexMass | massDensity | totalMass (+ rayleighMass of Mass)."""

    def __init__(self, mstate, elements=None, vertexMass=None, massDensity=None, totalMass=None, rayleighMass=0.0):
        from .topology import diagonal_mass
        self.mstate, self.ctx = mstate, mstate.ctx
        self.rayleighMass = float(rayleighMass)
        if vertexMass is not None:
            self.vertexMass_host = np.ascontiguousarray(vertexMass, mstate.ndtype)
        else:
            self.vertexMass_host = diagonal_mass(mstate.rest_position_host, elements, mstate.ndtype, massDensity, totalMass)
        self.vertexMass = torch.from_numpy(self.vertexMass_host).to(self.ctx.device)

    def addMDx(self, res, dx, factor=1.0):
        check(self.ctx.L.sofab200_mass_add_mdx(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(res), _dptr(dx), _dptr(self.vertexMass), float(factor)))

    def addForce(self, f, gravity):
        g = (C.c_double * 3)(*[float(v) for v in gravity])
        check(self.ctx.L.sofab200_mass_add_force(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(f), _dptr(self.vertexMass), g))

    def accFromF(self, a, f):
        check(self.ctx.L.sofab200_mass_acc_from_f(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(a), _dptr(f), _dptr(self.vertexMass)))


class MeshMatrixMass:
    """MeshMatrixMass<B200Vec3Types> on a tetrahedral topology: Data massDensity, lumping (vertexMass / edgeMass computed as the CPU
    class does, MeshMatrixMass.inl:547-665), or vertexMass + edges + edgeMass given."""

    def __init__(self, mstate, tetrahedra=None, massDensity=1.0, lumping=False, vertexMass=None, edges=None, edgeMass=None, massLumpingCoeff=2.5, rayleighMass=0.0):
        from .topology import mesh_matrix_mass
        self.mstate, self.ctx = mstate, mstate.ctx
        self.lumping, self.rayleighMass = bool(lumping), float(rayleighMass)
        if vertexMass is None:
            vertexMass, edges, edgeMass, massLumpingCoeff = mesh_matrix_mass(mstate.rest_position_host, tetrahedra, mstate.ndtype, massDensity, lumping)
        self.vertexMass_host = np.ascontiguousarray(vertexMass, mstate.ndtype)
        self.edges = np.ascontiguousarray(edges if edges is not None else np.zeros((0, 2)), np.uint32).reshape(-1, 2)
        self.edgeMass_host = np.ascontiguousarray(edgeMass if edgeMass is not None else np.zeros(0), mstate.ndtype)
        self.massLumpingCoeff = float(massLumpingCoeff)
        self.h = _P()
        check(self.ctx.L.sofab200_meshmass_create(self.ctx.h, mstate.real, mstate.size, self.vertexMass_host.ctypes.data_as(_P), self.edges.shape[0],
                                                  self.edges.ctypes.data_as(_P), self.edgeMass_host.ctypes.data_as(_P), int(self.lumping),
                                                  self.massLumpingCoeff, C.byref(self.h)))

    def addMDx(self, res, dx, factor=1.0):
        check(self.ctx.L.sofab200_meshmass_add_mdx(self.h, _dptr(res), _dptr(dx), float(factor)))

    def addForce(self, f, gravity):
        g = (C.c_double * 3)(*[float(v) for v in gravity])
        check(self.ctx.L.sofab200_meshmass_add_force(self.h, _dptr(f), g))

    def accFromF(self, a, f):
        check(self.ctx.L.sofab200_meshmass_acc_from_f(self.h, _dptr(a), _dptr(f)))

    def __del__(self):
        try:
            self.ctx.L.sofab200_meshmass_destroy(self.h)
        except Exception:
            pass


class UniformMass:
    """UniformMass<B200Vec3Types>: Data vertexMass | totalMass (+ rayleighMass of Mass); Mass/.../UniformMass.inl:300-345,403-496."""

    def __init__(self, mstate, vertexMass=None, totalMass=None, rayleighMass=0.0):
        self.mstate, self.ctx = mstate, mstate.ctx
        self.rayleighMass = float(rayleighMass)
        real = np.dtype(mstate.ndtype).type
        if totalMass is not None:      # initFromTotalMass: *m = d_totalMass.getValue() / Real(size), stored as MassType = Real
            self.vertexMass_value = float(real(float(totalMass) / float(real(mstate.size))))
        else:
            self.vertexMass_value = float(real(1.0 if vertexMass is None else vertexMass))
        self.vertexMass_host = np.full(mstate.size, self.vertexMass_value, mstate.ndtype)

    def addMDx(self, res, dx, factor=1.0):
        check(self.ctx.L.sofab200_uniform_mass_add_mdx(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(res), _dptr(dx), self.vertexMass_value, float(factor)))

    def addForce(self, f, gravity):
        g = (C.c_double * 3)(*[float(v) for v in gravity])
        check(self.ctx.L.sofab200_uniform_mass_add_force(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(f), self.vertexMass_value, g))


class PlaneForceField:
    """PlaneForceField<B200Vec3Types>: Data normal, d, stiffness, damping, maxForce, bilateral (+ rayleighStiffness of BaseForceField);
    per-operation level (MechanicalLoad/.../PlaneForceField.inl:158-226)."""

    def __init__(self, mstate, normal=(0.0, 1.0, 0.0), d=0.0, stiffness=500.0, damping=5.0, maxForce=0.0, bilateral=False, rayleighStiffness=0.0):
        self.mstate, self.ctx = mstate, mstate.ctx
        self.rayleighStiffness = float(rayleighStiffness)
        self.desc = _lib.PlaneDesc()
        self.desc.normal = (C.c_double * 3)(*[float(v) for v in normal])
        self.desc.d, self.desc.stiffness, self.desc.damping, self.desc.max_force, self.desc.bilateral = float(d), float(stiffness), float(damping), float(maxForce), int(bilateral)
        self.contacts = torch.zeros(mstate.size, dtype=torch.uint8, device=self.ctx.device)      # m_contacts as a per-node flag

    def addForce(self, f, x, v):
        check(self.ctx.L.sofab200_plane_add_force(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(f), _dptr(x), _dptr(v), C.byref(self.desc), _dptr(self.contacts)))

    def addDForce(self, df, dx, kFactor=1.0, bFactor=0.0):
        check(self.ctx.L.sofab200_plane_add_dforce(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(df), _dptr(dx), C.byref(self.desc), _dptr(self.contacts),
                                                   float(kFactor) + float(bFactor) * self.rayleighStiffness))


class FixedProjectiveConstraint:
    """FixedProjectiveConstraint<B200Vec3Types>: Data indices, fixAll."""

    def __init__(self, mstate, indices=(), fixAll=False):
        self.mstate, self.ctx = mstate, mstate.ctx
        self.indices_host = np.ascontiguousarray(indices, np.uint32)
        self.fixAll = bool(fixAll)
        self.indices = torch.from_numpy(self.indices_host.astype(np.int32)).to(self.ctx.device)

    def projectResponse(self, res):
        check(self.ctx.L.sofab200_fixed_project_response(self.ctx.h, self.mstate.real, self.mstate.size, _dptr(res), len(self.indices_host),
                                                         _dptr(self.indices) if len(self.indices_host) else None, int(self.fixAll)))


class SolverNode:
    """EulerImplicitSolver + CGLinearSolver<GraphScattered> over one MechanicalObject, resident on the device.

    Data (EulerImplicitSolver): rayleighStiffness, rayleighMass, vdamping, firstOrder, trapezoidalScheme;
    Data (CGLinearSolver): iterations, tolerance, threshold, warmStart; context: dt, gravity."""

    def __init__(self, mstate, forcefield, mass=None, constraint=None, massFirst=True, plane=None, dt=0.01, gravity=(0.0, -9.81, 0.0), rayleighStiffness=0.0,
                 rayleighMass=0.0, vdamping=0.0, firstOrder=False, trapezoidalScheme=False, iterations=25, tolerance=1e-5, threshold=1e-5,
                 warmStart=False):
        self.mstate, self.ctx, self.forcefield, self.mass, self.constraint = mstate, mstate.ctx, forcefield, mass, constraint
        d = _lib.NodeDesc()
        if isinstance(forcefield, TetrahedronFEMForceField):
            d.tetfem = forcefield.h
        else:
            d.hexfem = forcefield.h
        if isinstance(mass, UniformMass):
            d.uniform_mass, d.uniform_vertex_mass = 1, mass.vertexMass_value
        elif mass is not None and not isinstance(mass, MeshMatrixMass):
            d.vertex_mass_host = mass.vertexMass_host.ctypes.data_as(_P)
        if constraint is not None:
            d.n_fixed = len(constraint.indices_host)
            d.fixed_host = constraint.indices_host.ctypes.data_as(C.POINTER(C.c_uint32))
            d.fix_all = int(constraint.fixAll)
        d.mass_first = int(massFirst)
        if plane is not None:          # PlaneForceField, the node's last force field: fused into the element passes' epilogue
            self.plane = plane
            d.plane = C.pointer(plane.desc); d.plane_rayleigh_stiffness = plane.rayleighStiffness
        self.h = _P()
        check(self.ctx.L.sofab200_node_create(self.ctx.h, mstate.real, mstate.size, C.byref(d), C.byref(self.h)))
        if isinstance(mass, MeshMatrixMass):
            check(self.ctx.L.sofab200_node_set_mesh_mass(self.h, mass.h))
        self.params = dict(dt=dt, gravity=tuple(gravity), rayleighStiffness=rayleighStiffness, rayleighMass=rayleighMass, vdamping=vdamping,
                           firstOrder=firstOrder, trapezoidalScheme=trapezoidalScheme, iterations=iterations, tolerance=tolerance,
                           threshold=threshold, warmStart=warmStart)
        self._push()

    def _push(self):
        p, s = self.params, _lib.SolverParams()
        s.gravity = (C.c_double * 3)(*p["gravity"]); s.dt = p["dt"]
        s.rayleigh_stiffness, s.rayleigh_mass, s.vdamping = p["rayleighStiffness"], p["rayleighMass"], p["vdamping"]
        s.first_order, s.trapezoidal = int(p["firstOrder"]), int(p["trapezoidalScheme"])
        s.iterations, s.tolerance, s.threshold, s.warm_start = int(p["iterations"]), p["tolerance"], p["threshold"], int(p["warmStart"])
        s.ff_rayleigh_stiffness = self.forcefield.rayleighStiffness
        s.mass_rayleigh_mass = self.mass.rayleighMass if self.mass is not None else 0.0
        check(self.ctx.L.sofab200_node_set_params(self.h, C.byref(s)))

    def set_params(self, **kw):
        for k in kw:
            if k not in self.params:
                raise KeyError(k)
        self.params.update(kw)
        self._push()

    def computeForce(self, f, x):
        check(self.ctx.L.sofab200_node_compute_force(self.h, _dptr(f), _dptr(x)))

    def apply(self, q, p, mFactor, bFactor, kFactor):
        """GraphScatteredMatrix::apply: q = project((m M + b B + k K) p)."""
        check(self.ctx.L.sofab200_node_apply(self.h, _dptr(q), _dptr(p), float(mFactor), float(bFactor), float(kFactor)))

    def addMBKdx(self, out, d, mFactor, bFactor, kFactor, init=None, scale=None, project=False):
        """out = [init +] (m M + b B + k K) d, optionally * scale and projected (mop.addMBKdx / addMBKv)."""
        check(self.ctx.L.sofab200_node_add_mbkdx(self.h, _dptr(out), _dptr(init), _dptr(d), float(mFactor), float(bFactor), float(kFactor),
                                                 int(scale is not None), float(scale if scale is not None else 1.0), int(project)))

    def set_vertex_mass(self, m):
        m = np.ascontiguousarray(m, self.mstate.ndtype)
        check(self.ctx.L.sofab200_node_set_vertex_mass(self.h, m.ctypes.data_as(_P)))

    def set_distributed(self, comm, rank_mesh):
        """Attach an NCCL communicator and the halo plan of a parallel.RankMesh: step / cg_solve / apply / computeForce then run
        the distributed algorithm inside the library."""
        rm = rank_mesh
        owned = np.ascontiguousarray(rm.owned, np.uint8)
        interface = np.ascontiguousarray(rm.interface, np.uint32)
        my_slot = np.ascontiguousarray(rm.my_slot, np.int32)
        nbs = sorted(rm.neighbours.items())
        nb_rank = (C.c_int * max(len(nbs), 1))(*[s for s, _ in nbs])
        nb_count = (C.c_size_t * max(len(nbs), 1))(*[len(v["rows"]) for _, v in nbs])
        rows = [np.ascontiguousarray(v["rows"], np.uint32) for _, v in nbs]
        slots = [np.ascontiguousarray(v["slot"], np.int32) for _, v in nbs]
        rows_p = (C.POINTER(C.c_uint32) * max(len(nbs), 1))(*[r.ctypes.data_as(C.POINTER(C.c_uint32)) for r in rows])
        slots_p = (C.POINTER(C.c_int32) * max(len(nbs), 1))(*[r.ctypes.data_as(C.POINTER(C.c_int32)) for r in slots])
        d = _lib.HaloDesc()
        d.owned = owned.ctypes.data_as(C.POINTER(C.c_ubyte))
        d.n_interface = len(interface); d.interface = interface.ctypes.data_as(C.POINTER(C.c_uint32)); d.my_slot = my_slot.ctypes.data_as(C.POINTER(C.c_int32))
        d.max_sharers = int(rm.max_sharers); d.n_neighbours = len(nbs)
        d.nb_rank, d.nb_count, d.nb_rows, d.nb_slot = nb_rank, nb_count, rows_p, slots_p
        check(self.ctx.L.sofab200_node_set_distributed(self.h, comm.h, C.byref(d)))
        self._comm = comm
        self._nb_ranks = [s for s, _ in nbs]
        self._nb_counts = [len(v["rows"]) for _, v in nbs]

    def peer_bytes(self, inbox_rows):
        return int(self.ctx.L.sofab200_node_peer_bytes(self.h, int(inbox_rows)))

    def set_peer(self, rank, world, peer_bases, inbox_rows, remote_off):
        """Multi-GPU CG in one persistent kernel per GPU over peer memory (sofab200_node_set_peer).  peer_bases: every rank's
        mailbox as mapped here; remote_off[k]: first inbox row of our block in neighbour k's mailbox."""
        d = _lib.PeerDesc(); d.rank, d.world = int(rank), int(world)
        bases = (C.c_void_p * world)(*[int(b) for b in peer_bases])
        offs = (C.c_size_t * max(len(remote_off), 1))(*[int(o) for o in remote_off])
        d.peer_base, d.remote_off, d.inbox_rows = bases, offs, int(inbox_rows)
        check(self.ctx.L.sofab200_node_set_peer(self.h, C.byref(d)))

    def clear_peer(self):
        d = _lib.PeerDesc()
        check(self.ctx.L.sofab200_node_set_peer(self.h, C.byref(d)))

    def cg_solve(self, x, b, mFactor, bFactor, kFactor, sync=True):
        it = C.c_int()
        check(self.ctx.L.sofab200_node_cg_solve(self.h, _dptr(x), _dptr(b), float(mFactor), float(bFactor), float(kFactor), C.byref(it) if sync else None))
        return it.value if sync else None

    def step(self, x=None, v=None):
        """EulerImplicitSolver::solve on the MechanicalObject's device-resident x, v (asynchronous)."""
        check(self.ctx.L.sofab200_node_step(self.h, _dptr(self.mstate.x if x is None else x), _dptr(self.mstate.v if v is None else v)))

    def step_host(self, x_host, v_host):
        """The same step for host arrays (numpy or pinned torch CPU tensors): H2D, step, D2H."""
        xp = x_host.ctypes.data_as(_P) if isinstance(x_host, np.ndarray) else C.c_void_p(x_host.data_ptr())
        vp = v_host.ctypes.data_as(_P) if isinstance(v_host, np.ndarray) else C.c_void_p(v_host.data_ptr())
        check(self.ctx.L.sofab200_node_step_host(self.h, xp, vp))

    def step_host_x(self, x_host, v_host_in=None, v_host_out=None):
        """The step for a host-owned position array: x up, step on the device-resident velocities, x back (v only on request)."""
        ptr = lambda a: None if a is None else (a.ctypes.data_as(_P) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr()))
        check(self.ctx.L.sofab200_node_step_host_x(self.h, ptr(x_host), ptr(v_host_in), ptr(v_host_out)))

    def set_external_force(self, ext_host):
        """MechanicalObject's externalForce (accumulateForce, MechanicalObject.inl:1356-1375): n x 3 host array, or None to remove it."""
        if ext_host is not None:
            ext_host = np.ascontiguousarray(ext_host, self.mstate.ndtype)
        check(self.ctx.L.sofab200_node_set_external_force(self.h, None if ext_host is None else ext_host.ctypes.data_as(_P)))

    def step_pipelined(self, ext_host, x_out_host, x=None, v=None):
        """One step of the device-resident state: this step's external forces up from (pinned) host memory, the new positions down into
        x_out_host on a copy stream while the next step is already being submitted (alternate two output buffers; flush() at the end)."""
        ptr = lambda a: None if a is None else (a.ctypes.data_as(_P) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr()))
        x = self.mstate.x if x is None else x
        v = self.mstate.v if v is None else v
        check(self.ctx.L.sofab200_node_step_pipelined(self.h, _dptr(x), _dptr(v), ptr(ext_host), ptr(x_out_host)))

    def flush(self):
        check(self.ctx.L.sofab200_node_flush(self.h))

    def fused_info(self):
        out = (C.c_int * 8)()
        check(self.ctx.L.sofab200_node_cg_kernel_info(self.h, out))
        return dict(zip(["grid", "tiles_per_cta", "tile_state_in_shared_memory", "dynamic_smem_bytes", "element_threads", "dedicated_shared_node_threads",
                         "fused_enabled", "persistent_enabled"], list(out)))

    def last_solve(self):
        it, ec = C.c_int(), C.c_int()
        ne, nd = C.c_size_t(), C.c_size_t()
        cap = 1100
        ge, gd = (C.c_double * cap)(), (C.c_double * cap)()
        check(self.ctx.L.sofab200_node_last_solve(self.h, C.byref(it), C.byref(ec), ge, C.byref(ne), gd, C.byref(nd), cap))
        return dict(iterations=it.value, end_condition=ec.value, graph_error=np.array(ge[:ne.value]), graph_den=np.array(gd[:nd.value]))

    def get(self, what):
        out = np.empty((self.mstate.size, 3), self.mstate.ndtype)
        check(self.ctx.L.sofab200_node_get(self.h, what.encode(), out.ctypes.data_as(_P)))
        return out

    def get_plane_contacts(self):
        out = np.empty(self.mstate.size, np.uint8)
        check(self.ctx.L.sofab200_node_get(self.h, b"plane_contacts", out.ctypes.data_as(_P)))
        return out

    def reset(self):
        check(self.ctx.L.sofab200_node_reset(self.h))

    def __del__(self):
        try:
            self.ctx.L.sofab200_node_destroy(self.h)
        except Exception:
            pass
