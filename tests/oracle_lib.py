"""ctypes binding of the CPU oracle (oracle/libsofa_oracle.so) and of the reference's own
math compiled from /root/reference (oracle/_ref/libsofa_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under sofa_b200/ may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_ORACLE_SO = os.path.join(ORACLE_DIR, "libsofa_oracle.so")
_REF_SO = os.path.join(ORACLE_DIR, "_ref", "libsofa_ref.so")

TET_METHODS = {"small": 0, "large": 1, "polar": 2, "svd": 3}
FAST_METHODS = {"polar": 0, "qr": 1, "large": 1, "polar2": 2, "none": 3, "linear": 3, "small": 3}
HEX_METHODS = {"large": 0, "polar": 1, "small": 2}
_P = C.c_void_p


def build_oracle():
    """Compile oracle/libsofa_oracle.so (and oracle/_ref when the reference tree exists)."""
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "libsofa_oracle.so"])
    if os.path.isdir("/root/reference/Sofa/framework") and not os.path.exists(_REF_SO):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


_lib = None


def lib():
    global _lib
    if _lib is None:
        src = [os.path.join(ORACLE_DIR, f) for f in ("sofa_oracle.hpp", "sofa_oracle_capi.cpp")]
        if not os.path.exists(_ORACLE_SO) or any(os.path.getmtime(s) > os.path.getmtime(_ORACLE_SO) for s in src):
            build_oracle()
        L = C.CDLL(_ORACLE_SO)
        L.orc_scene_create.restype = _P
        L.orc_scene_get.restype = C.c_size_t
        L.orc_scene_graph.restype = C.c_size_t
        L.orc_vdot.restype = C.c_double
        L.orc_meshmass_create.restype = _P
        L.orc_meshmass_n_edges.restype = C.c_size_t
        L.orc_meshmass_total.restype = C.c_double
        L.orc_scene_hex_potential_energy.restype = C.c_double
        for sfx, ct in (("f", C.c_float), ("d", C.c_double)):
            getattr(L, "orc_polar_" + sfx).restype = ct
            getattr(L, "orc_mat3_det_" + sfx).restype = ct
            getattr(L, "orc_tet_volume_" + sfx).restype = ct
        _lib = L
    return _lib


_ref = None


def ref_lib():
    """The reference's own object code (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None and os.path.exists(_REF_SO):
        L = C.CDLL(_REF_SO)
        for sfx, ct in (("f", C.c_float), ("d", C.c_double)):
            getattr(L, "ref_polar_" + sfx).restype = ct
            getattr(L, "ref_mat3_det_" + sfx).restype = ct
            getattr(L, "ref_tet_volume_" + sfx).restype = ct
        _ref = L
    return _ref


def regular_grid(n, mn, mx):
    """RegularGridTopology positions (float64, N x 3) and hexahedra (H x 8, uint32)."""
    nx, ny, nz = n
    pos = np.empty((nx * ny * nz, 3), np.float64)
    H = max(nx - 1, 0) * max(ny - 1, 0) * max(nz - 1, 0)
    hexas = np.empty((H, 8), np.uint32)
    lib().orc_grid(nx, ny, nz, _ptr(np.asarray(mn, np.float64)), _ptr(np.asarray(mx, np.float64)), _ptr(pos), _ptr(hexas))
    return pos, hexas


def hexas_to_tetras(n, mode):
    """mode 0/1: Hexa2TetraTopologicalMapping swapping=false/true; 2: TetrahedronFEMForceField::init's own."""
    nx, ny, nz = n
    H = (nx - 1) * (ny - 1) * (nz - 1)
    tets = np.empty((H * 6, 4), np.uint32)
    lib().orc_hexas_to_tetras(nx, ny, nz, mode, _ptr(tets))
    return tets


def box_roi(pos, box):
    """BoxROI aligned box (closed intervals), BoxROI.inl:189-201."""
    b = np.asarray(box, np.float64)
    m = np.all((pos >= b[:3]) & (pos <= b[3:]), axis=1)
    return np.nonzero(m)[0].astype(np.uint32)


class OracleScene:
    """One solver node of the reference, restated on the CPU (see oracle/sofa_oracle.hpp)."""

    def __init__(self, dtype, x, v=None):
        self.dtype = np.dtype(dtype)
        self.real = 0 if self.dtype == np.float32 else 1
        self.L = lib()
        self.h = _P(self.L.orc_scene_create(self.real))
        x = np.ascontiguousarray(x, self.dtype)
        self.n = x.shape[0]
        vv = None if v is None else np.ascontiguousarray(v, self.dtype)
        self.L.orc_scene_set_state(self.h, C.c_size_t(self.n), _ptr(x), _ptr(vv))
        self.params = dict(gravity=(0.0, -9.81, 0.0), dt=0.01, rayleighStiffness=0.0, rayleighMass=0.0, vdamping=0.0,
                           firstOrder=0, trapezoidal=0, iterations=25, tolerance=1e-5, threshold=1e-5, warmStart=0,
                           massFirst=1, ffRayleighStiffness=0.0, massRayleighMass=0.0)
        self._push_params()

    def __del__(self):
        try:
            self.L.orc_scene_destroy(self.h)
        except Exception:
            pass

    def _push_params(self):
        p = self.params
        a = np.array(list(p["gravity"]) + [p["dt"], p["rayleighStiffness"], p["rayleighMass"], p["vdamping"], p["firstOrder"],
                                           p["trapezoidal"], p["iterations"], p["tolerance"], p["threshold"], p["warmStart"],
                                           p["massFirst"], p["ffRayleighStiffness"], p["massRayleighMass"]], np.float64)
        self.L.orc_scene_set_params(self.h, _ptr(a))

    def set_params(self, **kw):
        for k in kw:
            assert k in self.params, k
        self.params.update(kw)
        self._push_params()

    def set_tets(self, tets, method="large", young=5000.0, poisson=0.45, local_stiffness_factor=None):
        t = np.ascontiguousarray(tets, np.uint32)
        y = np.atleast_1d(np.asarray(young, np.float64))
        p = np.atleast_1d(np.asarray(poisson, np.float64))
        lsf = None if local_stiffness_factor is None else np.atleast_1d(np.asarray(local_stiffness_factor, np.float64))
        self.L.orc_scene_set_tets(self.h, C.c_size_t(t.shape[0]), _ptr(t), TET_METHODS[method], len(y), _ptr(y), len(p), _ptr(p),
                                  0 if lsf is None else len(lsf), _ptr(lsf))
        self.tets = t

    def set_external_force(self, f):
        """MechanicalObject's externalForce vector (None removes it)."""
        self._ext = None if f is None else np.ascontiguousarray(f, self.dtype)
        self.L.orc_scene_set_external_force(self.h, _ptr(self._ext))

    def set_fast_tets(self, tets, method="qr", young=5000.0, poisson=0.45, edges=None):
        """FastTetrahedralCorotationalForceField (d_method: "polar", "qr"/"large", "polar2", "none"/"linear"/"small")."""
        t = np.ascontiguousarray(tets, np.uint32)
        y = np.atleast_1d(np.asarray(young, np.float64))
        p = np.atleast_1d(np.asarray(poisson, np.float64))
        e = None if edges is None else np.ascontiguousarray(edges, np.uint32)
        self.L.orc_scene_set_fast_tets(self.h, C.c_size_t(t.shape[0]), _ptr(t), FAST_METHODS[method], len(y), _ptr(y), len(p), _ptr(p),
                                       C.c_size_t(0 if e is None else e.shape[0]), _ptr(e))
        self.tets = t

    def set_mesh_mass(self, tets, density=1.0, lumping=False):
        t = np.ascontiguousarray(tets, np.uint32)
        self.L.orc_scene_set_mesh_mass(self.h, C.c_size_t(t.shape[0]), _ptr(t), C.c_double(density), int(lumping))

    def set_plastic(self, max_threshold, yield_threshold=0.0001, creep=0.9):
        self.L.orc_scene_tet_set_plastic(self.h, C.c_double(max_threshold), C.c_double(yield_threshold), C.c_double(creep))

    def set_update_stiffness_matrix(self, on=True):
        self.L.orc_scene_tet_set_update_stiffness(self.h, int(bool(on)))

    def set_tetrahedral_corotational(self, on=True):
        """The tetra force field is the sibling class TetrahedralCorotationalFEMForceField."""
        self.L.orc_scene_tet_set_sibling(self.h, int(bool(on)))

    def tet_reset(self):
        self.L.orc_scene_tet_reset(self.h)

    def set_hexas(self, hexas, method="large", young=5000.0, poisson=0.45):
        hx = np.ascontiguousarray(hexas, np.uint32)
        y = np.atleast_1d(np.asarray(young, np.float64))
        p = np.atleast_1d(np.asarray(poisson, np.float64))
        self.L.orc_scene_set_hexas(self.h, C.c_size_t(hx.shape[0]), _ptr(hx), HEX_METHODS[method], len(y), _ptr(y), len(p), _ptr(p))
        self.hexas = hx

    def set_mass_density(self, density, elems):
        e = np.ascontiguousarray(elems, np.uint32)
        self.L.orc_scene_set_mass(self.h, 0, C.c_double(density), C.c_size_t(e.shape[0]), _ptr(e), e.shape[1], None)

    def set_total_mass(self, total, elems):
        e = np.ascontiguousarray(elems, np.uint32)
        self.L.orc_scene_set_mass(self.h, 1, C.c_double(total), C.c_size_t(e.shape[0]), _ptr(e), e.shape[1], None)

    def set_uniform_mass(self, vertexMass=None, totalMass=None):
        """UniformMass (Data vertexMass | totalMass) instead of DiagonalMass."""
        kind, val = (3, vertexMass) if vertexMass is not None else (4, totalMass)
        self.L.orc_scene_set_mass(self.h, kind, C.c_double(val), C.c_size_t(0), None, 4, None)

    def set_vertex_mass(self, m):
        m = np.ascontiguousarray(m, self.dtype)
        self.L.orc_scene_set_mass(self.h, 2, C.c_double(0), C.c_size_t(0), None, 4, _ptr(m))

    def set_plane(self, prm, rayleighStiffness=0.0):
        """PlaneForceField as the node's last force field; prm = (normal[3], d, stiffness, damping, maxForce, bilateral)."""
        p = np.ascontiguousarray(prm, np.float64)
        self.L.orc_scene_set_plane(self.h, _ptr(p), C.c_double(rayleighStiffness))

    def set_fixed(self, indices, fix_all=False):
        i = np.ascontiguousarray(indices, np.uint32)
        self.L.orc_scene_set_fixed(self.h, C.c_size_t(len(i)), _ptr(i), int(fix_all))

    def set_x(self, x):
        self.L.orc_scene_set_x(self.h, _ptr(np.ascontiguousarray(x, self.dtype)))

    def set_v(self, v):
        self.L.orc_scene_set_v(self.h, _ptr(np.ascontiguousarray(v, self.dtype)))

    def set_dot_double(self, on=True, reverse=False):
        """Test knobs: accumulate CG dot products in double / in reverse order (Scene::dotDouble, dotReverse)."""
        self.L.orc_scene_set_dot_double(self.h, int(bool(on)) | (2 if reverse else 0))

    def set_threads(self, n):
        self.L.orc_scene_set_threads(self.h, int(n))

    def step(self):
        return int(self.L.orc_scene_step(self.h))

    @property
    def end_condition(self):
        return int(self.L.orc_scene_end_condition(self.h))

    def get(self, what, dtype=None):
        cnt = self.L.orc_scene_get(self.h, what.encode(), None)
        out = np.empty(cnt, dtype or self.dtype)
        self.L.orc_scene_get(self.h, what.encode(), _ptr(out))
        if what in ("x", "v", "f", "b", "sol", "x0"):
            return out.reshape(-1, 3)
        if what.endswith("otations") or what.endswith("Transformation") or what in ("fast.linearDfDx", "fast.linearDfDxDiag", "fast.edgeInfo"):
            return out.reshape(-1, 3, 3)
        if what in ("tet.J", "tet.Jsh"):
            return out.reshape(-1, 4, 3)
        if what == "tet.plasticStrains":
            return out.reshape(-1, 6)
        if what == "tet.K":
            return out.reshape(-1, 3)
        if what in ("fast.shapeVectors", "fast.restEdgeVectors"):
            return out.reshape(-1, 3)
        if what == "fast.edges":
            return out.astype(np.int64).reshape(-1, 2)
        if what == "tet.X0":
            return out.reshape(-1, 4, 3)
        if what == "hex.X0":
            return out.reshape(-1, 8, 3)
        if what == "hex.Ke":
            return out.reshape(-1, 24, 24)
        return out

    def tet_matrices(self, e):
        J = np.empty((12, 6), np.float64)
        K = np.empty((6, 6), np.float64)
        self.L.orc_scene_tet_matrices(self.h, C.c_size_t(e), _ptr(J), _ptr(K))
        return J, K

    def fem_add_force(self, f, x):
        f = np.array(f, self.dtype, order="C")
        self.L.orc_scene_fem_add_force(self.h, _ptr(f), _ptr(np.ascontiguousarray(x, self.dtype)))
        return f

    def fem_add_dforce(self, df, dx, k_factor):
        df = np.array(df, self.dtype, order="C")
        self.L.orc_scene_fem_add_dforce(self.h, _ptr(df), _ptr(np.ascontiguousarray(dx, self.dtype)), C.c_double(k_factor))
        return df

    def hex_get_rotations(self):
        """HexahedronFEMForceField::getNodeRotation for every node."""
        out = np.empty((self.n, 3, 3), self.dtype)
        self.L.orc_scene_hex_get_rotations(self.h, _ptr(out))
        return out

    def tet_von_mises(self, x, how=1):
        """computeVonMisesStress at positions x: (per element, per node)."""
        pe = np.zeros(self.tets.shape[0], self.dtype); pn = np.zeros(self.n, self.dtype)
        self.L.orc_scene_tet_von_mises(self.h, _ptr(np.ascontiguousarray(x, self.dtype)), int(how), _ptr(pe), _ptr(pn))
        return pe, pn

    def tet_get_rotations(self):
        """TetrahedronFEMForceField::getRotations(VecReal&): per-node 3x3."""
        out = np.empty((self.n, 3, 3), self.dtype)
        self.L.orc_scene_tet_get_rotations(self.h, _ptr(out))
        return out

    def compute_force(self):
        f = np.empty((self.n, 3), self.dtype)
        self.L.orc_scene_compute_force(self.h, _ptr(f))
        return f

    def apply(self, p, m, b, k):
        q = np.empty((self.n, 3), self.dtype)
        self.L.orc_scene_apply(self.h, _ptr(q), _ptr(np.ascontiguousarray(p, self.dtype)), C.c_double(m), C.c_double(b), C.c_double(k))
        return q

    def cg(self, b, m, bf, k, x0=None):
        x = np.zeros((self.n, 3), self.dtype) if x0 is None else np.array(x0, self.dtype, order="C")
        it = self.L.orc_scene_cg(self.h, _ptr(x), _ptr(np.ascontiguousarray(b, self.dtype)), C.c_double(m), C.c_double(bf), C.c_double(k))
        return x, int(it)

    def graph(self, which="Error"):
        out = np.empty(4096, np.float64)
        n = self.L.orc_scene_graph(self.h, 0 if which == "Error" else 1, _ptr(out), C.c_size_t(4096))
        return out[:n].copy()

    @property
    def hex_potential_energy(self):
        return float(self.L.orc_scene_hex_potential_energy(self.h))


def vop(dtype, r, a, b, k):
    """MechanicalObject::vOp(r, a, b, k) on arrays; pass the same array object to alias operands."""
    real = 0 if np.dtype(dtype) == np.float32 else 1
    lib().orc_vop(real, C.c_size_t(r.shape[0]), _ptr(r), _ptr(a), _ptr(b), C.c_double(k))
    return r


def plane_add_force(dtype, prm, f, x, v):
    """PlaneForceField::addForce on (f, x, v); prm = (normal[3], d, stiffness, damping, maxForce, bilateral).  Returns (f, contacts)."""
    L = lib(); dtype = np.dtype(dtype)
    f = np.ascontiguousarray(f, dtype).copy(); x = np.ascontiguousarray(x, dtype); v = np.ascontiguousarray(v, dtype)
    p = np.ascontiguousarray(prm, np.float64); c = np.zeros(f.shape[0], np.uint8)
    L.orc_plane(0 if dtype == np.float32 else 1, C.c_size_t(f.shape[0]), _ptr(p), _ptr(f), _ptr(x), _ptr(v), _ptr(c), None, C.c_double(0.0), 0)
    return f, c


def plane_add_dforce(dtype, prm, df, dx, contacts, k_factor):
    L = lib(); dtype = np.dtype(dtype)
    df = np.ascontiguousarray(df, dtype).copy(); dx = np.ascontiguousarray(dx, dtype)
    p = np.ascontiguousarray(prm, np.float64); c = np.ascontiguousarray(contacts, np.uint8)
    L.orc_plane(0 if dtype == np.float32 else 1, C.c_size_t(df.shape[0]), _ptr(p), _ptr(df), None, None, _ptr(c), _ptr(dx), C.c_double(k_factor), 1)
    return df


def vdot(dtype, a, b):
    real = 0 if np.dtype(dtype) == np.float32 else 1
    return float(lib().orc_vdot(real, C.c_size_t(a.shape[0]), _ptr(a), _ptr(b)))


class OracleMeshMatrixMass:
    """MeshMatrixMass on tetrahedra in the oracle (oracle/sofa_oracle.hpp: MeshMatrixMass)."""

    def __init__(self, dtype, pos, tets, density=1.0, lumping=False):
        self.L = lib()
        self.dtype = np.dtype(dtype).type
        self.real = 0 if self.dtype == np.float32 else 1
        p = np.ascontiguousarray(pos, self.dtype); t = np.ascontiguousarray(tets, np.uint32)
        self.n = p.shape[0]
        self.h = _P(self.L.orc_meshmass_create(self.real, C.c_size_t(self.n), _ptr(p), C.c_size_t(t.shape[0]), _ptr(t), C.c_double(density), int(lumping)))
        E = self.L.orc_meshmass_n_edges(self.h)
        self.edges = np.zeros((E, 2), np.uint32); self.vertexMass = np.zeros(self.n, self.dtype); self.edgeMass = np.zeros(E, self.dtype)
        self.L.orc_meshmass_arrays(self.h, _ptr(self.edges), _ptr(self.vertexMass), _ptr(self.edgeMass))
        self.totalMass = self.L.orc_meshmass_total(self.h)

    def _op(self, op, res, dx=None, factor=1.0, g=(0, 0, 0)):
        r = np.array(res, self.dtype, order="C")
        d = None if dx is None else np.ascontiguousarray(dx, self.dtype)
        gg = (C.c_double * 3)(*[float(v) for v in g])
        ok = self.L.orc_meshmass_op(self.h, op, C.c_size_t(self.n), _ptr(r), _ptr(d) if d is not None else None, C.c_double(factor), gg)
        return r, ok

    def addMDx(self, res, dx, factor=1.0):
        return self._op(0, res, dx, factor)[0]

    def addForce(self, f, g):
        return self._op(1, f, None, 1.0, g)[0]

    def accFromF(self, f):
        return self._op(2, np.zeros_like(f), f)

    def __del__(self):
        try:
            self.L.orc_meshmass_destroy(self.h)
        except Exception:
            pass
