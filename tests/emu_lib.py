"""ctypes binding of tests/emu/libtet_emu.so: the product's tile/gather plan and HD element functions executed
on the CPU (test infrastructure; see tests/emu/tet_emu.cu)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, "emu", "libtet_emu.so")
_SRC = os.path.join(_HERE, "emu", "tet_emu.cu")
_P = C.c_void_p


class EmuEpilogue(C.Structure):
    _fields_ = [("init_src", _P), ("out", _P), ("sign", C.c_int), ("pre_kind", C.c_int), ("post_kind", C.c_int), ("mass", _P),
                ("mdx_src", _P), ("mass_factor", C.c_double), ("gravity", C.c_double * 3), ("has_scale", C.c_int), ("scale", C.c_double),
                ("fixed", _P), ("dot_with", _P)]


def build():
    csrc = os.path.join(_ROOT, "sofa_b200", "csrc")
    deps = [_SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if os.path.exists(_SO) and all(os.path.getmtime(d) <= os.path.getmtime(_SO) for d in deps):
        return
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off",
                           "-shared", "-cudart", "shared", _SRC, "-I", csrc, "-o", _SO])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.emu_tet_create.restype = _P
        L.emu_tet_error.restype = C.c_char_p
        L.emu_tet_run.restype = C.c_double
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


class EmuTet:
    METHODS = {"small": 0, "large": 1, "polar": 2, "svd": 3}

    def __init__(self, dtype, rest, tets, method, young, poisson, tile_e=256, shared_nodes=None):
        self.dtype = np.dtype(dtype)
        self.real = 0 if self.dtype == np.float32 else 1
        rest = np.ascontiguousarray(rest, self.dtype)
        tets = np.ascontiguousarray(tets, np.uint32)
        y = np.atleast_1d(np.asarray(young, np.float64)); p = np.atleast_1d(np.asarray(poisson, np.float64))
        self.n = rest.shape[0]
        shared = None
        if shared_nodes is not None:
            shared = np.zeros(self.n, np.uint8); shared[np.asarray(shared_nodes, np.int64)] = 1
        self.h = _P(lib().emu_tet_create(self.real, C.c_size_t(self.n), _ptr(rest), C.c_size_t(tets.shape[0]), _ptr(tets), self.METHODS[method],
                                         C.c_size_t(len(y)), _ptr(y), C.c_size_t(len(p)), _ptr(p), int(tile_e), _ptr(shared) if shared is not None else None))
        err = lib().emu_tet_error(self.h).decode()
        if err:
            raise RuntimeError(err)
        self.n_tets = tets.shape[0]

    def __del__(self):
        try:
            lib().emu_tet_destroy(self.h)
        except Exception:
            pass

    def stats(self):
        out = (C.c_uint64 * 8)()
        lib().emu_tet_stats(self.h, out)
        return dict(zip(["tiles", "tile_e", "interior", "shared", "staged", "smem", "maxval", "n_tets"], list(out)))

    def run(self, dforce, vec, kf=0.0, init=None, sign=None, pre_kind=0, post_kind=0, mass=None, mass_factor=0.0, gravity=(0, 0, 0),
            scale=None, fixed=None, dot=False):
        vec = np.ascontiguousarray(vec, self.dtype)
        out = np.full((self.n, 3), 999.0, self.dtype)
        q = EmuEpilogue()
        keep = [vec, out]
        if init is not None:
            init = np.ascontiguousarray(init, self.dtype); keep.append(init); q.init_src = _ptr(init)
        q.out = _ptr(out)
        q.sign = sign if sign is not None else (-1 if dforce else +1)
        q.pre_kind, q.post_kind = pre_kind, post_kind
        if mass is not None:
            mass = np.ascontiguousarray(mass, self.dtype); keep.append(mass); q.mass = _ptr(mass)
        q.mdx_src = _ptr(vec)
        q.mass_factor = mass_factor
        q.gravity = (C.c_double * 3)(*gravity)
        q.has_scale = int(scale is not None); q.scale = 1.0 if scale is None else scale
        if fixed is not None:
            fixed = np.ascontiguousarray(fixed, np.uint8); keep.append(fixed); q.fixed = _ptr(fixed)
        if dot:
            q.dot_with = _ptr(vec)
        d = lib().emu_tet_run(self.h, int(dforce), _ptr(vec), C.c_double(kf), C.byref(q))
        return (out, d) if dot else out

    def rotations(self):
        out = np.zeros((self.n_tets, 3, 3), self.dtype)
        lib().emu_tet_rotations(self.h, _ptr(out))
        return out
