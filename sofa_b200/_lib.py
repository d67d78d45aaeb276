"""ctypes binding of libsofa_b200.so (the C ABI declared in include/sofa_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no device is present, every
entry point raises.  Build it with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C sofa_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsofa_b200.so")

F32, F64 = 0, 1
_P = C.c_void_p


class Sofab200Error(RuntimeError):
    pass


class TetFemDesc(C.Structure):
    _fields_ = [("method", C.c_int), ("n_young", C.c_size_t), ("young", C.POINTER(C.c_double)), ("n_poisson", C.c_size_t),
                ("poisson", C.POINTER(C.c_double)), ("n_local_stiffness", C.c_size_t), ("local_stiffness", C.POINTER(C.c_double)),
                ("tile_elems", C.c_int), ("shared_nodes", C.POINTER(C.c_ubyte)),
                ("plastic_max_threshold", C.c_double), ("plastic_yield_threshold", C.c_double), ("plastic_creep", C.c_double), ("update_stiffness_matrix", C.c_int), ("tetrahedral_corotational", C.c_int), ("compute_von_mises", C.c_int),
                ("fast_corotational", C.c_int), ("n_edges", C.c_size_t), ("edges", C.POINTER(C.c_uint32))]


class HexFemDesc(C.Structure):
    _fields_ = [("method", C.c_int), ("n_young", C.c_size_t), ("young", C.POINTER(C.c_double)), ("n_poisson", C.c_size_t),
                ("poisson", C.POINTER(C.c_double)), ("tile_elems", C.c_int)]


class PlaneDesc(C.Structure):
    _fields_ = [("normal", C.c_double * 3), ("d", C.c_double), ("stiffness", C.c_double), ("damping", C.c_double), ("max_force", C.c_double),
                ("bilateral", C.c_int)]


class NodeDesc(C.Structure):
    _fields_ = [("tetfem", _P), ("hexfem", _P), ("vertex_mass_host", _P), ("n_fixed", C.c_size_t), ("fixed_host", C.POINTER(C.c_uint32)),
                ("fix_all", C.c_int), ("mass_first", C.c_int), ("uniform_mass", C.c_int), ("uniform_vertex_mass", C.c_double), ("plane", C.POINTER(PlaneDesc)), ("plane_rayleigh_stiffness", C.c_double)]


class HaloDesc(C.Structure):
    _fields_ = [("owned", C.POINTER(C.c_ubyte)), ("n_interface", C.c_size_t), ("interface", C.POINTER(C.c_uint32)), ("my_slot", C.POINTER(C.c_int32)),
                ("max_sharers", C.c_int), ("n_neighbours", C.c_int), ("nb_rank", C.POINTER(C.c_int)), ("nb_count", C.POINTER(C.c_size_t)),
                ("nb_rows", C.POINTER(C.POINTER(C.c_uint32))), ("nb_slot", C.POINTER(C.POINTER(C.c_int32)))]


class PeerDesc(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("peer_base", C.POINTER(C.c_void_p)), ("inbox_rows", C.c_size_t), ("remote_off", C.POINTER(C.c_size_t))]


class SolverParams(C.Structure):
    _fields_ = [("gravity", C.c_double * 3), ("dt", C.c_double), ("rayleigh_stiffness", C.c_double), ("rayleigh_mass", C.c_double),
                ("vdamping", C.c_double), ("first_order", C.c_int), ("trapezoidal", C.c_int), ("iterations", C.c_uint),
                ("tolerance", C.c_double), ("threshold", C.c_double), ("warm_start", C.c_int), ("ff_rayleigh_stiffness", C.c_double),
                ("mass_rayleigh_mass", C.c_double)]


# every symbol include/sofa_b200.h declares, with (restype, argtypes)
_SZ, _I, _D, _U64 = C.c_size_t, C.c_int, C.c_double, C.c_uint64
SYMBOLS = {
    "sofab200_version": (C.c_char_p, []),
    "sofab200_last_error": (C.c_char_p, []),
    "sofab200_ctx_create": (_I, [_I, _P, C.POINTER(_P)]),
    "sofab200_ctx_destroy": (_I, [_P]),
    "sofab200_ctx_set_stream": (_I, [_P, _P]),
    "sofab200_ctx_synchronize": (_I, [_P]),
    "sofab200_ctx_launch_count": (_U64, [_P]),
    "sofab200_ctx_profile_begin": (_I, [_P]),
    "sofab200_ctx_profile_end": (_I, [_P, C.POINTER(_D), C.POINTER(_U64)]),
    "sofab200_ctx_trace_begin": (_I, [_P]),
    "sofab200_ctx_trace_end": (_I, [_P, C.POINTER(_U64), _SZ]),
    "sofab200_mo_vop": (_I, [_P, _I, _SZ, _P, _P, _P, _D]),
    "sofab200_mo_vdot": (_I, [_P, _I, _SZ, _P, _P, C.POINTER(_D)]),
    "sofab200_mo_vdot_dev": (_I, [_P, _I, _SZ, _P, _P, _P, _P]),
    "sofab200_mo_vmultiop_integrate": (_I, [_P, _I, _SZ, _P, _P, _P, _D, _D]),
    "sofab200_mo_accumulate_force": (_I, [_P, _I, _SZ, _P, _P]),
    "sofab200_mass_add_mdx": (_I, [_P, _I, _SZ, _P, _P, _P, _D]),
    "sofab200_plane_add_force": (_I, [_P, _I, _SZ, _P, _P, _P, C.POINTER(PlaneDesc), _P]),
    "sofab200_plane_add_dforce": (_I, [_P, _I, _SZ, _P, _P, C.POINTER(PlaneDesc), _P, _D]),
    "sofab200_uniform_mass_add_mdx": (_I, [_P, _I, _SZ, _P, _P, _D, _D]),
    "sofab200_meshmass_create": (_I, [_P, _I, _SZ, _P, _SZ, _P, _P, _I, _D, C.POINTER(_P)]),
    "sofab200_meshmass_destroy": (_I, [_P]),
    "sofab200_meshmass_add_mdx": (_I, [_P, _P, _P, _D]),
    "sofab200_meshmass_add_force": (_I, [_P, _P, C.POINTER(_D)]),
    "sofab200_meshmass_acc_from_f": (_I, [_P, _P, _P]),
    "sofab200_uniform_mass_add_force": (_I, [_P, _I, _SZ, _P, _D, C.POINTER(_D)]),
    "sofab200_mass_add_force": (_I, [_P, _I, _SZ, _P, _P, C.POINTER(_D)]),
    "sofab200_mass_acc_from_f": (_I, [_P, _I, _SZ, _P, _P, _P]),
    "sofab200_fixed_project_response": (_I, [_P, _I, _SZ, _P, _SZ, _P, _I]),
    "sofab200_tetfem_create": (_I, [_P, _I, _SZ, _P, _SZ, _P, C.POINTER(TetFemDesc), C.POINTER(_P)]),
    "sofab200_tetfem_destroy": (_I, [_P]),
    "sofab200_tetfem_add_force": (_I, [_P, _P, _P]),
    "sofab200_tetfem_add_dforce": (_I, [_P, _P, _P, _D]),
    "sofab200_tetfem_get": (_I, [_P, C.c_char_p, _P]),
    "sofab200_tetfem_stats": (_I, [_P, C.POINTER(_U64)]),
    "sofab200_tetfem_get_rotations": (_I, [_P, _P]),
    "sofab200_tetfem_reset": (_I, [_P]),
    "sofab200_tetfem_compute_von_mises": (_I, [_P, _P, _P, _P]),
    "sofab200_hexfem_create": (_I, [_P, _I, _SZ, _P, _SZ, _P, C.POINTER(HexFemDesc), C.POINTER(_P)]),
    "sofab200_hexfem_destroy": (_I, [_P]),
    "sofab200_hexfem_add_force": (_I, [_P, _P, _P]),
    "sofab200_hexfem_add_dforce": (_I, [_P, _P, _P, _D]),
    "sofab200_hexfem_get": (_I, [_P, C.c_char_p, _P]),
    "sofab200_hexfem_stats": (_I, [_P, C.POINTER(_U64)]),
    "sofab200_hexfem_get_rotations": (_I, [_P, _P]),
    "sofab200_node_create": (_I, [_P, _I, _SZ, C.POINTER(NodeDesc), C.POINTER(_P)]),
    "sofab200_node_destroy": (_I, [_P]),
    "sofab200_node_set_params": (_I, [_P, C.POINTER(SolverParams)]),
    "sofab200_node_compute_force": (_I, [_P, _P, _P]),
    "sofab200_node_apply": (_I, [_P, _P, _P, _D, _D, _D]),
    "sofab200_node_add_mbkdx": (_I, [_P, _P, _P, _P, _D, _D, _D, _I, _D, _I]),
    "sofab200_node_set_vertex_mass": (_I, [_P, _P]),
    "sofab200_node_set_mesh_mass": (_I, [_P, _P]),
    "sofab200_node_cg_solve": (_I, [_P, _P, _P, _D, _D, _D, C.POINTER(_I)]),
    "sofab200_node_step": (_I, [_P, _P, _P]),
    "sofab200_node_step_host": (_I, [_P, _P, _P]),
    "sofab200_node_step_host_x": (_I, [_P, _P, _P, _P]),
    "sofab200_node_set_external_force": (_I, [_P, _P]),
    "sofab200_node_step_pipelined": (_I, [_P, _P, _P, _P, _P]),
    "sofab200_node_flush": (_I, [_P]),
    "sofab200_node_cg_kernel_info": (_I, [_P, C.POINTER(C.c_int)]),
    "sofab200_node_last_solve": (_I, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_D), C.POINTER(_SZ), C.POINTER(_D), C.POINTER(_SZ), _SZ]),
    "sofab200_node_get": (_I, [_P, C.c_char_p, _P]),
    "sofab200_node_reset": (_I, [_P]),
    "sofab200_comm_get_unique_id": (_I, [_P]),
    "sofab200_comm_create": (_I, [_P, _I, _I, _P, C.POINTER(_P)]),
    "sofab200_comm_destroy": (_I, [_P]),
    "sofab200_node_set_distributed": (_I, [_P, _P, C.POINTER(HaloDesc)]),
    "sofab200_peer_alloc": (_I, [_P, _SZ, C.POINTER(_P), C.POINTER(C.c_ubyte)]),
    "sofab200_peer_open": (_I, [_P, C.POINTER(C.c_ubyte), C.POINTER(_P)]),
    "sofab200_peer_close": (_I, [_P, _P]),
    "sofab200_peer_free": (_I, [_P, _P]),
    "sofab200_node_peer_bytes": (_SZ, [_P, _SZ]),
    "sofab200_node_set_peer": (_I, [_P, C.POINTER(PeerDesc)]),
}

_lib = None


def load():
    """Load libsofa_b200.so and bind every symbol of the C ABI.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Sofab200Error(f"{LIB_PATH} is missing: build it with `make -C sofa_b200/csrc` "
                                "(or __graft_entry__.build()).  sofa_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError here == header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise Sofab200Error(f"sofa_b200 error {rc}: {load().sofab200_last_error().decode()}")
