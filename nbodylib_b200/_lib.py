"""Loader + ctypes prototypes for libnbk.so (the C ABI of include/nbk.h).

The library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is no fallback:
if the shared object is missing this module raises, and every entry point needs a CUDA device.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnbk.so")


class NbkParticles(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("pos_stride", C.c_int64),
                ("vel", C.c_void_p), ("vel_stride", C.c_int64),
                ("mass", C.c_void_p), ("mass_stride", C.c_int64),
                ("real_bytes", C.c_int32), ("on_device", C.c_int32)]


class NbkInfo(C.Structure):
    _fields_ = [("n", C.c_int64),
                ("bucket", C.c_int32), ("treetype", C.c_int32), ("kerntype", C.c_int32), ("kernres", C.c_int32), ("nd", C.c_int32),
                ("num_nodes", C.c_int32), ("num_leaves", C.c_int32), ("depth", C.c_int32),
                ("store_bytes", C.c_int32), ("periodic", C.c_int32),
                ("inexact_coords", C.c_int64),
                ("kernnorm", C.c_double), ("period", C.c_double * 3),
                ("build_ms", C.c_double), ("h2d_ms", C.c_double), ("last_kernel_ms", C.c_double), ("last_call_ms", C.c_double),
                ("last_launches", C.c_int64), ("device_bytes", C.c_int64), ("last_flagged", C.c_int64),
                ("warp_aligned", C.c_int32), ("reserved", C.c_int32)]


class NbkFofLists(C.Structure):
    _fields_ = [("head", C.c_void_p), ("next", C.c_void_p), ("tail", C.c_void_p), ("len", C.c_void_p)]


# flags (include/nbk.h)
DEVICE_PTRS, TREE_ORDER, STRICT_PERIODIC, KNN_TREE_FORM, STORE_F64, STORE_F32, OUT_IDS, WARP_ALIGNED = (1 << i for i in range(8))

EXPORTS = [
    "nbk_last_error", "nbk_device_count", "nbk_create", "nbk_destroy", "nbk_get_info", "nbk_get_order",
    "nbk_get_kernel_table", "nbk_get_nodes", "nbk_knn_particles", "nbk_knn_points", "nbk_ball_particles",
    "nbk_ball_points", "nbk_calc_density", "nbk_calc_density_subset", "nbk_calc_veldensity", "nbk_smoothing_scale", "nbk_fof",
    "nbk_fof_criterion", "nbk_fof_criterion_basis", "nbk_attach_halo", "nbk_device_arrays", "nbk_release_cached_memory",
    "nbk_search_criterion_particles", "nbk_search_criterion_points", "nbk_calc_density_particles",
    "nbk_calc_veldensity_particles", "nbk_calc_density_points", "nbk_calc_veldensity_points",
    "nbk_knn_filtered_particles", "nbk_knn_filtered_points", "nbk_calc_smooth_vel", "nbk_calc_smooth_veldisp",
    "nbk_set_option", "nbk_fof_roots", "nbk_union_pairs", "nbk_calc_smooth_velskew", "nbk_calc_smooth_velkurtosis",
    "nbk_knn_phase_particles", "nbk_knn_phase_points",
]

SHARDED_PATH = os.path.join(HERE, "libnbk_sharded.so")
SHARDED_EXPORTS = ["nbk_comm_unique_id", "nbk_comm_init_rank", "nbk_comm_destroy", "nbk_sharded_create", "nbk_sharded_destroy",
                   "nbk_sharded_calc_density", "nbk_sharded_fof", "nbk_sharded_get_info", "nbk_sharded_release"]


class NbkShardedInfo(C.Structure):
    _fields_ = [("n_local", C.c_int64), ("n_global", C.c_int64), ("first_global_id", C.c_int64),
                ("rank", C.c_int32), ("nranks", C.c_int32), ("h_knn", C.c_double),
                ("ghosts_knn", C.c_int64), ("ghosts_fof", C.c_int64), ("density_setups", C.c_int64), ("fof_setups", C.c_int64),
                ("last_kernel_ms", C.c_double), ("last_call_ms", C.c_double), ("last_launches", C.c_int64), ("last_flagged", C.c_int64)]


_lib = None
_sharded = None


def load_sharded():
    """libnbk_sharded.so (include/nbk_sharded.h).  torch is imported first so that the process holds ONE NCCL (the loader
    resolves the library's libnccl.so.2 to the copy torch already mapped)."""
    global _sharded
    if _sharded is not None:
        return _sharded
    load()
    if not os.path.exists(SHARDED_PATH):
        raise RuntimeError("nbodylib_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`" % SHARDED_PATH)
    import torch  # noqa: F401
    S = C.CDLL(SHARDED_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    S.nbk_comm_unique_id.argtypes = [vp]
    S.nbk_comm_init_rank.argtypes = [i32, i32, vp, i32, C.POINTER(vp)]
    S.nbk_comm_destroy.argtypes = [vp]
    S.nbk_sharded_create.argtypes = [vp, C.POINTER(NbkParticles), i64, vp, vp, i32, i32, dbl, C.POINTER(vp)]
    S.nbk_sharded_destroy.argtypes = [vp]
    S.nbk_sharded_calc_density.argtypes = [vp, i32, vp, i32]
    S.nbk_sharded_fof.argtypes = [vp, i32, dbl, vp, i32, i32, vp, C.POINTER(i64), i32]
    S.nbk_sharded_get_info.argtypes = [vp, C.POINTER(NbkShardedInfo)]
    S.nbk_sharded_release.argtypes = [vp]
    for name in SHARDED_EXPORTS:
        getattr(S, name).restype = i32
    _sharded = S
    return S


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "nbodylib_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.nbk_last_error.restype = C.c_char_p
    L.nbk_device_count.restype = i32
    L.nbk_create.argtypes = [C.POINTER(NbkParticles), i64, i32, i32, i32, i32, i32, vp, i32, i32, C.POINTER(vp)]
    L.nbk_destroy.argtypes = [vp]
    L.nbk_get_info.argtypes = [vp, C.POINTER(NbkInfo)]
    L.nbk_get_order.argtypes = [vp, vp, i32]
    L.nbk_get_kernel_table.argtypes = [vp, vp]
    L.nbk_get_nodes.argtypes = [vp, C.POINTER(i64), vp, vp, vp, vp]
    L.nbk_knn_particles.argtypes = [vp, i32, i64, i64, vp, vp, i32]
    L.nbk_knn_points.argtypes = [vp, i32, i64, vp, vp, vp, i32]
    L.nbk_knn_phase_particles.argtypes = [vp, i32, i64, i64, vp, vp, i32]
    L.nbk_knn_phase_points.argtypes = [vp, i32, i64, vp, vp, vp, vp, i32]
    L.nbk_knn_filtered_particles.argtypes = [vp, i32, i64, i64, i32, vp, vp, vp, vp, i32]
    L.nbk_knn_filtered_points.argtypes = [vp, i32, i64, vp, vp, i32, vp, vp, vp, vp, i32]
    L.nbk_calc_smooth_vel.argtypes = [vp, i32, vp, vp, i32]
    L.nbk_calc_smooth_veldisp.argtypes = [vp, i32, vp, vp, vp, i32]
    L.nbk_calc_smooth_velskew.argtypes = [vp, i32, vp, vp, vp, vp, i32]
    L.nbk_calc_smooth_velkurtosis.argtypes = [vp, i32, vp, vp, vp, vp, i32]
    L.nbk_ball_particles.argtypes = [vp, dbl, i64, vp, vp, vp, vp, i64, C.POINTER(i64), i32]
    L.nbk_ball_points.argtypes = [vp, dbl, i64, vp, vp, vp, vp, i64, C.POINTER(i64), i32]
    L.nbk_search_criterion_particles.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp, i64, C.POINTER(i64), i32]
    L.nbk_search_criterion_points.argtypes = [vp, i32, vp, i64, vp, vp, vp, vp, vp, i64, C.POINTER(i64), i32]
    L.nbk_calc_density_particles.argtypes = [vp, i32, i64, vp, vp, i32]
    L.nbk_calc_veldensity_particles.argtypes = [vp, i32, i32, i64, vp, vp, i32]
    L.nbk_calc_density_points.argtypes = [vp, i32, i64, vp, vp, i32]
    L.nbk_calc_veldensity_points.argtypes = [vp, i32, i32, i64, vp, vp, vp, i32]
    L.nbk_calc_density.argtypes = [vp, i32, vp, vp, i32]
    L.nbk_calc_density_subset.argtypes = [vp, i32, vp, vp, vp, i32]
    L.nbk_calc_veldensity.argtypes = [vp, i32, i32, vp, i32]
    L.nbk_smoothing_scale.argtypes = [vp, i32, vp, i32]
    L.nbk_fof.argtypes = [vp, dbl, i32, i32, vp, vp, C.POINTER(i64), C.POINTER(NbkFofLists), i32]
    L.nbk_fof_criterion.argtypes = [vp, i32, vp, i32, i32, vp, vp, C.POINTER(i64), C.POINTER(NbkFofLists), i32]
    L.nbk_fof_criterion_basis.argtypes = [vp, i32, vp, i32, i32, vp, vp, C.POINTER(i64), C.POINTER(NbkFofLists), i32]
    L.nbk_attach_halo.argtypes = [vp, vp]
    L.nbk_release_cached_memory.argtypes = [i32]
    L.nbk_set_option.argtypes = [C.c_char_p, i64]
    L.nbk_fof_roots.argtypes = [vp, i32, dbl, vp, vp, vp, i32]
    L.nbk_union_pairs.argtypes = [i32, i64, i64, vp, vp, vp]
    L.nbk_device_arrays.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    for name in EXPORTS:
        if name not in ("nbk_last_error", "nbk_device_count"):
            getattr(L, name).restype = i32
    _lib = L
    return L


def set_option(name, value):
    """nbk_set_option: process-wide tuning overrides (include/nbk.h lists the names)"""
    check(load().nbk_set_option(name.encode(), int(value)))


class NbkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("nbk error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise NbkError(rc, load().nbk_last_error().decode())
