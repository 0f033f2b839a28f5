"""ctypes binding of libsom_b200.so (the C ABI declared in include/som_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` /
``python -m xpysom_dask_b200.build``.  There is no fallback: if the shared
object is missing, or a call fails, this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsom_b200.so")

c_f32p = ctypes.c_void_p
c_i32p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/som_b200.h exactly
SIGNATURES = {
    "som_b200_abi_version": (ctypes.c_int, []),
    "som_b200_last_error": (ctypes.c_char_p, []),
    "som_b200_device_info": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                            ctypes.POINTER(ctypes.c_size_t)]),
    "som_b200_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "som_b200_shard_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "som_b200_neigh_table_floats": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "som_b200_neigh_scratch_floats": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "som_b200_prepare_codebook": (ctypes.c_int, [c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                 ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_pick_algo": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "som_b200_prepare_samples": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, c_f32p,
                                                ctypes.c_void_p]),
    "som_b200_accum_scales": (ctypes.c_int, [c_f32p, ctypes.c_int, ctypes.c_double, c_f32p, c_f32p, ctypes.c_void_p]),
    "som_b200_accum_words": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "som_b200_accum_replicas": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "som_b200_accum_fold_replicas": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "som_b200_filter_eligible": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "som_b200_filter_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "som_b200_filter_overflow_offset": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "som_b200_filter_prepare_samples": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                                       ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_bmu_filter": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, ctypes.c_int, c_i32p,
                                           c_f32p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_accum_finalize": (ctypes.c_int, [ctypes.c_void_p, c_f32p, ctypes.c_int, ctypes.c_int, c_f32p, c_f32p,
                                               ctypes.c_void_p]),
    "som_b200_accum_fold": (ctypes.c_int, [ctypes.c_void_p, c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                           ctypes.c_void_p]),
    "som_b200_accum_finalize_f64": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_f32p, c_f32p,
                                                   ctypes.c_void_p]),
    "som_b200_bmu": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, c_f32p, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_float, ctypes.c_int, c_i32p, c_f32p,
                                    ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_distances": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_float, ctypes.c_int, c_f32p, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_top2": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, ctypes.c_int, c_i32p,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_accumulate": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_i32p, ctypes.c_int,
                                           c_f32p, ctypes.c_void_p, ctypes.c_void_p]),
    "som_b200_epoch_accumulate": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, c_f32p,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_int,
                                                 c_f32p, ctypes.c_void_p, c_i32p, ctypes.c_void_p, ctypes.c_size_t,
                                                 ctypes.c_void_p]),
    "som_b200_neigh_apply": (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_int, c_f32p, c_f32p, c_f32p, ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_neigh_apply_sched": (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double,
                                                  ctypes.c_int, c_f32p, c_f32p, c_f32p, ctypes.c_size_t,
                                                  ctypes.c_void_p]),
    "som_b200_epoch_advance": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "som_b200_merge": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]),
    "som_b200_epoch_tail": (ctypes.c_int, [ctypes.c_void_p, c_f32p, c_f32p, c_f32p, c_f32p, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_float, c_f32p, c_f32p, c_f32p, ctypes.c_size_t,
                                           ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "som_b200_peer_create": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p),
                                            ctypes.c_void_p]),
    "som_b200_peer_connect": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "som_b200_peer_accumulator": (ctypes.c_void_p, [ctypes.c_void_p]),
    "som_b200_peer_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "som_b200_quantize": (ctypes.c_int, [c_f32p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, c_f32p, ctypes.c_int,
                                         c_i32p, c_f32p, c_f32p, ctypes.c_void_p]),
    "som_b200_distance_map": (ctypes.c_int, [c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                             c_f32p, ctypes.c_void_p]),
    "som_b200_debug_timeline": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "som_b200_train_host": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.POINTER(ctypes.c_double),
                                           ctypes.POINTER(ctypes.c_double), ctypes.c_int]),
}


class TrainConfig(ctypes.Structure):
    """struct som_b200_train_config"""
    _fields_ = [("gx", ctypes.c_int), ("gy", ctypes.c_int), ("d", ctypes.c_int),
                ("topology", ctypes.c_int), ("neigh_kind", ctypes.c_int), ("dist_kind", ctypes.c_int),
                ("algo", ctypes.c_int), ("compact_support", ctypes.c_int),
                ("p", ctypes.c_float), ("std_coeff", ctypes.c_double)]


# enum values of include/som_b200.h
DIST = {"euclidean": 0, "euclidean_no_opt": 0, "cosine": 1, "manhattan": 2, "manhattan_no_opt": 2,
        "chebyshev": 3, "norm_p": 4, "norm_p_no_opt": 4}
NEIGH = {"gaussian": 0, "mexican_hat": 1, "bubble": 2, "triangle": 3}
TOPO = {"rectangular": 0, "hexagonal": 1}
ALGO = {"auto": 0, "simt": 1, "tc": 2, "tc16": 3}

ABI_VERSION = 2
_lib = None


class SomB200Error(RuntimeError):
    pass


def load():
    """Load the shared object once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SomB200Error(
            "libsom_b200.so not found at %s — build it with `python -m xpysom_dask_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)           # AttributeError if the ABI lost a symbol
        fn.restype = res
        fn.argtypes = args
    if lib.som_b200_abi_version() != ABI_VERSION:
        raise SomB200Error("libsom_b200.so ABI version %d, expected %d" % (lib.som_b200_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().som_b200_last_error()
        raise SomB200Error("%s failed (rc=%d): %s" % (what, rc, (msg or b"").decode("utf-8", "replace")))
