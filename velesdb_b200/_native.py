"""ctypes binding of libveles_b200.so (include/veles_b200.h).

The library is the product; this module only loads it.  There is no fallback: if the shared
object is missing or no CUDA device is present, calls raise ``VelesError``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libveles_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_IO, ERR_OOM, ERR_OVERFLOW, ERR_UNSUPPORTED = -1, -2, -3, -4, -5, -6
COSINE, EUCLIDEAN, DOT, HAMMING, JACCARD = 0, 1, 2, 3, 4
F32, F16, BIN1 = 0, 1, 2
FAST, BALANCED, ACCURATE, PERFECT, CUSTOM = 0, 1, 2, 3, 4
INVALID_ID = 0xFFFFFFFF

# every symbol include/veles_b200.h declares: (name, restype, argtypes)
_vp, _u32, _u64, _i32, _f = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32, C.c_float
SYMBOLS = [
    ("veles_init", _i32, [_i32]),
    ("veles_shutdown", _i32, []),
    ("veles_last_error", C.c_char_p, []),
    ("veles_version", C.c_char_p, []),
    ("veles_launch_count", _u64, []),
    ("veles_ef_search", _u64, [_i32, _u64, _u64]),
    ("veles_transform_score", _f, [_i32, _f]),
    ("veles_index_from_reference_files", _i32, [C.c_char_p, C.c_char_p, _i32, _i32, C.POINTER(_vp)]),
    ("veles_index_from_arrays", _i32,
     [_vp, _u64, _u32, _i32, _i32, _i32, _u32, C.POINTER(_vp), C.POINTER(_vp), _vp, _u32, _u32, _u64, _u32,
      C.POINTER(_vp)]),
    ("veles_index_from_vectors", _i32, [_vp, _u64, _u32, _i32, _i32, _i32, C.POINTER(_vp)]),
    ("veles_index_create", _i32, [_u64, _u32, _i32, _i32, C.POINTER(_vp)]),
    ("veles_index_set_rows_d", _i32, [_vp, _u64, _u64, _vp, _i32, _vp]),
    ("veles_index_get_rows", _i32, [_vp, _u64, _u64, _vp]),
    ("veles_index_free", _i32, [_vp]),
    ("veles_index_len", _u64, [_vp]),
    ("veles_index_dim", _u32, [_vp]),
    ("veles_index_metric", _i32, [_vp]),
    ("veles_index_max_layer", _u32, [_vp]),
    ("veles_index_entry_point", _u64, [_vp]),
    ("veles_index_device_bytes", _u64, [_vp]),
    ("veles_index_dump", _i32, [_vp, C.c_char_p, C.c_char_p]),
    ("veles_index_dump_graph", _i32, [_vp, C.c_char_p, C.c_char_p]),
    ("veles_index_export_layer", _i32, [_vp, _u32, C.POINTER(_u64), C.POINTER(_u64), _vp, _vp]),
    ("veles_search_batch", _i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    ("veles_search_batch_d", _i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    ("veles_search_status", _i32, [_vp, _vp]),
    ("veles_search_submit", _i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, C.POINTER(_u64)]),
    ("veles_search_wait", _i32, [_vp, _u64]),
    ("veles_bruteforce_batch", _i32, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    ("veles_bruteforce_batch_d", _i32, [_vp, _vp, _u32, _u32, _vp, _vp, _vp]),
    ("veles_bruteforce_batch_relaxed", _i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp]),
    ("veles_bruteforce_batch_relaxed_d", _i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, C.POINTER(_f), _vp]),
    ("veles_rerank_batch", _i32, [_vp, _vp, _u32, _vp, _u32, _vp, _vp]),
    ("veles_distance_pairs", _i32, [_i32, _vp, _vp, _u32, _u32, _i32, _vp, _vp]),
    ("veles_bm25_from_csr", _i32, [_u32, _vp, _vp, _vp, _vp, _u32, _vp, _u64, _u64, _f, _f, C.POINTER(_vp)]),
    ("veles_bm25_free", _i32, [_vp]),
    ("veles_bm25_search_batch", _i32, [_vp, _vp, _vp, _u32, _u32, _vp, _vp, _vp, _vp]),
    ("veles_rrf_hybrid", _i32, [_vp, _vp, _vp, _vp, _u32, _u32, _f, _u32, _vp, _vp, _vp, _vp]),
    ("veles_hybrid_search_batch", _i32, [_vp, _vp, _vp, _vp, _vp, _u32, _u32, _u32, _f, _vp, _vp, _vp, _vp]),
    ("veles_fuse", _i32, [_i32, _vp, _u32, _vp, _vp, _u32, _f, _f, _f, _u32, _vp, _vp, _vp, _vp]),
    ("veles_index_build_graph", _i32, [_vp, _u32, _u32, _vp]),
    ("veles_index_build_graph_exact", _i32, [_vp, _u32, _u32, _vp]),
    ("veles_index_append", _i32, [_vp, _vp, _u64, _i32, _u32, _vp]),
    ("veles_index_attach_sq8", _i32, [_vp, _u64, _vp]),
    ("veles_index_has_sq8", _i32, [_vp]),
    ("veles_index_sq8_export", _i32, [_vp, _vp, _vp, _vp, _vp]),
    ("veles_search_batch_sq8", _i32, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    ("veles_search_batch_sq8_d", _i32, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
    ("veles_search_batch_multi_entry", _i32, [_vp, _vp, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("veles_index_set_id_map", _i32, [_vp, _vp, _vp]),
    ("veles_comm_handle_bytes", _i32, []),
    ("veles_comm_create", _i32, [_i32, _i32, _u32, _u32, C.POINTER(_vp), _vp]),
    ("veles_comm_connect", _i32, [_vp, _vp]),
    ("veles_search_batch_gather_d", _i32, [_vp, _vp, _vp, _u32, _u32, _u32, _vp, _vp]),
    ("veles_comm_window", _i32, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    ("veles_comm_status", _i32, [_vp, _vp]),
    ("veles_comm_destroy", _i32, [_vp]),
    ("veles_search_batch_mapped", _i32, [_vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp]),
]


class VelesError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"veles status {status}: {message}")
        self.status = status


_lib = None


def lib():
    """Loads libveles_b200.so (once).  Raises VelesError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VelesError(ERR_UNSUPPORTED,
                             f"{LIB_PATH} is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != OK:
        raise VelesError(status, lib().veles_last_error().decode("utf-8", "replace"))


_initialised = {}


def init(device=None):
    """Selects the CUDA device for this thread (veles_init).  ``init()`` with no argument only makes sure
    *some* device has been initialised and never switches away from the one chosen earlier -- the
    multi-GPU runtime calls ``init(local_rank)`` once and every later implicit ``init()`` must keep it."""
    if device is None:
        if _initialised:
            return
        device = 0
    if not _initialised.get(device):
        check(lib().veles_init(device))
        _initialised[device] = True


def ptr(a):
    """host numpy array or torch tensor (host or device) -> void*"""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)
