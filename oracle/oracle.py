"""ctypes binding of the CPU oracle (oracle/veles_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- see the header of veles_oracle.cpp.  Importable from
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs, nowhere else.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libveles_oracle.so")

COSINE, EUCLIDEAN, DOT, HAMMING, JACCARD = 0, 1, 2, 3, 4
FAST, BALANCED, ACCURATE, PERFECT, CUSTOM = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "veles_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def _load():
    build()
    try:
        return C.CDLL(_SO)
    except OSError:
        build(force=True)
        return C.CDLL(_SO)


_lib = _load()

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def _sig(name, res, args):
    f = getattr(_lib, name)
    f.restype = res
    f.argtypes = args
    return f


_sig("vo_force_scalar", None, [C.c_int])
_sig("vo_have_avx2", C.c_int, [])
_sig("vo_graph_distance", C.c_float, [C.c_int, _f32p, _f32p, C.c_uint64, C.c_int])
_sig("vo_metric_value", C.c_float, [C.c_int, _f32p, _f32p, C.c_uint64, C.c_int])
_sig("vo_dot", C.c_float, [_f32p, _f32p, C.c_uint64, C.c_int])
_sig("vo_l2sq", C.c_float, [_f32p, _f32p, C.c_uint64, C.c_int])
_sig("vo_norm_sq", C.c_float, [_f32p, C.c_uint64, C.c_int])
_sig("vo_hamming_binary", C.c_uint32, [_u64p, _u64p, C.c_uint64])
_sig("vo_transform_score", C.c_float, [C.c_int, C.c_float])
_sig("vo_higher_is_better", C.c_int, [C.c_int])
_sig("vo_ef_search", C.c_uint64, [C.c_int, C.c_uint64, C.c_uint64])
_sig("vo_hnsw_new", C.c_void_p, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_int])
_sig("vo_hnsw_free", None, [C.c_void_p])
_sig("vo_hnsw_insert", C.c_uint64, [C.c_void_p, _f32p])
_sig("vo_hnsw_insert_many", None, [C.c_void_p, _f32p, C.c_uint64])
_sig("vo_hnsw_len", C.c_uint64, [C.c_void_p])
_sig("vo_hnsw_dim", C.c_uint32, [C.c_void_p])
_sig("vo_hnsw_num_layers", C.c_uint32, [C.c_void_p])
_sig("vo_hnsw_max_layer", C.c_uint32, [C.c_void_p])
_sig("vo_hnsw_M", C.c_uint32, [C.c_void_p])
_sig("vo_hnsw_M0", C.c_uint32, [C.c_void_p])
_sig("vo_hnsw_has_entry", C.c_int, [C.c_void_p])
_sig("vo_hnsw_entry_point", C.c_uint64, [C.c_void_p])
_sig("vo_hnsw_layer_nodes", C.c_uint64, [C.c_void_p, C.c_uint32])
_sig("vo_hnsw_layer_edges", C.c_uint64, [C.c_void_p, C.c_uint32])
_sig("vo_hnsw_export_layer", None, [C.c_void_p, C.c_uint32, _u64p, _u32p])
_sig("vo_hnsw_vectors", C.POINTER(C.c_float), [C.c_void_p])
_sig("vo_levels", None, [C.c_uint32, C.c_uint64, _u8p])
_sig("vo_hnsw_from_arrays", C.c_void_p,
     [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _f32p, C.c_uint64, C.c_uint32,
      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _u64p, C.c_uint64, C.c_uint32])
_sig("vo_hnsw_search", C.c_uint32, [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_int, _u64p, _f32p, _u64p])
_sig("vo_hnsw_search_multi_entry", C.c_uint32,
     [C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _u64p, _f32p, _u64p, _u64p])
_sig("vo_hnsw_rng_state", C.c_uint64, [C.c_void_p])
_sig("vo_hnsw_set_rng_state", None, [C.c_void_p, C.c_uint64])
_sig("vo_hnsw_search_batch", None,
     [C.c_void_p, _f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u64p, _f32p, _u32p, _u64p])
_sig("vo_hnsw_search_layer", C.c_uint32, [C.c_void_p, _f32p, C.c_uint64, C.c_uint32, C.c_uint32, _u64p, _f32p])
_sig("vo_hnsw_select_neighbors", C.c_uint32, [C.c_void_p, _u64p, _f32p, C.c_uint32, C.c_uint32, _u32p])
_sig("vo_bruteforce", C.c_uint32,
     [C.c_int, _f32p, C.c_uint64, C.c_uint32, _f32p, C.c_uint32, C.c_int, _u64p, _f32p])
_sig("vo_bruteforce_batch", None,
     [C.c_int, _f32p, C.c_uint64, C.c_uint32, _f32p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, _u64p, _f32p])
_sig("vo_bruteforce_binary", C.c_uint32, [_u64p, C.c_uint64, C.c_uint32, _u64p, C.c_uint32, _u64p, _u32p])
_sig("vo_hnsw_dump", C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p])
_sig("vo_hnsw_load", C.c_void_p, [C.c_char_p, C.c_char_p, C.c_int, C.c_int])
_sig("vo_hnsw_frozen", C.c_void_p,
     [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint32,
      C.c_void_p, C.c_void_p, _u64p, C.c_uint64, C.c_uint32])
_sig("vo_hnsw_open", C.c_void_p, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.c_uint32])
_sig("vo_bm25_new", C.c_void_p, [C.c_float, C.c_float])
_sig("vo_bm25_free", None, [C.c_void_p])
_sig("vo_bm25_add", None, [C.c_void_p, C.c_uint64, _u32p, C.c_uint64])
_sig("vo_bm25_remove", C.c_int, [C.c_void_p, C.c_uint64])
_sig("vo_bm25_len", C.c_uint64, [C.c_void_p])
_sig("vo_bm25_term_count", C.c_uint64, [C.c_void_p])
_sig("vo_bm25_search", C.c_uint32, [C.c_void_p, _u32p, C.c_uint64, C.c_uint32, _u64p, _f32p])
_sig("vo_bm25_search_batch", None,
     [C.c_void_p, _u32p, _u32p, C.c_uint64, C.c_uint32, C.c_int, _u64p, _f32p, _u32p])
_sig("vo_rrf_hybrid", C.c_uint32, [_u64p, C.c_uint32, _u64p, C.c_uint32, C.c_float, C.c_uint32, _u64p, _f32p])
_sig("vo_fuse", C.c_uint32,
     [C.c_int, _u32p, C.c_uint32, _u64p, _f32p, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_uint32, _u64p,
      _f32p])

_sig("vo_sq8_train", C.c_void_p, [_f32p, C.c_uint64, C.c_uint32])
_sig("vo_sq8_free", None, [C.c_void_p])
_sig("vo_sq8_params", None, [C.c_void_p, _f32p, _f32p, _f32p])
_sig("vo_sq8_quantize", None, [C.c_void_p, _f32p, C.c_uint64, _u8p])
_sig("vo_sq8_push", None, [C.c_void_p, _f32p, C.c_uint64])
_sig("vo_sq8_len", C.c_uint64, [C.c_void_p])
_sig("vo_sq8_codes", C.POINTER(C.c_uint8), [C.c_void_p])
_sig("vo_sq8_distance_quantized", C.c_uint32, [_u8p, _u8p, C.c_uint32])
_sig("vo_sq8_distance_asymmetric", C.c_float, [C.c_void_p, _f32p, _u8p])
_sig("vo_dual_search_int8", C.c_uint32,
     [C.c_void_p, C.c_void_p, _f32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _u64p, _f32p, _u64p])
_sig("vo_dual_search_int8_batch", None,
     [C.c_void_p, C.c_void_p, _f32p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int, _u64p, _f32p,
      _u32p, _u64p])

def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def force_scalar(on: bool) -> None:
    _lib.vo_force_scalar(1 if on else 0)


def have_avx2() -> bool:
    return bool(_lib.vo_have_avx2())


def graph_distance(metric, a, b, fma=True) -> float:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    return float(_lib.vo_graph_distance(metric, a, b, a.size, int(fma)))


def metric_value(metric, a, b, fma=True) -> float:
    a, b = _f32(a), _f32(b)
    assert a.shape == b.shape
    return float(_lib.vo_metric_value(metric, a, b, a.size, int(fma)))


def dot(a, b, fma=True) -> float:
    a, b = _f32(a), _f32(b)
    return float(_lib.vo_dot(a, b, a.size, int(fma)))


def l2sq(a, b, fma=True) -> float:
    a, b = _f32(a), _f32(b)
    return float(_lib.vo_l2sq(a, b, a.size, int(fma)))


def norm_sq(a, fma=True) -> float:
    a = _f32(a)
    return float(_lib.vo_norm_sq(a, a.size, int(fma)))


def hamming_binary(a, b) -> int:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    return int(_lib.vo_hamming_binary(a, b, a.size))


def transform_score(metric, raw) -> float:
    return float(_lib.vo_transform_score(metric, float(raw)))


def higher_is_better(metric) -> bool:
    return bool(_lib.vo_higher_is_better(metric))


def ef_search(quality, k, custom_ef=0) -> int:
    return int(_lib.vo_ef_search(quality, k, custom_ef))


def levels(M: int, count: int) -> np.ndarray:
    out = np.zeros(count, dtype=np.uint8)
    _lib.vo_levels(M, count, out)
    return out


class Hnsw:
    """NativeHnsw<SimdDistance> restated (graph.rs).  Node ids are insertion indices."""

    def __init__(self, metric, dim, M=32, ef_construction=400, alpha=1.0, fma=True, _handle=None):
        self.metric, self.fma = metric, fma
        self._keep = None
        self._h = _handle if _handle is not None else _lib.vo_hnsw_new(metric, dim, M, ef_construction, alpha, int(fma))
        if not self._h:
            raise RuntimeError("oracle: could not create index")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vo_hnsw_free(self._h)
            self._h = None

    @property
    def dim(self):
        return int(_lib.vo_hnsw_dim(self._h))

    def __len__(self):
        return int(_lib.vo_hnsw_len(self._h))

    def insert(self, v) -> int:
        v = _f32(v)
        assert v.size == self.dim
        return int(_lib.vo_hnsw_insert(self._h, v))

    def insert_many(self, vs) -> None:
        vs = _f32(vs)
        assert vs.ndim == 2 and vs.shape[1] == self.dim
        _lib.vo_hnsw_insert_many(self._h, vs, vs.shape[0])

    @property
    def M(self):
        return int(_lib.vo_hnsw_M(self._h))

    @property
    def M0(self):
        return int(_lib.vo_hnsw_M0(self._h))

    @property
    def max_layer(self):
        return int(_lib.vo_hnsw_max_layer(self._h))

    @property
    def num_layers(self):
        return int(_lib.vo_hnsw_num_layers(self._h))

    @property
    def entry_point(self):
        return int(_lib.vo_hnsw_entry_point(self._h)) if _lib.vo_hnsw_has_entry(self._h) else None

    def vectors(self) -> np.ndarray:
        n, d = len(self), self.dim
        if n == 0:
            return np.zeros((0, d), dtype=np.float32)
        p = _lib.vo_hnsw_vectors(self._h)
        return np.ctypeslib.as_array(p, shape=(n, d)).copy()

    def export_layer(self, l):
        nodes = int(_lib.vo_hnsw_layer_nodes(self._h, l))
        edges = int(_lib.vo_hnsw_layer_edges(self._h, l))
        row_ptr = np.zeros(nodes + 1, dtype=np.uint64)
        cols = np.zeros(max(edges, 1), dtype=np.uint32)
        _lib.vo_hnsw_export_layer(self._h, l, row_ptr, cols)
        return row_ptr, cols[:edges]

    def export_graph(self):
        """[(row_ptr, cols)] per layer, CSR."""
        return [self.export_layer(l) for l in range(self.num_layers)]

    def search(self, q, k, ef, order="reference", with_stats=False):
        q = _f32(q)
        assert q.size == self.dim
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        d = np.zeros(max(k, 1), dtype=np.float32)
        st = np.zeros(6, dtype=np.uint64)
        n = _lib.vo_hnsw_search(self._h, q, k, ef, 0 if order == "reference" else 1, ids, d, st)
        if with_stats:
            return ids[:n].copy(), d[:n].copy(), dict(zip(("ndc0", "hops0", "ndc_up", "hops_up", "tie_at_k", "adj0"),
                                                          (int(x) for x in st)))
        return ids[:n].copy(), d[:n].copy()

    def search_multi_entry(self, q, k, ef, num_probes, order="reference"):
        """NativeHnsw::search_multi_entry (graph.rs:288-348).  Returns ids, dist, stats dict, entry points used."""
        q = _f32(q)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        d = np.zeros(max(k, 1), dtype=np.float32)
        st = np.zeros(6, dtype=np.uint64)
        ent = np.zeros(4, dtype=np.uint64)
        n = _lib.vo_hnsw_search_multi_entry(self._h, q, k, ef, num_probes, 0 if order == "reference" else 1, ids, d, st, ent)
        stats = dict(zip(("ndc0", "hops0", "ndc_up", "hops_up", "tie_at_k", "adj0"), (int(x) for x in st)))
        return ids[:n].copy(), d[:n].copy(), stats, [int(e) for e in ent if e != np.uint64(0xFFFFFFFFFFFFFFFF)]

    @property
    def rng_state(self) -> int:
        return int(_lib.vo_hnsw_rng_state(self._h))

    @rng_state.setter
    def rng_state(self, s: int) -> None:
        _lib.vo_hnsw_set_rng_state(self._h, s)

    def search_batch(self, qs, k, ef, order="canonical", threads=1):
        qs = _f32(qs)
        nq = qs.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint64)
        d = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        st = np.zeros((nq, 6), dtype=np.uint64)
        _lib.vo_hnsw_search_batch(self._h, qs, nq, k, ef, 0 if order == "reference" else 1, threads, ids, d, cnt, st)
        return ids, d, cnt, st

    def search_layer(self, q, entry, ef, layer):
        q = _f32(q)
        ids = np.zeros(max(ef, 1) + 8, dtype=np.uint64)
        d = np.zeros(max(ef, 1) + 8, dtype=np.float32)
        n = _lib.vo_hnsw_search_layer(self._h, q, entry, ef, layer, ids, d)
        return ids[:n].copy(), d[:n].copy()

    def select_neighbors(self, cand_ids, cand_d, maxn):
        ci = np.ascontiguousarray(cand_ids, dtype=np.uint64)
        cd = _f32(cand_d)
        out = np.zeros(max(maxn, 1), dtype=np.uint32)
        n = _lib.vo_hnsw_select_neighbors(self._h, ci, cd, ci.size, maxn, out)
        return out[:n].copy()

    def dump(self, directory, basename="native_hnsw"):
        if _lib.vo_hnsw_dump(self._h, os.fsencode(directory), basename.encode()) != 0:
            raise OSError("oracle: dump failed")

    @classmethod
    def load(cls, directory, metric, basename="native_hnsw", fma=True):
        h = _lib.vo_hnsw_load(os.fsencode(directory), basename.encode(), metric, int(fma))
        if not h:
            raise OSError("oracle: load failed")
        return cls(metric, 0, _handle=h, fma=fma)

    @classmethod
    def from_arrays(cls, metric, vectors, layers, M, M0, entry_point, max_layer, ef_construction=400, fma=True):
        """layers: [(row_ptr u64[nodes+1], cols u32[edges])]"""
        vectors = _f32(vectors)
        n, dim = vectors.shape
        rps = [np.ascontiguousarray(rp, dtype=np.uint64) for rp, _ in layers]
        cls_ = [np.ascontiguousarray(c if len(c) else np.zeros(1, np.uint32), dtype=np.uint32) for _, c in layers]
        nodes = np.array([rp.size - 1 for rp in rps], dtype=np.uint64)
        rp_arr = (C.c_void_p * len(layers))(*[rp.ctypes.data for rp in rps])
        c_arr = (C.c_void_p * len(layers))(*[c.ctypes.data for c in cls_])
        h = _lib.vo_hnsw_from_arrays(metric, dim, M, M0, ef_construction, int(fma), vectors, n, len(layers), rp_arr,
                                     c_arr, nodes, entry_point, max_layer)
        return cls(metric, dim, _handle=h, fma=fma)


STORE_F32, STORE_F16, STORE_BIN = 0, 1, 2
_STORE_OF = {np.dtype(np.float32): STORE_F32, np.dtype(np.float16): STORE_F16, np.dtype(np.uint64): STORE_BIN}


def frozen(metric, vectors, layers, M, M0, entry_point, max_layer, dim=None, ef_construction=400, fma=True):
    """A search-only index over the caller's arrays, nothing copied: CSR adjacency per layer and `vectors` as
    float32 [n, dim], float16 [n, dim] (up-converted per evaluation) or uint64 [n, dim/64] (packed bits, Hamming).
    For the CPU arm of the 10M / 50M node configs, where per-node lists and f32 lanes would not fit."""
    vectors = np.ascontiguousarray(vectors)
    store = _STORE_OF[vectors.dtype]
    n = vectors.shape[0]
    d = dim if dim is not None else (vectors.shape[1] * 64 if store == STORE_BIN else vectors.shape[1])
    rps = [np.ascontiguousarray(rp, dtype=np.uint64) for rp, _ in layers]
    cols = [np.ascontiguousarray(c if len(c) else np.zeros(1, np.uint32), dtype=np.uint32) for _, c in layers]
    nodes = np.array([rp.size - 1 for rp in rps], dtype=np.uint64)
    rp_arr = (C.c_void_p * len(layers))(*[rp.ctypes.data for rp in rps])
    c_arr = (C.c_void_p * len(layers))(*[c.ctypes.data for c in cols])
    h = _lib.vo_hnsw_frozen(metric, d, M, M0, ef_construction, int(fma), store, vectors.ctypes.data, n, len(layers),
                            C.cast(rp_arr, C.c_void_p), C.cast(c_arr, C.c_void_p), nodes, entry_point, max_layer)
    if not h:
        raise ValueError("oracle: unsupported storage / metric combination")
    g = Hnsw(metric, d, _handle=h, fma=fma)
    g._keep = (vectors, rps, cols, nodes, rp_arr, c_arr)  # borrowed by the C side
    return g


def open_index(directory, metric, basename="native_hnsw", vectors=None, dim=None, fma=True):
    """Search-only index from a format-v1 `.graph` file; vectors from `{basename}.vectors` (f32) or, when given, from
    the caller's array in float32 / float16 / packed uint64 form (borrowed)."""
    if vectors is None:
        h = _lib.vo_hnsw_open(os.fsencode(directory), basename.encode(), metric, int(fma), STORE_F32, None, 0, 0)
        keep = None
    else:
        vectors = np.ascontiguousarray(vectors)
        store = _STORE_OF[vectors.dtype]
        d = dim if dim is not None else (vectors.shape[1] * 64 if store == STORE_BIN else vectors.shape[1])
        h = _lib.vo_hnsw_open(os.fsencode(directory), basename.encode(), metric, int(fma), store, vectors.ctypes.data,
                              vectors.shape[0], d)
        keep = vectors
    if not h:
        raise OSError("oracle: open failed")
    g = Hnsw(metric, 0, _handle=h, fma=fma)
    g._keep = keep
    return g


class ScalarQuantizer:
    """ScalarQuantizer + QuantizedVectorStore restated (native/quantization.rs:160-374)."""

    def __init__(self, train_vectors):
        tv = _f32(train_vectors)
        if tv.ndim != 2 or tv.shape[0] == 0:
            raise ValueError("Cannot train on empty vectors")  # quantization.rs:191
        self.dimension = tv.shape[1]
        self._s = _lib.vo_sq8_train(tv, tv.shape[0], tv.shape[1])
        self.min_vals = np.zeros(self.dimension, np.float32)
        self.scales = np.zeros(self.dimension, np.float32)
        self.inv_scales = np.zeros(self.dimension, np.float32)
        _lib.vo_sq8_params(self._s, self.min_vals, self.scales, self.inv_scales)

    def __del__(self):
        if getattr(self, "_s", None) and _lib is not None:
            _lib.vo_sq8_free(self._s)
            self._s = None

    def quantize(self, v) -> np.ndarray:
        v = _f32(v)
        one = v.ndim == 1
        v2 = v.reshape(-1, self.dimension)
        out = np.zeros(v2.shape, np.uint8)
        _lib.vo_sq8_quantize(self._s, v2, v2.shape[0], out)
        return out[0] if one else out

    def dequantize(self, codes) -> np.ndarray:  # quantization.rs:254-268
        return np.asarray(codes, np.uint8).astype(np.float32) * self.inv_scales + self.min_vals

    def push(self, vs) -> None:
        vs = _f32(vs).reshape(-1, self.dimension)
        _lib.vo_sq8_push(self._s, vs, vs.shape[0])

    def __len__(self):
        return int(_lib.vo_sq8_len(self._s))

    def codes(self) -> np.ndarray:
        n = len(self)
        if n == 0:
            return np.zeros((0, self.dimension), np.uint8)
        return np.ctypeslib.as_array(_lib.vo_sq8_codes(self._s), shape=(n, self.dimension)).copy()

    @staticmethod
    def distance_l2_quantized(a, b) -> int:
        a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
        return int(_lib.vo_sq8_distance_quantized(a, b, a.size))

    def distance_l2_asymmetric(self, q, codes) -> float:
        return float(_lib.vo_sq8_distance_asymmetric(self._s, _f32(q), np.ascontiguousarray(codes, np.uint8)))


class DualPrecisionHnsw:
    """DualPrecisionHnsw<SimdDistance> restated (native/dual_precision.rs:60-441)."""

    def __init__(self, metric, dimension, max_connections, ef_construction, max_elements, fma=True, graph=None):
        self.inner = graph if graph is not None else Hnsw(metric, dimension, max_connections, ef_construction, fma=fma)
        self.dimension = dimension
        self.training_sample_size = min(1000, max_elements)  # dual_precision.rs:100
        self.quantizer = None
        self._buffer = []

    @classmethod
    def from_graph(cls, graph: Hnsw, train_count=None):
        """Wraps an existing oracle graph as if its vectors had been inserted in id order."""
        dp = cls(graph.metric, graph.dim, graph.M, 0, max(len(graph), 1), graph=graph)
        vs = graph.vectors()
        t = dp.training_sample_size if train_count is None else train_count
        if len(vs) >= t > 0:
            dp.quantizer = ScalarQuantizer(vs[:t])
            dp.quantizer.push(vs)
        else:
            dp._buffer = [v for v in vs]
        return dp

    def __len__(self):
        return len(self.inner)

    def is_quantizer_trained(self) -> bool:
        return self.quantizer is not None

    def insert(self, v) -> int:  # dual_precision.rs:122-143
        v = _f32(v)
        node = self.inner.insert(v)
        if self.quantizer is not None:
            self.quantizer.push(v)
        else:
            self._buffer.append(v.copy())
            if len(self._buffer) >= self.training_sample_size:
                self._train()
        return node

    def _train(self):  # dual_precision.rs:146-169
        if not self._buffer:
            return
        buf = np.stack(self._buffer)
        self.quantizer = ScalarQuantizer(buf)
        self.quantizer.push(buf)
        self._buffer = []

    def force_train_quantizer(self):  # dual_precision.rs:172-176
        if self.quantizer is None and self._buffer:
            self._train()

    def search(self, q, k, ef_search):  # dual_precision.rs:179-228 (f32 traversal + exact re-rank = inner.search)
        if self.quantizer is None:
            return self.inner.search(q, k, ef_search)
        rerank_k = max(ef_search * 2, k * 4)
        ids, d = self.inner.search(q, rerank_k, ef_search)
        order = np.argsort(np.array([_total_key(x) for x in d], dtype=np.int64), kind="stable")[:k]
        return ids[order], d[order]

    def search_with_config(self, q, k, ef_search, oversampling_ratio=4, use_int8_traversal=True, min_index_size=10_000,
                           order="reference", with_stats=False):  # dual_precision.rs:263-325
        if self.quantizer is None or not use_int8_traversal or len(self.inner) < min_index_size:
            r = self.inner.search(q, k, ef_search, order=order, with_stats=with_stats)
            return r
        q = _f32(q)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        d = np.zeros(max(k, 1), dtype=np.float32)
        st = np.zeros(6, dtype=np.uint64)
        n = _lib.vo_dual_search_int8(self.inner._h, self.quantizer._s, q, k, ef_search, oversampling_ratio,
                                     0 if order == "reference" else 1, ids, d, st)
        if with_stats:
            return ids[:n].copy(), d[:n].copy(), dict(zip(("ndc0", "hops0", "ndc_up", "hops_up", "tie_at_k", "adj0"),
                                                          (int(x) for x in st)))
        return ids[:n].copy(), d[:n].copy()

    def search_int8_batch(self, qs, k, ef_search, oversampling_ratio=4, order="canonical", threads=1):
        assert self.quantizer is not None
        qs = _f32(qs)
        nq = qs.shape[0]
        ids = np.zeros((nq, k), dtype=np.uint64)
        d = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        st = np.zeros((nq, 6), dtype=np.uint64)
        _lib.vo_dual_search_int8_batch(self.inner._h, self.quantizer._s, qs, nq, k, ef_search, oversampling_ratio,
                                       0 if order == "reference" else 1, threads, ids, d, cnt, st)
        return ids, d, cnt, st


def _total_key(x) -> int:
    b = int(np.float32(x).view(np.int32))
    return b ^ (((b >> 31) & 0xFFFFFFFF) >> 1) if b < 0 else b


def bruteforce(metric, vectors, q, k, fma=True):
    vectors, q = _f32(vectors), _f32(q)
    n, dim = vectors.shape
    ids = np.zeros(max(k, 1), dtype=np.uint64)
    sc = np.zeros(max(k, 1), dtype=np.float32)
    m = _lib.vo_bruteforce(metric, vectors, n, dim, q, k, int(fma), ids, sc)
    return ids[:m].copy(), sc[:m].copy()


def bruteforce_batch(metric, vectors, qs, k, fma=True, threads=1):
    vectors, qs = _f32(vectors), _f32(qs)
    n, dim = vectors.shape
    nq = qs.shape[0]
    ids = np.zeros((nq, k), dtype=np.uint64)
    sc = np.zeros((nq, k), dtype=np.float32)
    _lib.vo_bruteforce_batch(metric, vectors, n, dim, qs, nq, k, int(fma), threads, ids, sc)
    return ids, sc


def bruteforce_binary(vectors_u64, q_u64, k):
    v = np.ascontiguousarray(vectors_u64, dtype=np.uint64)
    q = np.ascontiguousarray(q_u64, dtype=np.uint64)
    n, words = v.shape
    ids = np.zeros(max(k, 1), dtype=np.uint64)
    d = np.zeros(max(k, 1), dtype=np.uint32)
    m = _lib.vo_bruteforce_binary(v, n, words, q, k, ids, d)
    return ids[:m].copy(), d[:m].copy()


_TOKEN_SPLIT = re.compile(r"[^\w]|_", re.UNICODE)


def tokenize(text: str):
    """bm25.rs:114-120: lowercase, split on non-alphanumeric, drop tokens of byte length <= 1."""
    out = []
    cur = []
    for ch in text.lower():
        if ch.isalnum():
            cur.append(ch)
        else:
            if cur:
                out.append("".join(cur))
                cur = []
    if cur:
        out.append("".join(cur))
    return [t for t in out if len(t.encode("utf-8")) > 1]


class Bm25:
    """Bm25Index restated (index/bm25.rs); strings are mapped to term ids here."""

    def __init__(self, k1=1.2, b=0.75):
        self._h = _lib.vo_bm25_new(k1, b)
        self.vocab = {}

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.vo_bm25_free(self._h)
            self._h = None

    def _ids(self, tokens, grow):
        out = []
        for t in tokens:
            if t not in self.vocab:
                if not grow:
                    out.append(0xFFFFFFFF)  # unknown term: df = 0
                    continue
                self.vocab[t] = len(self.vocab)
            out.append(self.vocab[t])
        return np.array(out, dtype=np.uint32)

    def add_document(self, doc_id, text):
        assert 0 <= doc_id <= 0xFFFFFFFF, "BM25 document ID exceeds u32::MAX"
        t = self._ids(tokenize(text), True)
        if t.size:
            _lib.vo_bm25_add(self._h, doc_id, t, t.size)

    def add_document_terms(self, doc_id, term_ids):
        t = np.ascontiguousarray(term_ids, dtype=np.uint32)
        if t.size:
            _lib.vo_bm25_add(self._h, doc_id, t, t.size)

    def remove_document(self, doc_id) -> bool:
        return bool(_lib.vo_bm25_remove(self._h, doc_id))

    def __len__(self):
        return int(_lib.vo_bm25_len(self._h))

    def term_count(self):
        return int(_lib.vo_bm25_term_count(self._h))

    def search_terms(self, term_ids, k):
        t = np.ascontiguousarray(term_ids, dtype=np.uint32)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        if t.size == 0:
            return ids[:0], sc[:0]
        n = _lib.vo_bm25_search(self._h, t, t.size, k, ids, sc)
        return ids[:n].copy(), sc[:n].copy()

    def search(self, text, k):
        return self.search_terms(self._ids(tokenize(text), False), k)

    def search_batch_terms(self, q_ptr, q_terms, k, threads=1):
        q_ptr = np.ascontiguousarray(q_ptr, dtype=np.uint32)
        q_terms = np.ascontiguousarray(q_terms, dtype=np.uint32)
        nq = q_ptr.size - 1
        ids = np.zeros((nq, k), dtype=np.uint64)
        sc = np.zeros((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        _lib.vo_bm25_search_batch(self._h, q_ptr, q_terms, nq, k, threads, ids, sc, cnt)
        return ids, sc, cnt


def rrf_hybrid(vec_ids, txt_ids, k, vector_weight=0.5):
    v = np.ascontiguousarray(vec_ids, dtype=np.uint64)
    t = np.ascontiguousarray(txt_ids, dtype=np.uint64)
    ids = np.zeros(max(k, 1), dtype=np.uint64)
    sc = np.zeros(max(k, 1), dtype=np.float32)
    vv = v if v.size else np.zeros(1, np.uint64)
    tt = t if t.size else np.zeros(1, np.uint64)
    n = _lib.vo_rrf_hybrid(vv, v.size, tt, t.size, vector_weight, k, ids, sc)
    return ids[:n].copy(), sc[:n].copy()


AVERAGE, MAXIMUM, RRF, WEIGHTED = 0, 1, 2, 3


def fuse(strategy, lists, rrf_k=60, avg_w=0.0, max_w=0.0, hit_w=0.0):
    """lists: [[(id, score), ...], ...] -> (ids, scores) sorted score-desc (ties id-asc)."""
    ptr = np.zeros(len(lists) + 1, dtype=np.uint32)
    ids, sc = [], []
    for i, l in enumerate(lists):
        for d, s in l:
            ids.append(d)
            sc.append(s)
        ptr[i + 1] = len(ids)
    cap = max(len(ids), 1)
    ia = np.array(ids if ids else [0], dtype=np.uint64)
    sa = np.array(sc if sc else [0], dtype=np.float32)
    oi = np.zeros(cap, dtype=np.uint64)
    os_ = np.zeros(cap, dtype=np.float32)
    n = _lib.vo_fuse(strategy, ptr, len(lists), ia, sa, rrf_k, avg_w, max_w, hit_w, cap, oi, os_)
    return oi[:n].copy(), os_[:n].copy()
