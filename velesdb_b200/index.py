"""Host-side mirror of velesdb-core's ``HnswIndex`` over the C ABI (include/veles_b200.h).

Same method names, argument meaning and error behaviour as the Rust type
(crates/velesdb-core/src/index/hnsw/index/{mod,search,batch,trait_impl,constructors,vacuum}.rs and
the ``VectorIndex`` trait, index/mod.rs:30-83), so the parity tests read like the reference's.
Everything numeric happens in libveles_b200.so on the GPU; this file only does what the Rust
wrapper does on the host: id <-> node mapping, tombstones, quality -> ef, score transform.

A Rust panic (``assert_eq!`` on dimensions) is raised here as ``DimensionMismatch`` (a ValueError).
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import struct

import numpy as np

from . import _native as nv


class DistanceMetric(enum.IntEnum):
    """core/distance.rs:16-39"""
    Cosine = 0
    Euclidean = 1
    DotProduct = 2
    Hamming = 3
    Jaccard = 4

    def higher_is_better(self) -> bool:  # core/distance.rs:76-81
        return self in (DistanceMetric.Cosine, DistanceMetric.DotProduct, DistanceMetric.Jaccard)


class SearchQuality:
    """index/hnsw/params.rs:283-320"""

    def __init__(self, kind, ef=0):
        self.kind, self.ef = kind, ef

    def ef_search(self, k: int) -> int:
        return int(nv.lib().veles_ef_search(self.kind, k, self.ef))

    @staticmethod
    def Custom(ef: int) -> "SearchQuality":
        return SearchQuality(nv.CUSTOM, ef)

    def __repr__(self):
        return f"SearchQuality({['Fast', 'Balanced', 'Accurate', 'Perfect', 'Custom'][self.kind]}, {self.ef})"


SearchQuality.Fast = SearchQuality(nv.FAST)
SearchQuality.Balanced = SearchQuality(nv.BALANCED)
SearchQuality.Accurate = SearchQuality(nv.ACCURATE)
SearchQuality.Perfect = SearchQuality(nv.PERFECT)


class HnswParams:
    """``HnswParams`` (index/hnsw/params.rs:13-289): M, ef_construction, initial capacity and storage mode, with the
    reference's presets.  ``storage_mode`` is one of "full", "sq8", "binary" (quantization.rs ``StorageMode``)."""

    # (dimension <= 256, dimension > 256) -> (M, ef_construction) per expected dataset size (params.rs:71-141)
    _BY_SIZE = ((10_000, 20_000, (24, 200), (32, 400)), (100_000, 150_000, (64, 800), (128, 1600)),
                (500_000, 750_000, (96, 1200), (128, 2000)), (None, 1_500_000, (64, 800), (128, 1600)))

    def __init__(self, max_connections=32, ef_construction=400, max_elements=100_000, storage_mode="full"):
        self.max_connections, self.ef_construction, self.max_elements = max_connections, ef_construction, max_elements
        self.storage_mode = storage_mode

    def __eq__(self, other):
        return isinstance(other, HnswParams) and vars(self) == vars(other)

    def __repr__(self):
        return (f"HnswParams(max_connections={self.max_connections}, ef_construction={self.ef_construction}, "
                f"max_elements={self.max_elements}, storage_mode={self.storage_mode!r})")

    @staticmethod
    def default() -> "HnswParams":                                   # params.rs:29-33
        return HnswParams.auto(768)

    @staticmethod
    def auto(dimension: int) -> "HnswParams":                        # params.rs:40-57
        return HnswParams(24, 300) if dimension <= 256 else HnswParams(32, 400)

    @staticmethod
    def for_dataset_size(dimension: int, expected_vectors: int) -> "HnswParams":   # params.rs:71-141
        for limit, cap, small, large in HnswParams._BY_SIZE:
            if limit is None or expected_vectors <= limit:
                m, efc = small if dimension <= 256 else large
                return HnswParams(m, efc, cap)
        raise AssertionError("unreachable")

    @staticmethod
    def large_dataset(dimension: int) -> "HnswParams":               # params.rs:147-150
        return HnswParams.for_dataset_size(dimension, 500_000)

    @staticmethod
    def million_scale(dimension: int) -> "HnswParams":               # params.rs:155-158
        return HnswParams.for_dataset_size(dimension, 1_000_000)

    @staticmethod
    def fast() -> "HnswParams":                                      # params.rs:162-170
        return HnswParams(16, 150)

    @staticmethod
    def turbo() -> "HnswParams":                                     # params.rs:189-197
        return HnswParams(12, 100)

    @staticmethod
    def high_recall(dimension: int) -> "HnswParams":                 # params.rs:200-208
        b = HnswParams.auto(dimension)
        return HnswParams(b.max_connections + 8, b.ef_construction + 200, b.max_elements)

    @staticmethod
    def max_recall(dimension: int) -> "HnswParams":                  # params.rs:211-233
        if dimension <= 256:
            return HnswParams(32, 500)
        return HnswParams(48, 800) if dimension <= 768 else HnswParams(64, 1000)

    @staticmethod
    def fast_indexing(dimension: int) -> "HnswParams":               # params.rs:236-244
        b = HnswParams.auto(dimension)
        return HnswParams(max(b.max_connections // 2, 8), b.ef_construction // 2, b.max_elements)

    @staticmethod
    def custom(max_connections: int, ef_construction: int, max_elements: int) -> "HnswParams":   # params.rs:247-259
        return HnswParams(max_connections, ef_construction, max_elements)

    @staticmethod
    def with_sq8(dimension: int) -> "HnswParams":                    # params.rs:271-276
        p = HnswParams.auto(dimension)
        p.storage_mode = "sq8"
        return p

    @staticmethod
    def with_binary(dimension: int) -> "HnswParams":                 # params.rs:280-285
        p = HnswParams.auto(dimension)
        p.storage_mode = "binary"
        return p


class VacuumError(RuntimeError):
    """index/hnsw/index/vacuum.rs:17-29"""


class DimensionMismatch(ValueError):
    pass


_DT = {"f32": nv.F32, "f16": nv.F16, "bin1": nv.BIN1}


def _as_f32_2d(a, dim, what):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[1] != dim:
        raise DimensionMismatch(f"{what} dimension mismatch: expected {dim}, got {a.shape[1]}")
    return a


class DeviceSnapshot:
    """Owns one ``veles_index_t*``."""

    def __init__(self, handle):
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            try:
                nv.lib().veles_index_free(self.h)
            except Exception:  # interpreter shutdown
                pass
            self.h = None

    @staticmethod
    def from_vectors(vectors, metric, store_dtype="f32", src_dtype="f32", dim=None):
        nv.init()
        vectors = np.ascontiguousarray(vectors)
        n = vectors.shape[0]
        d = dim if dim is not None else vectors.shape[1]
        h = C.c_void_p()
        nv.check(nv.lib().veles_index_from_vectors(nv.ptr(vectors), n, d, _DT[src_dtype], _DT[store_dtype], int(metric),
                                                   C.byref(h)))
        return DeviceSnapshot(h)

    @staticmethod
    def create(n, dim, metric, store_dtype="f32"):
        """An empty snapshot of n zero rows, to be filled from device memory with set_rows_device."""
        nv.init()
        h = C.c_void_p()
        nv.check(nv.lib().veles_index_create(n, dim, _DT[store_dtype], int(metric), C.byref(h)))
        return DeviceSnapshot(h)

    def set_rows_device(self, first, rows_t, src_dtype="f32", stream=None):
        """rows_t: a contiguous CUDA tensor holding rows [first, first + len(rows_t)) as f32, or in the store type."""
        nv.check(nv.lib().veles_index_set_rows_d(self.h, first, rows_t.shape[0], nv.ptr(rows_t), _DT[src_dtype], stream))

    def get_rows(self, first, count, dtype):
        """Rows back on the host in the store type: dtype np.float32 / np.float16 / np.uint64 (packed bits)."""
        width = {np.dtype(np.float32): self.dim, np.dtype(np.float16): self.dim, np.dtype(np.uint64): self.dim // 64}[np.dtype(dtype)]
        out = np.empty((count, width), dtype=dtype)
        nv.check(nv.lib().veles_index_get_rows(self.h, first, count, nv.ptr(out)))
        return out

    def dump_graph(self, directory, basename="native_hnsw"):
        nv.check(nv.lib().veles_index_dump_graph(self.h, os.fsencode(directory), basename.encode()))

    @staticmethod
    def from_arrays(vectors, metric, layers, M, M0, entry_point, max_layer, store_dtype="f32", src_dtype="f32",
                    dim=None):
        """layers: [(row_ptr u64[nodes+1], cols u32[edges])] per layer (CSR)."""
        nv.init()
        vectors = np.ascontiguousarray(vectors)
        n = vectors.shape[0]
        d = dim if dim is not None else vectors.shape[1]
        rps = [np.ascontiguousarray(rp, dtype=np.uint64) for rp, _ in layers]
        cls = [np.ascontiguousarray(c if len(c) else np.zeros(1, np.uint32), dtype=np.uint32) for _, c in layers]
        nodes = np.array([rp.size - 1 for rp in rps], dtype=np.uint64)
        rp_arr = (C.c_void_p * len(layers))(*[rp.ctypes.data for rp in rps])
        c_arr = (C.c_void_p * len(layers))(*[c.ctypes.data for c in cls])
        h = C.c_void_p()
        nv.check(nv.lib().veles_index_from_arrays(nv.ptr(vectors), n, d, _DT[src_dtype], _DT[store_dtype], int(metric),
                                                  len(layers), rp_arr, c_arr, nv.ptr(nodes), M, M0,
                                                  entry_point or 0, max_layer, C.byref(h)))
        return DeviceSnapshot(h)

    @staticmethod
    def from_reference_files(directory, metric, basename="native_hnsw", store_dtype="f32"):
        nv.init()
        h = C.c_void_p()
        nv.check(nv.lib().veles_index_from_reference_files(os.fsencode(directory), basename.encode(), int(metric),
                                                           _DT[store_dtype], C.byref(h)))
        return DeviceSnapshot(h)

    # -- properties
    def __len__(self):
        return int(nv.lib().veles_index_len(self.h))

    @property
    def dim(self):
        return int(nv.lib().veles_index_dim(self.h))

    @property
    def metric(self):
        return DistanceMetric(nv.lib().veles_index_metric(self.h))

    @property
    def max_layer(self):
        return int(nv.lib().veles_index_max_layer(self.h))

    @property
    def entry_point(self):
        return int(nv.lib().veles_index_entry_point(self.h))

    @property
    def device_bytes(self):
        return int(nv.lib().veles_index_device_bytes(self.h))

    # -- calls
    def search_batch(self, queries, k, ef, with_stats=False, stream=None):
        q = _as_f32_2d(queries, self.dim, "Query")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        dist = np.empty((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        st = np.zeros((nq, 4), dtype=np.uint32) if with_stats else None
        nv.check(nv.lib().veles_search_batch(self.h, nv.ptr(q), nq, k, ef, nv.ptr(ids), nv.ptr(dist), nv.ptr(cnt),
                                             nv.ptr(st), stream))
        return (ids, dist, cnt, st) if with_stats else (ids, dist, cnt)

    def search_batch_device(self, q_t, k, ef, ids_t, dist_t, cnt_t, stats_t=None, stream=None):
        """torch CUDA tensors in, results written to torch CUDA tensors; only enqueues on `stream`."""
        nv.check(nv.lib().veles_search_batch_d(self.h, nv.ptr(q_t), q_t.shape[0], k, ef, nv.ptr(ids_t), nv.ptr(dist_t),
                                               nv.ptr(cnt_t), nv.ptr(stats_t), stream))

    def bruteforce_batch(self, queries, k, stream=None):
        q = _as_f32_2d(queries, self.dim, "Query")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        sc = np.empty((nq, k), dtype=np.float32)
        nv.check(nv.lib().veles_bruteforce_batch(self.h, nv.ptr(q), nq, k, nv.ptr(ids), nv.ptr(sc), stream))
        return ids, sc

    def bruteforce_batch_device(self, q_t, k, ids_t, score_t, stream=None):
        nv.check(nv.lib().veles_bruteforce_batch_d(self.h, nv.ptr(q_t), q_t.shape[0], k, nv.ptr(ids_t), nv.ptr(score_t),
                                                   stream))

    def bruteforce_batch_relaxed(self, queries, k, oversample=4, stream=None):
        """Tensor-core candidate GEMM (fp16) + exact re-rank: veles_bruteforce_batch_relaxed."""
        q = _as_f32_2d(queries, self.dim, "Query")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        sc = np.empty((nq, k), dtype=np.float32)
        nv.check(nv.lib().veles_bruteforce_batch_relaxed(self.h, nv.ptr(q), nq, k, oversample, nv.ptr(ids), nv.ptr(sc), stream))
        return ids, sc

    def bruteforce_batch_relaxed_device(self, q_t, k, oversample, ids_t, score_t, stream=None, want_gemm_ms=False):
        ms = C.c_float(0.0)
        nv.check(nv.lib().veles_bruteforce_batch_relaxed_d(self.h, nv.ptr(q_t), q_t.shape[0], k, oversample, nv.ptr(ids_t),
                                                           nv.ptr(score_t), C.byref(ms) if want_gemm_ms else None, stream))
        return ms.value

    def rerank_batch(self, queries, cand, stream=None):
        q = _as_f32_2d(queries, self.dim, "Query")
        cand = np.ascontiguousarray(cand, dtype=np.uint32).reshape(q.shape[0], -1)
        out = np.empty(cand.shape, dtype=np.float32)
        nv.check(nv.lib().veles_rerank_batch(self.h, nv.ptr(q), q.shape[0], nv.ptr(cand), cand.shape[1], nv.ptr(out),
                                             stream))
        return out

    def build_graph(self, M, ef_construction=0, stream=None):
        """Bulk construction by block insertion (insert_batch_parallel's role, batch.rs:82-108); 0 = library default."""
        nv.check(nv.lib().veles_index_build_graph(self.h, M, ef_construction, stream))

    def search_status(self, stream=None):
        """Waits for `stream`; raises VelesError(OVERFLOW) if a `_d` search enqueued on it overflowed its tie list."""
        nv.check(nv.lib().veles_search_status(self.h, stream))

    def search_submit(self, q, k, ef, ids, dist, cnt):
        """veles_search_submit: host arrays (pinned for true overlap) in/out, returns a ticket for search_wait."""
        t = C.c_uint64()
        nv.check(nv.lib().veles_search_submit(self.h, nv.ptr(q), q.shape[0], k, ef, nv.ptr(ids), nv.ptr(dist), nv.ptr(cnt),
                                              C.byref(t)))
        return t.value

    def search_wait(self, ticket):
        nv.check(nv.lib().veles_search_wait(self.h, ticket))

    def append(self, vectors, ef_construction=0, src_dtype="f32", stream=None):
        """veles_index_append: more vectors linked into the live graph (insert_batch_parallel on an existing index)."""
        vectors = np.ascontiguousarray(vectors)
        nv.check(nv.lib().veles_index_append(self.h, nv.ptr(vectors), vectors.shape[0], _DT[src_dtype], ef_construction, stream))

    def build_graph_exact(self, M, ef_construction, stream=None):
        """NativeHnsw::insert for nodes 0..n-1 in order (graph.rs:158-237): the reference's deterministic graph."""
        nv.check(nv.lib().veles_index_build_graph_exact(self.h, M, ef_construction, stream))

    # -- SQ8 dual precision (native/dual_precision.rs)
    def attach_sq8(self, train_count, stream=None):
        """Trains the ScalarQuantizer on the first `train_count` vectors and codes every vector."""
        nv.check(nv.lib().veles_index_attach_sq8(self.h, train_count, stream))

    @property
    def has_sq8(self) -> bool:
        return bool(nv.lib().veles_index_has_sq8(self.h))

    def sq8_export(self, with_codes=True):
        d = self.dim
        mn, sc, inv = (np.empty(d, np.float32) for _ in range(3))
        codes = np.empty((len(self), d), np.uint8) if with_codes else None
        nv.check(nv.lib().veles_index_sq8_export(self.h, nv.ptr(mn), nv.ptr(sc), nv.ptr(inv), nv.ptr(codes)))
        return mn, sc, inv, codes

    def search_batch_sq8(self, queries, k, ef_search, oversampling=4, with_stats=False, stream=None):
        q = _as_f32_2d(queries, self.dim, "Query")
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.uint32)
        dist = np.empty((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        st = np.zeros((nq, 4), dtype=np.uint32) if with_stats else None
        nv.check(nv.lib().veles_search_batch_sq8(self.h, nv.ptr(q), nq, k, ef_search, oversampling, nv.ptr(ids),
                                                 nv.ptr(dist), nv.ptr(cnt), nv.ptr(st), stream))
        return (ids, dist, cnt, st) if with_stats else (ids, dist, cnt)

    def search_batch_sq8_device(self, q_t, k, ef_search, oversampling, ids_t, dist_t, cnt_t, stats_t=None, stream=None):
        nv.check(nv.lib().veles_search_batch_sq8_d(self.h, nv.ptr(q_t), q_t.shape[0], k, ef_search, oversampling,
                                                   nv.ptr(ids_t), nv.ptr(dist_t), nv.ptr(cnt_t), nv.ptr(stats_t), stream))

    # -- NativeHnsw::search_multi_entry (native/graph.rs:288-348)
    def search_batch_multi_entry(self, queries, k, ef, extra_entries, with_stats=False, stream=None):
        """extra_entries: [nq, 3] node ids (INVALID padded) added to the greedy result as layer-0 entry points."""
        q = _as_f32_2d(queries, self.dim, "Query")
        nq = q.shape[0]
        ex = np.ascontiguousarray(extra_entries, dtype=np.uint32).reshape(nq, 3)
        ids = np.empty((nq, k), dtype=np.uint32)
        dist = np.empty((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        st = np.zeros((nq, 4), dtype=np.uint32) if with_stats else None
        nv.check(nv.lib().veles_search_batch_multi_entry(self.h, nv.ptr(q), nq, k, ef, nv.ptr(ex), nv.ptr(ids), nv.ptr(dist),
                                                         nv.ptr(cnt), nv.ptr(st), stream))
        return (ids, dist, cnt, st) if with_stats else (ids, dist, cnt)

    # -- id map, tombstones, filtered search (ShardedMappings on the device)
    def set_id_map(self, ext_ids=None, live_bits=None):
        """ext_ids: u64[n] external id per node (None = identity); live_bits: one bit per node, 0 = removed."""
        e = None if ext_ids is None else np.ascontiguousarray(ext_ids, dtype=np.uint64)
        b = None if live_bits is None else np.ascontiguousarray(live_bits, dtype=np.uint32)
        assert e is None or e.size == len(self)
        assert b is None or b.size == (len(self) + 31) // 32
        nv.check(nv.lib().veles_index_set_id_map(self.h, nv.ptr(e), nv.ptr(b)))

    def search_batch_mapped(self, queries, k, ef, k_fetch=None, allow_bits=None, stream=None):
        """search + tombstone/filter drop + node -> external id + transform_score in one call (batch.rs:178-196)."""
        q = _as_f32_2d(queries, self.dim, "Query")
        nq = q.shape[0]
        kf = k if k_fetch is None else k_fetch
        a = None if allow_bits is None else np.ascontiguousarray(allow_bits, dtype=np.uint32)
        ids = np.empty((nq, k), dtype=np.uint64)
        sc = np.empty((nq, k), dtype=np.float32)
        cnt = np.zeros(nq, dtype=np.uint32)
        nv.check(nv.lib().veles_search_batch_mapped(self.h, nv.ptr(q), nq, kf, k, ef, nv.ptr(a), nv.ptr(ids), nv.ptr(sc),
                                                    nv.ptr(cnt), stream))
        return ids, sc, cnt

    def export_layer(self, layer):
        nodes, edges = C.c_uint64(), C.c_uint64()
        nv.check(nv.lib().veles_index_export_layer(self.h, layer, C.byref(nodes), C.byref(edges), None, None))
        rp = np.zeros(nodes.value + 1, dtype=np.uint64)
        cols = np.zeros(max(edges.value, 1), dtype=np.uint32)
        nv.check(nv.lib().veles_index_export_layer(self.h, layer, C.byref(nodes), C.byref(edges), nv.ptr(rp),
                                                   nv.ptr(cols)))
        return rp, cols[:edges.value]

    def export_graph(self):
        return [self.export_layer(l) for l in range(self.max_layer + 1)]

    def dump(self, directory, basename="native_hnsw"):
        nv.check(nv.lib().veles_index_dump(self.h, os.fsencode(directory), basename.encode()))


def multi_entry_probes(rng_state: int, count: int, num_probes: int):
    """The extra entry points NativeHnsw::search_multi_entry draws (graph.rs:313-340): xorshift64 (13, 7, 17) on the
    shared state, `state % count`, only when count > 10 and num_probes > 1, at most 3 draws.  Returns the three-slot
    row for veles_search_batch_multi_entry (INVALID padded) and the advanced state, which the caller keeps."""
    m64 = (1 << 64) - 1
    row = [nv.INVALID_ID] * 3
    if num_probes > 1 and count > 10:
        for i in range(min(num_probes, 4) - 1):
            rng_state ^= (rng_state << 13) & m64
            rng_state ^= rng_state >> 7
            rng_state ^= (rng_state << 17) & m64
            row[i] = rng_state % count
    return row, rng_state


def distance_pairs(metric, a, b, as_metric_value=False):
    """DistanceEngine::batch_distance on explicit pairs (native/distance.rs:22-24)."""
    nv.init()
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    if a.ndim == 1:
        a, b = a.reshape(1, -1), b.reshape(1, -1)
    assert a.shape == b.shape, "Vector dimensions must match"
    out = np.empty(a.shape[0], dtype=np.float32)
    nv.check(nv.lib().veles_distance_pairs(int(metric), nv.ptr(a), nv.ptr(b), a.shape[0], a.shape[1],
                                           1 if as_metric_value else 0, nv.ptr(out), None))
    return out


def _bincode_mappings(id_to_idx, idx_to_id, next_idx) -> bytes:
    out = [struct.pack("<Q", len(id_to_idx))]
    out += [struct.pack("<QQ", int(k), int(v)) for k, v in id_to_idx.items()]
    out.append(struct.pack("<Q", len(idx_to_id)))
    out += [struct.pack("<QQ", int(k), int(v)) for k, v in idx_to_id.items()]
    out.append(struct.pack("<Q", int(next_idx)))
    return b"".join(out)


def _parse_bincode_mappings(raw: bytes):
    def read_map(off):
        if off + 8 > len(raw):
            raise OSError("native_mappings.bin is truncated")
        (n,) = struct.unpack_from("<Q", raw, off)
        off += 8
        if n > (len(raw) - off) // 16:
            raise OSError("native_mappings.bin is truncated")
        flat = np.frombuffer(raw, dtype="<u8", count=2 * n, offset=off)
        return {int(flat[2 * i]): int(flat[2 * i + 1]) for i in range(n)}, off + 16 * n

    a, off = read_map(0)
    b, off = read_map(off)
    if off + 8 > len(raw):
        raise OSError("native_mappings.bin is truncated")
    (next_idx,) = struct.unpack_from("<Q", raw, off)
    return a, b, int(next_idx)


class HnswIndex:
    """``HnswIndex`` (index/hnsw/index/mod.rs:93-131) over a GPU snapshot.

    The reference mutates its graph in place under a RwLock.  Here inserts are staged on the host
    and the immutable device snapshot is (re)built on the next search -- the bulk path the reference
    calls ``insert_batch_parallel`` (index/hnsw/index/batch.rs:82-108), whose graph is order
    dependent in the reference too.  Indexes loaded from the reference's files keep the file's graph.
    """

    EXACT_BUILD_LIMIT = 20_000  # above this the one-warp sequential builder is too slow; use the bulk builder

    def __init__(self, dimension, metric, params=None, enable_vector_storage=True, store_dtype="f32"):
        self._dimension = int(dimension)
        self._metric = DistanceMetric(metric)
        self._params = params or HnswParams.auto(self._dimension)
        self.enable_vector_storage = enable_vector_storage
        self._store_dtype = store_dtype
        self._id_to_idx = {}
        self._idx_to_id = {}
        self._next_idx = 0
        self._staged = []          # vectors by node index
        self._snapshot = None
        self._dirty = False
        self._map_dirty = True        # the device copy of the id map / live bitmap is stale
        self._bulk = False            # True once a batch of >= 100 vectors went through insert_batch_parallel
        self._vectors_present = True  # False after load(): ShardedVectors is left empty (constructors.rs:240)

    # ---- constructors (index/hnsw/index/constructors.rs:29-177)
    @classmethod
    def new(cls, dimension, metric):
        return cls(dimension, metric)

    @classmethod
    def new_fast_insert(cls, dimension, metric):
        """constructors.rs:63-66: no ShardedVectors copy -- no brute force, no re-rank, ``vacuum`` is an error."""
        return cls(dimension, metric, HnswParams.auto(int(dimension)), enable_vector_storage=False)

    @classmethod
    def new_turbo(cls, dimension, metric):
        """constructors.rs:86-91: auto parameters with ef_construction raised by half."""
        p = HnswParams.auto(int(dimension))
        p.ef_construction = (p.ef_construction * 3) // 2
        return cls(dimension, metric, p)

    @classmethod
    def with_params(cls, dimension, metric, params):
        return cls(dimension, metric, params)

    @classmethod
    def with_params_full(cls, dimension, metric, params, enable_vector_storage):
        return cls(dimension, metric, params, enable_vector_storage)

    @classmethod
    def from_snapshot(cls, snapshot: DeviceSnapshot, ids=None, vectors_present=False):
        ix = cls(snapshot.dim, snapshot.metric)
        n = len(snapshot)
        ids = list(range(n)) if ids is None else list(ids)
        ix._id_to_idx = {int(e): i for i, e in enumerate(ids)}
        ix._idx_to_id = {i: int(e) for i, e in enumerate(ids)}
        ix._next_idx = n
        ix._snapshot = snapshot
        ix._vectors_present = vectors_present
        return ix

    # ---- VectorIndex (index/mod.rs:30-83; trait_impl.rs)
    def insert(self, id, vector):
        v = np.asarray(vector, dtype=np.float32).reshape(-1)
        if v.size != self._dimension:
            raise DimensionMismatch(f"Vector dimension mismatch: expected {self._dimension}, got {v.size}")
        if id in self._id_to_idx:  # duplicate ids are skipped silently (trait_impl.rs:12-25)
            return
        self._materialise_staged()
        idx = self._next_idx
        self._next_idx += 1
        self._id_to_idx[id] = idx
        self._idx_to_id[idx] = id
        self._staged.append(v.copy())
        self._dirty = True
        self._map_dirty = True

    def insert_batch_parallel(self, vectors) -> int:
        """index/hnsw/index/batch.rs:82-108: (id, vector) pairs; returns how many were new.  Batches of fewer
        than 100 vectors are inserted sequentially by the reference too (backend_adapter.rs:110-118).  A batch of
        >= 100 vectors into an index whose snapshot is live goes straight into that graph on the device
        (veles_index_append), as the reference inserts into the graph it has; otherwise the vectors are staged and the
        next search builds the snapshot."""
        vectors = list(vectors)
        if len(vectors) >= 100:
            self._bulk = True
        live = self._snapshot is not None and not self._dirty and len(vectors) >= 100 and self._next_idx > 0
        if live and self._store_dtype != "bin1":
            self._materialise_staged()
            fresh, seen = [], set()
            for id, v in vectors:
                if id in self._id_to_idx or id in seen:
                    continue  # duplicate ids are skipped silently (trait_impl.rs:12-25)
                v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1)
                if v.shape[0] != self._dimension:
                    raise DimensionMismatch(f"Vector dimension mismatch: expected {self._dimension}, got {v.shape[0]}")
                seen.add(id)
                fresh.append((id, v))
            if not fresh:
                return 0
            self._snapshot.append(np.stack([v for _, v in fresh]), min(self._params.ef_construction, 4096))
            for id, v in fresh:
                idx = self._next_idx
                self._next_idx += 1
                self._id_to_idx[id] = idx
                self._idx_to_id[idx] = id
                self._staged.append(v.copy())
            self._map_dirty = True
            return len(fresh)
        count = 0
        for id, v in vectors:
            before = len(self._id_to_idx)
            self.insert(id, v)
            count += len(self._id_to_idx) - before
        return count

    def insert_batch_sequential(self, vectors) -> int:
        """index/hnsw/index/batch.rs:120-139 (deprecated in the reference): one NativeHnsw::insert per new vector,
        in order -- the exact sequential builder's path here."""
        count = 0
        for id, v in vectors:
            before = len(self._id_to_idx)
            self.insert(id, v)
            count += len(self._id_to_idx) - before
        return count

    def remove(self, id) -> bool:
        """Soft delete (trait_impl.rs:54-58): the mapping goes, the node stays in the graph."""
        idx = self._id_to_idx.pop(id, None)
        if idx is None:
            return False
        del self._idx_to_id[idx]
        self._map_dirty = True
        return True

    def len(self) -> int:
        return len(self._id_to_idx)

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def dimension(self) -> int:
        return self._dimension

    def metric(self) -> DistanceMetric:
        return self._metric

    def search(self, query, k):
        return self.search_with_quality(query, k, SearchQuality.Balanced)

    # ---- inherent search API (index/hnsw/index/search.rs, batch.rs)
    def search_with_quality(self, query, k, quality):
        q = _as_f32_2d(query, self._dimension, "Query")
        if quality.kind == nv.PERFECT:
            return self.search_brute_force(q[0], k)
        if self.len() <= 100 and self.enable_vector_storage and self._vectors_present and self._next_idx > 0:
            return self.search_brute_force(q[0], k)
        return self._graph_search(q, k, quality.ef_search(k))[0]

    def search_batch_parallel(self, queries, k, quality):
        """batch.rs:159-197 -- no brute-force short-circuits on this path."""
        qs = np.ascontiguousarray(queries, dtype=np.float32)
        if qs.ndim != 2:
            qs = np.stack([np.asarray(x, dtype=np.float32) for x in queries]) if len(queries) else np.zeros(
                (0, self._dimension), np.float32)
        for i in range(qs.shape[0]):
            if qs.shape[1] != self._dimension:
                raise DimensionMismatch(f"Query {i} dimension mismatch: expected {self._dimension}, got {qs.shape[1]}")
        if qs.shape[0] == 0:
            return []
        return self._graph_search(qs, k, quality.ef_search(k))

    def search_brute_force(self, query, k):
        q = _as_f32_2d(query, self._dimension, "Query")
        if not self.enable_vector_storage or not self._vectors_present or self._next_idx == 0:
            if self._next_idx == 0:
                return []
            return self._graph_search(q, k, SearchQuality.Accurate.ef_search(k))[0]  # search.rs:180-194
        return self._brute(q, k)[0]

    search_brute_force_buffered = search_brute_force
    brute_force_search_parallel = search_brute_force

    def search_brute_force_gpu(self, query, k):
        """search.rs:229-279 (cosine only in the reference's wgpu path); always available here."""
        return self.search_brute_force(query, k)

    def search_with_rerank(self, query, k, rerank_k):
        return self.search_with_rerank_quality(query, k, rerank_k, SearchQuality.Accurate)

    def search_with_rerank_quality(self, query, k, rerank_k, initial_quality):
        q = _as_f32_2d(query, self._dimension, "Query")
        if initial_quality.kind == nv.PERFECT:
            initial_quality = SearchQuality.Accurate
        cands = self.search_with_quality(q[0], rerank_k, initial_quality)
        if not cands:
            return []
        if not self._vectors_present:  # rerank finds no vectors after load() (search.rs:130-137)
            return []
        snap = self._ensure_snapshot()
        pairs = [(id, self._id_to_idx[id]) for id, _ in cands if id in self._id_to_idx]
        if not pairs:
            return []
        scores = snap.rerank_batch(q, np.array([[p[1] for p in pairs]], dtype=np.uint32))[0]
        out = [(p[0], float(s)) for p, s in zip(pairs, scores)]
        out = _sort_results(self._metric, out)
        return out[:k]

    def set_searching_mode(self):
        self._ensure_snapshot()

    # ---- persistence (constructors.rs:190-287)
    # native_mappings.bin / native_meta.bin are the reference's bincode 1.3 files (fixed-width little-endian
    # integers, usize as u64, a HashMap as its u64 length followed by the (key, value) pairs, bool as one byte):
    #   mappings = (HashMap<u64, usize> id_to_idx, HashMap<usize, u64> idx_to_id, usize next_idx)
    #   meta     = (usize dimension, u8 metric, bool enable_vector_storage)
    # so a directory written by either side opens on the other.
    def save(self, path):
        os.makedirs(path, exist_ok=True)
        snap = self._ensure_snapshot()
        snap.dump(path, "native_hnsw")
        with open(os.path.join(path, "native_mappings.bin"), "wb") as f:
            f.write(_bincode_mappings(self._id_to_idx, self._idx_to_id, self._next_idx))
        with open(os.path.join(path, "native_meta.bin"), "wb") as f:
            f.write(struct.pack("<QBB", self._dimension, int(self._metric), 1 if self.enable_vector_storage else 0))

    @classmethod
    def load(cls, path, dimension=None, metric=None):
        """`dimension` and `metric` are kept for API compatibility, as in the reference: both are read from
        native_meta.bin (constructors.rs:190-217)."""
        for name in ("native_meta.bin", "native_hnsw.vectors", "native_hnsw.graph", "native_mappings.bin"):
            if not os.path.exists(os.path.join(path, name)):
                raise FileNotFoundError(f"{name} not found in {path}")
        with open(os.path.join(path, "native_meta.bin"), "rb") as f:
            raw = f.read()
        if len(raw) < 10:
            raise OSError("native_meta.bin is truncated")
        dim_file, metric_u8, evs = struct.unpack_from("<QBB", raw)
        if metric_u8 > 4:
            raise OSError("Unknown distance metric")  # constructors.rs:211-216
        with open(os.path.join(path, "native_mappings.bin"), "rb") as f:
            id_to_idx, idx_to_id, next_idx = _parse_bincode_mappings(f.read())
        snap = DeviceSnapshot.from_reference_files(path, DistanceMetric(metric_u8))
        if len(snap) and snap.dim != dim_file:
            raise OSError(f"native_meta.bin says dimension {dim_file}, native_hnsw.vectors holds {snap.dim}")
        ix = cls(dim_file, DistanceMetric(metric_u8), enable_vector_storage=bool(evs))
        ix._id_to_idx, ix._idx_to_id, ix._next_idx = id_to_idx, idx_to_id, next_idx
        ix._snapshot = snap
        ix._vectors_present = False  # ShardedVectors stays empty after load (constructors.rs:240)
        return ix

    # ---- vacuum.rs
    def tombstone_count(self) -> int:
        return self._next_idx - len(self._id_to_idx)

    def tombstone_ratio(self) -> float:
        return 0.0 if self._next_idx == 0 else self.tombstone_count() / self._next_idx

    def needs_vacuum(self) -> bool:
        return self.tombstone_ratio() > 0.2

    def vacuum(self) -> int:
        """index/hnsw/index/vacuum.rs:110-190: rebuild from the live vectors only -- new graph with
        HnswParams::auto, `parallel_insert` (the bulk builder here), fresh dense node indices -- and return how many
        vectors it kept.  The reference walks its DashMap (unspecified order); here live vectors keep their relative
        insertion order.  The device snapshot is rebuilt on the next search."""
        if not self.enable_vector_storage:
            raise VacuumError("VectorStorageDisabled")
        if not self._vectors_present:
            return 0  # ShardedVectors is empty after load(): nothing is collected (vacuum.rs:116-126)
        live = sorted(self._idx_to_id.items())
        if not live:
            return 0
        self._staged = [self._staged[idx] for idx, _ in live]
        self._idx_to_id = {i: e for i, (_, e) in enumerate(live)}
        self._id_to_idx = {e: i for i, e in self._idx_to_id.items()}
        self._next_idx = len(live)
        self._params = HnswParams.auto(self._dimension)
        self._bulk = True
        self._dirty = True
        self._map_dirty = True
        return len(live)

    # ---- internals
    def _materialise_staged(self):
        """Before inserting into an index restored from files: bring its vectors back from the device snapshot
        (`veles_index_get_rows`; the graph's own vectors survive a load, native/backend_adapter.rs:286-307, even though
        ShardedVectors does not, constructors.rs:240).  The next search rebuilds the graph over old + new vectors with
        the builder the insert path selects.  Where this differs from the reference: its NativeHnsw restarts the level
        PRNG from the seed after file_load (backend_adapter.rs:373) and keeps the loaded adjacency; here levels follow
        the PRNG in node-id order across the whole collection, as for an index that was never saved."""
        if self._snapshot is not None and not self._staged and self._next_idx > 0:
            n = len(self._snapshot)
            dt = {"f32": np.float32, "f16": np.float16}.get(self._store_dtype)
            if dt is None:
                raise NotImplementedError("re-staging packed-bit snapshots is not supported")
            rows = self._snapshot.get_rows(0, n, dt).astype(np.float32)
            self._staged = [rows[i] for i in range(n)]

    def _ensure_snapshot(self) -> DeviceSnapshot:
        if self._snapshot is None or self._dirty:
            vecs = np.stack(self._staged) if self._staged else np.zeros((0, self._dimension), np.float32)
            snap = DeviceSnapshot.from_vectors(vecs, self._metric, self._store_dtype)
            if not self._bulk and self._store_dtype == "f32" and len(vecs) <= self.EXACT_BUILD_LIMIT:
                # sequential inserts: the reference's deterministic graph (graph.rs:158-237), id for id
                snap.build_graph_exact(self._params.max_connections, self._params.ef_construction)
            else:
                snap.build_graph(self._params.max_connections, min(self._params.ef_construction, 4096))
            self._snapshot, self._dirty = snap, False
            self._map_dirty = True
        return self._snapshot

    def _map(self, ids, vals, cnt, transform):
        lib = nv.lib()
        out = []
        for r in range(ids.shape[0]):
            row = []
            for j in range(int(cnt[r])):
                ext = self._idx_to_id.get(int(ids[r, j]))
                if ext is None:  # tombstoned: dropped silently (search.rs:86-91)
                    continue
                v = float(vals[r, j])
                row.append((ext, float(lib.veles_transform_score(int(self._metric), v)) if transform else v))
            out.append(row)
        return out

    def _sync_mappings(self, snap):
        """Uploads node -> external id and the live bitmap when they changed (insert / remove / new snapshot)."""
        if not self._map_dirty:
            return
        n = len(snap)
        ext = np.zeros(n, np.uint64)
        live = np.zeros((n + 31) // 32, np.uint32)
        for idx, e in self._idx_to_id.items():
            if idx < n:
                ext[idx] = e
                live[idx >> 5] |= np.uint32(1 << (idx & 31))
        snap.set_id_map(ext, live)
        self._map_dirty = False

    def _graph_search(self, q, k, ef):
        if self._next_idx == 0:
            return [[] for _ in range(q.shape[0])]
        snap = self._ensure_snapshot()
        self._sync_mappings(snap)
        ids, sc, cnt = snap.search_batch_mapped(q, k, ef)   # id map, tombstone drop, transform_score on the device
        return [[(int(ids[r, j]), float(sc[r, j])) for j in range(int(cnt[r]))] for r in range(q.shape[0])]

    def _brute(self, q, k):
        snap = self._ensure_snapshot()
        # tombstoned nodes are filtered before the sort in the reference (search.rs:205-208): over-fetch
        kk = min(len(snap), k + self.tombstone_count())
        if kk == 0:
            return [[] for _ in range(q.shape[0])]
        ids, sc = snap.bruteforce_batch(q, kk)
        cnt = (ids != nv.INVALID_ID).sum(axis=1)
        return [row[:k] for row in self._map(ids, sc, cnt, False)]


def _sort_results(metric, results):
    """DistanceMetric::sort_results (core/distance.rs:95-103): stable, total order on the score."""
    def key(x):
        b = np.float32(x[1]).view(np.int32).item()
        b ^= ((b >> 31) & 0x7FFFFFFF)
        return b
    return sorted(results, key=key, reverse=DistanceMetric(metric).higher_is_better())
