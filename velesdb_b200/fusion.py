"""Host-side mirror of ``FusionStrategy`` (crates/velesdb-core/src/fusion/strategy.rs) and of the RRF
inside ``Collection::hybrid_search`` (collection/search/text.rs:113-203) over the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nv


class FusionError(ValueError):
    pass


class FusionStrategy:
    AVERAGE, MAXIMUM, RRF_KIND, WEIGHTED = 0, 1, 2, 3

    def __init__(self, kind, k=60, avg_weight=0.0, max_weight=0.0, hit_weight=0.0):
        self.kind, self.k = kind, k
        self.avg_weight, self.max_weight, self.hit_weight = avg_weight, max_weight, hit_weight

    @staticmethod
    def Average():
        return FusionStrategy(FusionStrategy.AVERAGE)

    @staticmethod
    def Maximum():
        return FusionStrategy(FusionStrategy.MAXIMUM)

    @staticmethod
    def RRF(k=60):
        return FusionStrategy(FusionStrategy.RRF_KIND, k=k)

    @staticmethod
    def rrf_default():
        return FusionStrategy.RRF(60)

    @staticmethod
    def weighted(avg_weight, max_weight, hit_weight):
        """strategy.rs:88-122: weights must be non-negative and sum to 1 (+-0.001)."""
        if min(avg_weight, max_weight, hit_weight) < 0:
            raise FusionError("weights must be non-negative")
        if abs(avg_weight + max_weight + hit_weight - 1.0) > 0.001:
            raise FusionError("weights must sum to 1.0")
        return FusionStrategy(FusionStrategy.WEIGHTED, avg_weight=avg_weight, max_weight=max_weight,
                              hit_weight=hit_weight)

    def fuse(self, results):
        """results: [[(id, score), ...], ...] -> [(id, fused_score)] sorted score-descending."""
        nv.init()
        ptr = np.zeros(len(results) + 1, np.uint32)
        ids, sc = [], []
        for i, l in enumerate(results):
            for d, s in l:
                if not (0 <= d <= 0xFFFFFFFF):
                    raise FusionError("device fusion handles 32-bit ids")
                ids.append(d)
                sc.append(s)
            ptr[i + 1] = len(ids)
        cap = max(len(ids), 1)
        ia = np.array(ids or [0], np.uint32)
        sa = np.array(sc or [0], np.float32)
        oi = np.zeros(cap, np.uint32)
        os_ = np.zeros(cap, np.float32)
        cnt = C.c_uint32(0)
        nv.check(nv.lib().veles_fuse(self.kind, nv.ptr(ptr), len(results), nv.ptr(ia), nv.ptr(sa), self.k,
                                     self.avg_weight, self.max_weight, self.hit_weight, cap, nv.ptr(oi), nv.ptr(os_),
                                     C.byref(cnt), None))
        return [(int(oi[i]), float(os_[i])) for i in range(cnt.value)]


def rrf_hybrid_batch(vec_ids, vec_cnt, txt_ids, txt_cnt, k, vector_weight=0.5):
    """text.rs:133-180 for a batch: [nq, in_k] id lists with valid counts -> (ids, scores, counts)."""
    nv.init()
    vec_ids = np.ascontiguousarray(vec_ids, np.uint32)
    txt_ids = np.ascontiguousarray(txt_ids, np.uint32)
    nq, in_k = vec_ids.shape
    assert txt_ids.shape == (nq, in_k)
    vec_cnt = np.ascontiguousarray(vec_cnt, np.uint32)
    txt_cnt = np.ascontiguousarray(txt_cnt, np.uint32)
    oi = np.empty((nq, k), np.uint32)
    os_ = np.empty((nq, k), np.float32)
    oc = np.zeros(nq, np.uint32)
    nv.check(nv.lib().veles_rrf_hybrid(nv.ptr(vec_ids), nv.ptr(vec_cnt), nv.ptr(txt_ids), nv.ptr(txt_cnt), nq, in_k,
                                       float(vector_weight), k, nv.ptr(oi), nv.ptr(os_), nv.ptr(oc), None))
    return oi, os_, oc


def hybrid_search_batch(snapshot, text_snapshot, queries, q_ptr, q_terms, k, ef, vector_weight=None, stream=None):
    """``Collection::hybrid_search`` (text.rs:113-203, minus the storage fetch) for a batch in ONE device call
    (``veles_hybrid_search_batch``): vector top-2k on ``snapshot`` (a ``DeviceSnapshot``; ``ef`` as the caller's
    quality dictates, the reference uses ``ef_search(Balanced, 2k)``) and BM25 top-2k on ``text_snapshot`` (a
    ``Bm25Snapshot``) run concurrently, the RRF reads both lists on the device, one copy back.  ``q_ptr`` /
    ``q_terms``: per-query term-id lists as ``Bm25Snapshot.search_batch``.  Node ids are document ids.
    Returns ``(ids [nq, k] u32, scores [nq, k] f32, counts [nq] u32)``; identical to the three separate calls."""
    nv.init()
    w = 0.5 if vector_weight is None else vector_weight
    from .index import _as_f32_2d
    queries = _as_f32_2d(queries, snapshot.dim, "Query")                  # DimensionMismatch, text.rs:124-130
    q_ptr = np.ascontiguousarray(q_ptr, np.uint32)
    q_terms = np.ascontiguousarray(q_terms, np.uint32)
    nq = queries.shape[0]
    assert q_ptr.size == nq + 1
    qt = q_terms if q_terms.size else np.zeros(1, np.uint32)
    oi = np.empty((nq, k), np.uint32)
    os_ = np.empty((nq, k), np.float32)
    oc = np.zeros(nq, np.uint32)
    nv.check(nv.lib().veles_hybrid_search_batch(snapshot.h, text_snapshot.h, nv.ptr(queries), nv.ptr(q_ptr), nv.ptr(qt), nq, k,
                                                ef, float(w), nv.ptr(oi), nv.ptr(os_), nv.ptr(oc), stream))
    return oi, os_, oc


def hybrid_search(index, text_index, vector_query, text_query, k, vector_weight=None):
    """``Collection::hybrid_search`` (text.rs:113-203) minus the storage fetch: vector top-2k (Balanced) and
    BM25 top-2k, fused by RRF on the device.  Ids are the external ids of ``index`` (must fit u32)."""
    w = 0.5 if vector_weight is None else vector_weight
    vres = index.search(vector_query, k * 2)
    tres = text_index.search(text_query, k * 2)
    in_k = max(2 * k, 1)
    vi = np.full((1, in_k), nv.INVALID_ID, np.uint32)
    ti = np.full((1, in_k), nv.INVALID_ID, np.uint32)
    vi[0, :len(vres)] = [r[0] for r in vres]
    ti[0, :len(tres)] = [r[0] for r in tres]
    ids, sc, cnt = rrf_hybrid_batch(vi, [len(vres)], ti, [len(tres)], k, w)
    return [(int(ids[0, j]), float(sc[0, j])) for j in range(int(cnt[0]))]


def search_with_filter(index, query, k, predicate):
    """``Collection::search_with_filter`` (collection/search/vector.rs:164-239) minus the storage fetch:
    post-filtering over ``candidates_k = max(4k, k + 10)`` index results (``index.search`` = Balanced), first k
    matches in index order, then the reference's final sort by score (stable; descending for similarity
    metrics).  ``predicate(id) -> bool`` stands for ``filter.matches(payload)``."""
    candidates_k = max(k * 4, k + 10)
    hits = [(i, s) for i, s in index.search(query, candidates_k) if predicate(i)][:k]
    return sorted(hits, key=lambda h: -h[1] if index.metric().higher_is_better() else h[1])


def overfetch_k(top_k: int) -> int:
    """collection/search/batch.rs:270-275"""
    if top_k <= 10:
        return top_k * 20
    if top_k <= 50:
        return top_k * 10
    if top_k <= 100:
        return top_k * 5
    return top_k * 2


def multi_query_search(index, vectors, top_k, strategy: "FusionStrategy", predicate=None):
    """``Collection::multi_query_search`` (collection/search/batch.rs:238-330) minus the storage fetch: at most 10
    query vectors, one batched over-fetched search (Balanced), optional pre-fusion filter, ``FusionStrategy::fuse``
    on the device, first ``top_k``.  Ids are the external ids of ``index`` (must fit u32)."""
    from .index import SearchQuality
    if len(vectors) == 0:
        raise ValueError("multi_query_search requires at least one vector")      # batch.rs:241-245
    if len(vectors) > 10:
        raise ValueError(f"multi_query_search supports at most 10 vectors, got {len(vectors)}")  # batch.rs:247-253
    batch = index.search_batch_parallel(vectors, overfetch_k(top_k), SearchQuality.Balanced)
    if predicate is not None:
        batch = [[(i, s) for i, s in r if predicate(i)] for r in batch]
    fused = strategy.fuse(batch)
    return fused[:top_k]
