"""Host mirror of ``DualPrecisionHnsw`` (velesdb-core native/dual_precision.rs:60-441) over the C ABI.

Same names, argument meaning and defaults as the reference:

* ``DualPrecisionHnsw.new(metric, dimension, max_connections, ef_construction, max_elements)`` (:87-103)
* ``insert(vector) -> node id`` (:122-143): the quantizer trains itself on the first
  ``min(1000, max_elements)`` vectors, ``force_train_quantizer()`` (:172-176) trains on what is there
* ``search(query, k, ef_search)`` (:179-228) -- f32 traversal + exact re-rank, i.e. ``NativeHnsw::search``
* ``search_with_config(query, k, ef_search, config)`` (:263-325) -- int8 traversal + exact re-rank
* ``DualPrecisionConfig`` defaults 4 / True / 10_000 (:33-57)

The reference mutates its graph per insert; here vectors are staged and the device snapshot (graph by the
exact sequential builder, SQ8 store by ``veles_index_attach_sq8``) is rebuilt on the next search, as
``HnswIndex`` in index.py does.  No CPU path: every search is a device call.
"""
from __future__ import annotations

import dataclasses

import numpy as np

from .index import DeviceSnapshot, DimensionMismatch, DistanceMetric


@dataclasses.dataclass
class DualPrecisionConfig:
    """dual_precision.rs:33-57"""
    oversampling_ratio: int = 4
    use_int8_traversal: bool = True
    min_index_size: int = 10_000


class DualPrecisionHnsw:
    def __init__(self, metric, dimension, max_connections, ef_construction, max_elements):
        self.metric = DistanceMetric(metric)
        self.dimension = int(dimension)
        self.max_connections = int(max_connections)
        self.ef_construction = int(ef_construction)
        self.training_sample_size = min(1000, int(max_elements))  # dual_precision.rs:100
        self._vectors = []
        self._train_count = 0       # 0 = quantizer not trained
        self._snapshot = None
        self._dirty = False
        self._external = False      # True: wraps a caller's snapshot, read-only

    new = classmethod(lambda cls, *a, **kw: cls(*a, **kw))

    @classmethod
    def from_snapshot(cls, snapshot: DeviceSnapshot, max_connections=32, train_count=None):
        """Wraps an existing device snapshot (graph already present) as if its vectors had been inserted in id
        order: the quantizer is trained on the first min(1000, n) of them unless `train_count` says otherwise."""
        dp = cls(snapshot.metric, snapshot.dim, max_connections, 0, max(len(snapshot), 1))
        dp._snapshot = snapshot
        dp._external = True
        n = len(snapshot)
        t = dp.training_sample_size if train_count is None else int(train_count)
        if n >= t > 0:
            dp._train_count = t
            snapshot.attach_sq8(t)
        return dp

    # ---- dual_precision.rs:106-119
    def len(self) -> int:
        return len(self._snapshot) if self._external else len(self._vectors)

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def is_quantizer_trained(self) -> bool:
        return self._train_count > 0

    # ---- dual_precision.rs:122-143
    def insert(self, vector) -> int:
        if self._external:
            raise RuntimeError("a DualPrecisionHnsw wrapped around a device snapshot is read-only")
        v = np.asarray(vector, dtype=np.float32).reshape(-1)
        if v.size != self.dimension:
            raise DimensionMismatch(f"Vector dimension mismatch: expected {self.dimension}, got {v.size}")
        node = len(self._vectors)
        self._vectors.append(v.copy())
        self._dirty = True
        if self._train_count == 0 and len(self._vectors) >= self.training_sample_size:
            self._train_count = len(self._vectors)
        return node

    def force_train_quantizer(self) -> None:
        if self._train_count == 0 and self._vectors:
            self._train_count = len(self._vectors)
            self._dirty = True

    def quantizer(self):
        """(min_vals, scales, inv_scales) of the trained ScalarQuantizer, or None (dual_precision.rs:230-234)."""
        if self._train_count == 0:
            return None
        mn, sc, inv, _ = self._ensure().sq8_export(with_codes=False)
        return mn, sc, inv

    # ---- searches
    def search(self, query, k, ef_search):
        """dual_precision.rs:179-228.  With a trained quantizer the reference still traverses in f32, asks for
        max(2 ef, 4 k) candidates (it gets at most ef), recomputes the same exact distances and stably re-sorts
        an already sorted list: the result is NativeHnsw::search(query, k, ef_search)."""
        q = self._query(query)
        if self.is_empty():
            return []
        ids, dist, cnt = self._ensure().search_batch(q, k, ef_search)
        return [(int(ids[0, j]), float(dist[0, j])) for j in range(int(cnt[0]))]

    def search_with_config(self, query, k, ef_search, config: DualPrecisionConfig | None = None):
        config = config or DualPrecisionConfig()
        if self._train_count == 0 or not config.use_int8_traversal or self.len() < config.min_index_size:
            return self.search(query, k, ef_search)  # dual_precision.rs:271-279 (inner.search)
        q = self._query(query)
        ids, dist, cnt = self._ensure().search_batch_sq8(q, k, ef_search, config.oversampling_ratio)
        return [(int(ids[0, j]), float(dist[0, j])) for j in range(int(cnt[0]))]

    def search_batch_with_config(self, queries, k, ef_search, config: DualPrecisionConfig | None = None):
        """The batched form the GPU path exists for: one launch for all queries."""
        config = config or DualPrecisionConfig()
        snap = self._ensure()
        if self._train_count == 0 or not config.use_int8_traversal or self.len() < config.min_index_size:
            return snap.search_batch(queries, k, ef_search)
        return snap.search_batch_sq8(queries, k, ef_search, config.oversampling_ratio)

    # ---- internals
    def _query(self, query):
        q = np.asarray(query, dtype=np.float32).reshape(1, -1)
        if q.shape[1] != self.dimension:
            raise DimensionMismatch(f"Query dimension mismatch: expected {self.dimension}, got {q.shape[1]}")
        return q

    def _ensure(self) -> DeviceSnapshot:
        if self._external:
            return self._snapshot
        if self._snapshot is None or self._dirty:
            vs = np.stack(self._vectors)
            snap = DeviceSnapshot.from_vectors(vs, self.metric)
            snap.build_graph_exact(self.max_connections, self.ef_construction)
            if self._train_count:
                snap.attach_sq8(self._train_count)
            self._snapshot = snap
            self._dirty = False
        return self._snapshot
