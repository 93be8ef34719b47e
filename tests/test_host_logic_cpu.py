"""CPU tests of the host mirror's own logic (velesdb_b200/index.py): id <-> node mapping, tombstones,
quality -> ef routing, score transform, error behaviour.  The device snapshot is replaced by a stand-in
that answers from the CPU oracle (test infrastructure), so no GPU is needed."""
import numpy as np
import pytest

from oracle import oracle as vo
from velesdb_b200 import DimensionMismatch, DistanceMetric, HnswIndex, SearchQuality
from velesdb_b200 import _native as nv


class OracleSnapshot:
    """Duck-types DeviceSnapshot for HnswIndex: answers search / brute force from the oracle."""

    def __init__(self, metric, vectors, M=16, ef_c=100):
        self.metric_, self.x = metric, np.ascontiguousarray(vectors, np.float32)
        self.g = vo.Hnsw(int(metric), self.x.shape[1], M=M, ef_construction=ef_c)
        if len(self.x):
            self.g.insert_many(self.x)
        self.calls = []
        self.ext, self.live = None, None

    def __len__(self):
        return len(self.x)

    def search_batch(self, q, k, ef):
        self.calls.append(("search", k, ef))
        ids, d, cnt, _ = self.g.search_batch(q, k, ef, order="canonical")
        return ids.astype(np.uint32), d, cnt

    def set_id_map(self, ext_ids=None, live_bits=None):
        self.ext, self.live = ext_ids, live_bits

    def search_batch_mapped(self, q, k, ef, k_fetch=None, allow_bits=None, stream=None):
        """veles_search_batch_mapped restated on the host: live filter, id map, transform_score, first k."""
        ids, d, cnt = self.search_batch(q, k if k_fetch is None else k_fetch, ef)
        out_i = np.full((len(q), k), 0xFFFFFFFFFFFFFFFF, np.uint64)
        out_s = np.full((len(q), k), np.nan, np.float32)
        out_c = np.zeros(len(q), np.uint32)
        for r in range(len(q)):
            w = 0
            for j in range(int(cnt[r])):
                node = int(ids[r, j])
                if self.live is not None and not (int(self.live[node >> 5]) >> (node & 31)) & 1:
                    continue
                if w < k:
                    out_i[r, w] = node if self.ext is None else self.ext[node]
                    out_s[r, w] = vo.transform_score(int(self.metric_), float(d[r, j]))
                w += 1
            out_c[r] = min(w, k)
        return out_i, out_s, out_c

    def bruteforce_batch(self, q, k):
        self.calls.append(("brute", k))
        ids, sc = vo.bruteforce_batch(int(self.metric_), self.x, q, k)
        return ids.astype(np.uint32), sc

    def rerank_batch(self, q, cand):
        return np.array([[vo.metric_value(int(self.metric_), q[0], self.x[c]) for c in cand[0]]], np.float32)


def make_index(n=300, dim=16, metric=DistanceMetric.Cosine, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.normal(size=(n, dim)).astype(np.float32)
    ix = HnswIndex(dim, metric)
    ids = [1000 + 7 * i for i in range(n)]  # external ids differ from node indices
    for e, v in zip(ids, x):
        ix.insert(e, v)
    ix.insert(ids[0], x[1])  # duplicate id: skipped
    snap = OracleSnapshot(metric, x)
    ix._snapshot, ix._dirty = snap, False
    return ix, snap, x, ids


def test_len_dimension_metric_and_duplicates():
    ix, _, _, ids = make_index()
    assert ix.len() == 300 and ix.dimension() == 16 and ix.metric() == DistanceMetric.Cosine and not ix.is_empty()
    assert ix.tombstone_count() == 0


def test_quality_routes_to_ef_and_transform_score():
    ix, snap, x, ids = make_index()
    res = ix.search(x[5], 10)                      # Balanced -> ef = max(128, 4k)
    assert snap.calls[-1] == ("search", 10, 128)
    assert res[0][0] == ids[5] and 0.0 <= res[0][1] <= 1.0   # cosine scores clamp into [0, 1]
    ix.search_with_quality(x[5], 40, SearchQuality.Fast)
    assert snap.calls[-1] == ("search", 40, 80)
    ix.search_with_quality(x[5], 10, SearchQuality.Custom(77))
    assert snap.calls[-1] == ("search", 10, 77)
    ix.search_with_quality(x[5], 10, SearchQuality.Perfect)    # Perfect short-circuits to brute force
    assert snap.calls[-1][0] == "brute"
    ix.search_batch_parallel(x[:4], 10, SearchQuality.Accurate)
    assert snap.calls[-1] == ("search", 10, 512)


def test_small_index_uses_brute_force_only_on_the_single_query_path():
    ix, snap, x, ids = make_index(n=50)
    ix.search(x[0], 5)
    assert snap.calls[-1][0] == "brute"            # len <= 100 (search.rs:75-77)
    ix.search_batch_parallel(x[:2], 5, SearchQuality.Balanced)
    assert snap.calls[-1][0] == "search"           # no short-circuit in batch (batch.rs:159-197)


def test_soft_delete_filters_results_everywhere():
    ix, snap, x, ids = make_index()
    assert ix.remove(ids[5]) and not ix.remove(ids[5])
    assert ix.len() == 299 and ix.tombstone_count() == 1
    assert all(r[0] != ids[5] for r in ix.search(x[5], 10))
    bf = ix.search_brute_force(x[5], 10)
    assert len(bf) == 10 and all(r[0] != ids[5] for r in bf)   # over-fetch keeps k results after the filter
    assert ix.tombstone_ratio() == pytest.approx(1 / 300)


def test_dimension_mismatch_is_an_error_like_the_reference_panic():
    ix, _, x, _ = make_index()
    with pytest.raises(DimensionMismatch, match="Query dimension mismatch: expected 16, got 3"):
        ix.search(np.zeros(3, np.float32), 5)
    with pytest.raises(DimensionMismatch, match="Vector dimension mismatch"):
        ix.insert(99999, np.zeros(5, np.float32))
    with pytest.raises(DimensionMismatch):
        ix.search_batch_parallel(np.zeros((2, 4), np.float32), 5, SearchQuality.Balanced)


def test_euclidean_and_dot_score_transforms():
    for metric, sign in ((DistanceMetric.Euclidean, 1), (DistanceMetric.DotProduct, -1)):
        ix, snap, x, ids = make_index(metric=metric)
        res = ix.search(x[3], 5)
        raw = snap.g.search(x[3], 5, 128, order="canonical")[1]
        assert [r[1] for r in res] == [sign * float(d) for d in raw]


def test_rerank_sorts_by_metric_value():
    ix, snap, x, ids = make_index()
    rr = ix.search_with_rerank(x[9], 5, 30)
    assert len(rr) == 5 and rr[0][0] == ids[9]
    assert all(rr[i][1] >= rr[i + 1][1] for i in range(4))


def test_empty_index_returns_nothing():
    ix = HnswIndex(8, DistanceMetric.Cosine)
    assert ix.search_brute_force(np.zeros(8, np.float32), 3) == []
    assert ix.search_batch_parallel(np.zeros((0, 8), np.float32), 3, SearchQuality.Balanced) == []
    assert nv.lib().veles_ef_search(nv.BALANCED, 10, 0) == 128


def test_search_with_filter_overfetch_and_order():
    # collection/search/vector.rs:164-239: candidates_k = max(4k, k + 10), Balanced quality, first k matches
    from velesdb_b200 import overfetch_k, search_with_filter
    ix, snap, x, ids = make_index()
    odd = lambda e: ((e - 1000) // 7) % 2 == 1
    got = search_with_filter(ix, x[5], 3, odd)
    assert snap.calls[-1] == ("search", 13, 128)            # max(12, 13) candidates; ef = max(128, 4 * 13)
    assert len(got) == 3 and all(odd(e) for e, _ in got)
    assert [s for _, s in got] == sorted((s for _, s in got), reverse=True)
    search_with_filter(ix, x[5], 40, odd)
    assert snap.calls[-1] == ("search", 160, 640)
    assert [overfetch_k(k) for k in (10, 50, 100, 200)] == [200, 500, 500, 400]   # collection/search/batch.rs:270-275


def test_vacuum_and_sequential_batch_bookkeeping():
    # vacuum.rs:45-190, batch.rs:120-139: host-side state only (the snapshot is rebuilt lazily on the next search)
    from velesdb_b200 import VacuumError
    ix, snap, x, ids = make_index(n=120)
    for e in ids[:30]:
        assert ix.remove(e)
    assert not ix.remove(ids[0])
    assert ix.len() == 90 and ix.tombstone_count() == 30 and ix.needs_vacuum()
    assert ix.vacuum() == 90
    assert ix.len() == 90 and ix.tombstone_count() == 0 and not ix.needs_vacuum() and ix._dirty and ix._bulk
    assert sorted(ix._id_to_idx) == sorted(ids[30:]) and sorted(ix._idx_to_id) == list(range(90))
    assert all(np.array_equal(ix._staged[ix._id_to_idx[e]], x[i]) for i, e in enumerate(ids) if i >= 30)
    ix2 = HnswIndex(16, DistanceMetric.Cosine, enable_vector_storage=False)
    with pytest.raises(VacuumError, match="VectorStorageDisabled"):
        ix2.vacuum()
    ix3 = HnswIndex(16, DistanceMetric.Cosine)
    assert ix3.vacuum() == 0
    assert ix3.insert_batch_sequential([(1, x[0]), (2, x[1]), (1, x[2])]) == 2 and not ix3._bulk


def test_mappings_and_meta_files_follow_the_reference_bincode_layout():
    """constructors.rs:190-287 writes native_mappings.bin / native_meta.bin with bincode 1.3 (Cargo.lock): fixed-width
    little-endian integers, usize as u64, HashMap = u64 length + (key, value) pairs, bool = one byte.  The bytes below
    are written by hand from that spec; the mirror must parse them and write the same layout back."""
    import struct

    from velesdb_b200.index import _bincode_mappings, _parse_bincode_mappings

    raw = (struct.pack("<Q", 2) + struct.pack("<QQ", 1000, 0) + struct.pack("<QQ", 2 ** 40 + 7, 2) +
           struct.pack("<Q", 2) + struct.pack("<QQ", 0, 1000) + struct.pack("<QQ", 2, 2 ** 40 + 7) + struct.pack("<Q", 3))
    a, b, nxt = _parse_bincode_mappings(raw)
    assert a == {1000: 0, 2 ** 40 + 7: 2} and b == {0: 1000, 2: 2 ** 40 + 7} and nxt == 3
    assert _bincode_mappings(a, b, nxt) == raw
    import pytest

    with pytest.raises(OSError):
        _parse_bincode_mappings(raw[:-3])
    with pytest.raises(OSError):
        _parse_bincode_mappings(struct.pack("<Q", 2 ** 60) + raw[8:])   # absurd length: rejected before allocating


def test_distance_metric_semantics_known_answers():
    # collection/search/distance_semantics_tests.rs:41-118 and core/distance.rs:76-103: which metrics are similarities,
    # and the order DistanceMetric::sort_results puts scores in (most similar first)
    from velesdb_b200.index import _sort_results

    assert DistanceMetric.Cosine.higher_is_better() and DistanceMetric.DotProduct.higher_is_better()
    assert DistanceMetric.Jaccard.higher_is_better()
    assert not DistanceMetric.Euclidean.higher_is_better() and not DistanceMetric.Hamming.higher_is_better()
    scores = [(0, 0.3), (1, 0.9), (2, 0.5), (3, 0.7)]
    assert [s for _, s in _sort_results(DistanceMetric.Cosine, scores)] == [0.9, 0.7, 0.5, 0.3]
    assert [s for _, s in _sort_results(DistanceMetric.Euclidean, scores)] == [0.3, 0.5, 0.7, 0.9]
    # equal scores keep their input order (stable sort), negative zero sorts below zero (total order)
    assert [i for i, _ in _sort_results(DistanceMetric.Euclidean, [(5, 1.0), (4, 1.0), (3, 0.0), (2, -0.0)])] == [2, 3, 5, 4]


def test_hnsw_params_presets_known_answers():
    # index/hnsw/params_tests.rs:6-150, 198-246: every preset's M / ef_construction / capacity / storage mode
    from velesdb_b200 import HnswParams as P

    def t(p):
        return (p.max_connections, p.ef_construction, p.max_elements, p.storage_mode)

    assert t(P.default())[:2] == (32, 400) and P.default().storage_mode == "full"
    assert t(P.auto(128))[:2] == (24, 300) and t(P.auto(1024))[:2] == (32, 400) and t(P.auto(256))[:2] == (24, 300)
    assert t(P.fast()) == (16, 150, 100_000, "full") and t(P.turbo()) == (12, 100, 100_000, "full")
    assert t(P.high_recall(768))[:2] == (40, 600) and t(P.fast_indexing(768))[:2] == (16, 200)
    assert t(P.large_dataset(768))[:3] == (128, 2000, 750_000)
    assert t(P.for_dataset_size(768, 5_000))[:3] == (32, 400, 20_000)
    assert t(P.for_dataset_size(768, 50_000))[:3] == (128, 1600, 150_000)
    assert t(P.for_dataset_size(768, 300_000))[:3] == (128, 2000, 750_000)
    assert t(P.million_scale(768))[:3] == (128, 1600, 1_500_000)
    assert t(P.for_dataset_size(768, 100_000))[:2] == (128, 1600) and t(P.for_dataset_size(768, 500_000))[:2] == (128, 2000)
    assert t(P.for_dataset_size(128, 10_000))[:3] == (24, 200, 20_000) and t(P.for_dataset_size(128, 10_001))[:3] == (64, 800, 150_000)
    assert t(P.for_dataset_size(128, 400_000))[:3] == (96, 1200, 750_000) and t(P.for_dataset_size(128, 2_000_000))[:3] == (64, 800, 1_500_000)
    assert t(P.max_recall(128))[:2] == (32, 500) and t(P.max_recall(512))[:2] == (48, 800) and t(P.max_recall(1024))[:2] == (64, 1000)
    assert t(P.custom(32, 400, 50_000)) == (32, 400, 50_000, "full")
    assert t(P.with_sq8(768)) == (32, 400, 100_000, "sq8") and t(P.with_binary(768))[0] == 32 and P.with_binary(768).storage_mode == "binary"
    assert P.custom(32, 400, 50_000) == P.custom(32, 400, 50_000) and P.fast() != P.turbo()
    # params_tests.rs:153-196: quality -> ef_search
    assert [SearchQuality.Fast.ef_search(10), SearchQuality.Balanced.ef_search(10), SearchQuality.Accurate.ef_search(10),
            SearchQuality.Custom(50).ef_search(10)] == [64, 128, 512, 50]
    assert [SearchQuality.Perfect.ef_search(k) for k in (10, 50, 100)] == [4096, 5000, 10000]
    assert [SearchQuality.Fast.ef_search(100), SearchQuality.Balanced.ef_search(50), SearchQuality.Accurate.ef_search(40)] == [200, 200, 640]


def test_constructor_variants_known_answers():
    # index_tests.rs:118-134, 153-204 and constructors.rs:29-118: new / new_turbo / new_fast_insert / with_params
    from velesdb_b200 import HnswParams, VacuumError

    ix = HnswIndex.new(768, DistanceMetric.Cosine)
    assert ix.is_empty() and ix.len() == 0 and ix.dimension() == 768 and ix.metric() == DistanceMetric.Cosine
    assert ix._params == HnswParams.auto(768) and ix.enable_vector_storage
    turbo = HnswIndex.new_turbo(64, DistanceMetric.Cosine)
    assert (turbo._params.max_connections, turbo._params.ef_construction) == (24, 450) and turbo.enable_vector_storage
    fast = HnswIndex.new_fast_insert(64, DistanceMetric.Cosine)
    assert fast._params == HnswParams.auto(64) and not fast.enable_vector_storage
    rows = np.array([[(i + j) * 0.01 for j in range(64)] for i in range(100)], np.float32)
    for i in range(100):
        fast.insert(i, rows[i])
        turbo.insert(i, rows[i])
    assert fast.len() == 100 and turbo.len() == 100
    with pytest.raises(VacuumError):                       # VacuumError::VectorStorageDisabled
        fast.vacuum()
    assert HnswIndex.new(64, DistanceMetric.Cosine).vacuum() == 0
    # searches: the fast-insert index has no stored vectors, so even a 100-vector index goes through the graph
    for ix2 in (fast, turbo):
        snap = OracleSnapshot(DistanceMetric.Cosine, rows)
        ix2._snapshot, ix2._dirty = snap, False
    q = np.array([j * 0.01 for j in range(64)], np.float32)
    assert len(fast.search(q, 10)) == 10 and fast._snapshot.calls[-1][0] == "search"
    assert len(turbo.search(q, 10)) == 10 and turbo._snapshot.calls[-1][0] == "brute"     # <= 100 vectors, stored
    assert fast.search_brute_force(q, 5) == fast.search_with_quality(q, 5, SearchQuality.Accurate)   # search.rs:180-194
    custom = HnswIndex.with_params(128, DistanceMetric.Euclidean, HnswParams.custom(48, 600, 1_000_000))
    assert (custom._params.max_connections, custom._params.ef_construction) == (48, 600)


def _attach(ix, rows):
    snap = OracleSnapshot(ix.metric(), np.asarray(rows, np.float32))
    ix._snapshot, ix._dirty = snap, False
    return snap


def test_batch_insert_rerank_and_batch_search_known_answers():
    # index_tests.rs:441-494, 556-571, 633-650, 1018-1068 through the mirror (device snapshot = oracle stand-in)
    pts = [(1, [1.0, 0.0, 0.0]), (2, [0.0, 1.0, 0.0]), (3, [0.0, 0.0, 1.0]), (4, [0.5, 0.5, 0.0]), (5, [0.5, 0.0, 0.5])]
    ix = HnswIndex.new(3, DistanceMetric.Cosine)
    assert ix.insert_batch_parallel(pts) == 5 and ix.len() == 5
    _attach(ix, [v for _, v in pts])
    ix.set_searching_mode()                                        # nothing left to build
    res = ix.search([1.0, 0.0, 0.0], 3)
    assert len(res) == 3 and res[0][0] == 1
    rr = ix.search_with_rerank([1.0, 0.0, 0.0], 3, 100)            # rerank_k > index size
    assert 0 < len(rr) <= 5 and rr[0][0] == 1
    dup = HnswIndex.new(3, DistanceMetric.Cosine)
    dup.insert(1, [1.0, 0.0, 0.0])
    assert dup.insert_batch_parallel([(1, [0.0, 1.0, 0.0]), (2, [0.0, 0.0, 1.0])]) == 1 and dup.len() == 2
    with pytest.raises(DimensionMismatch, match="Vector dimension mismatch"):
        HnswIndex.new(3, DistanceMetric.Cosine).insert_batch_sequential([(1, [1.0, 0.0])])
    assert HnswIndex.new(3, DistanceMetric.Cosine).insert_batch_sequential([]) == 0
    # batch search = the individual searches, query by query (here even id for id: both paths end in the same graph
    # search once the index holds more than 100 vectors)
    big = HnswIndex.new(64, DistanceMetric.Cosine)
    rows = np.array([[np.sin((i + j) * 0.01) for j in range(64)] for i in range(150)], np.float32)
    for i in range(150):
        big.insert(i, rows[i])
    _attach(big, rows)
    qs = np.array([[np.sin((200 + i + j) * 0.01) for j in range(64)] for i in range(10)], np.float32)
    batch = big.search_batch_parallel(qs, 5, SearchQuality.Balanced)
    single = [big.search_with_quality(q, 5, SearchQuality.Balanced) for q in qs]
    assert len(batch) == 10 and all(len(b) == 5 for b in batch) and batch == single
    one = HnswIndex.new(3, DistanceMetric.Cosine)
    one.insert(1, [1.0, 0.0, 0.0])
    assert one.search_batch_parallel([], 5, SearchQuality.Fast) == []
    # tombstones (index_tests.rs:22-54)
    t = HnswIndex.new(64, DistanceMetric.Cosine)
    assert t.tombstone_count() == 0 and t.tombstone_ratio() == 0.0
    for i in range(10):
        t.insert(i, rows[i])
    for i in range(3):
        assert t.remove(i)
    assert t.tombstone_count() == 3 and t.len() == 7 and abs(t.tombstone_ratio() - 0.3) < 1e-9
