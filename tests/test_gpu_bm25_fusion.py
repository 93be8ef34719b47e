"""GPU parity for BM25 posting scan (index/bm25.rs:269-376), hybrid RRF (collection/search/text.rs:133-180)
and FusionStrategy (fusion/strategy.rs:138-300) against the CPU oracle.  Scores are compared bit for bit
(the bar in BASELINE.md is 1e-5 relative)."""
import numpy as np
import pytest

from oracle import oracle as vo
from tests.gpu_util import bits_equal
from velesdb_b200 import (Bm25Index, Bm25Snapshot, DeviceSnapshot, DimensionMismatch, DistanceMetric, FusionStrategy, hybrid_search,
                          hybrid_search_batch, rrf_hybrid_batch)

pytestmark = pytest.mark.gpu


def zipf_corpus(n_docs, vocab, seed, s=1.07, lo=8, hi=120):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, vocab + 1) ** s
    p /= p.sum()
    lens = np.clip(np.exp(rng.normal(np.log(40), 0.5, n_docs)).astype(int), lo, hi)
    return [rng.choice(vocab, size=l, p=p).astype(np.uint32) for l in lens], p


def build_both(docs, k1=1.2, b=0.75):
    o = vo.Bm25(k1, b)
    tf_lists = {}
    total = 0
    for d, terms in enumerate(docs):
        o.add_document_terms(d, terms)
        total += len(terms)
        u, c = np.unique(terms, return_counts=True)
        for t, f in zip(u, c):
            tf_lists.setdefault(int(t), []).append((d, int(f)))
    n_terms = max(tf_lists) + 1
    term_ptr = np.zeros(n_terms + 1, np.uint64)
    pd, pt, df = [], [], np.zeros(n_terms, np.uint32)
    for t in range(n_terms):
        l = tf_lists.get(t, [])
        df[t] = len(l)
        pd += [x[0] for x in l]
        pt += [x[1] for x in l]
        term_ptr[t + 1] = len(pd)
    doc_len = np.array([len(t) for t in docs], np.uint32)
    snap = Bm25Snapshot(term_ptr, np.array(pd, np.uint32), np.array(pt, np.uint32), df, doc_len, len(docs), total, k1, b)
    return o, snap


@pytest.mark.parametrize("n_docs,vocab", [(300, 50), (20000, 2000)])
def test_bm25_batch_bit_exact(n_docs, vocab):
    docs, p = zipf_corpus(n_docs, vocab, seed=n_docs)
    o, snap = build_both(docs)
    rng = np.random.default_rng(1)
    nq, k = 64, 20
    q_ptr, q_terms = [0], []
    for i in range(nq):
        l = int(rng.integers(1, 7))
        t = rng.choice(vocab, size=l, p=p).astype(np.uint32)
        if i % 7 == 0:
            t = np.concatenate([t, t[:1]])          # duplicate query term: scored again (bm25_tests.rs:270-281)
        if i % 11 == 0:
            t = np.concatenate([t, [0xFFFFFFFF]])   # term missing from the dictionary
        q_terms += t.tolist()
        q_ptr.append(len(q_terms))
    docs_g, sc_g, cnt_g = snap.search_batch(q_ptr, q_terms, k)
    oi, os_, oc = o.search_batch_terms(q_ptr, q_terms, k, threads=8)
    assert np.array_equal(cnt_g, oc)
    for i in range(nq):
        c = int(cnt_g[i])
        assert np.array_equal(docs_g[i, :c], oi[i, :c].astype(np.uint32)), i
        assert bits_equal(sc_g[i, :c], os_[i, :c]), i
        assert (docs_g[i, c:] == 0xFFFFFFFF).all()
    assert cnt_g.max() == k


def test_bm25_reference_known_answers_on_gpu():
    # bm25_tests.rs:109-122, 160-179, 217-236, 270-281 through the host mirror
    ix = Bm25Index()
    ix.add_document(1, "rust programming language")
    ix.add_document(2, "python programming language")
    ix.add_document(3, "rust is fast")
    r = ix.search("rust", 10)
    assert sorted(x[0] for x in r) == [1, 3]
    assert Bm25Index().search("rust", 10) == []
    ix = Bm25Index()
    for i in range(1, 101):
        ix.add_document(i, f"document number {i} about rust")
    assert len(ix.search("rust", 5)) == 5
    ix = Bm25Index()
    ix.add_document(1, "rust")
    ix.add_document(2, "rust is a systems programming language that runs blazingly fast")
    r = ix.search("rust", 10)
    assert len(r) == 2 and r[0][0] == 1
    ix = Bm25Index()
    ix.add_document(1, "rust programming")
    assert len(ix.search("rust rust rust", 10)) == 1
    ix.add_document(1, "updated text")
    assert len(ix) == 1 and ix.search("rust", 10) == []       # stale posting scores 0 and is filtered
    assert ix.remove_document(1) and not ix.remove_document(1)
    # same answers as the oracle, including the replaced-document df quirk
    o, g = vo.Bm25(), Bm25Index()
    for d, t in [(1, "alpha beta gamma"), (2, "beta beta delta"), (1, "gamma delta delta epsilon"), (3, "alpha")]:
        o.add_document(d, t)
        g.add_document(d, t)
    for q in ("alpha", "beta delta", "gamma gamma epsilon", "zeta"):
        oi, os_ = o.search(q, 5)
        r = g.search(q, 5)
        assert [x[0] for x in r] == oi.tolist() and bits_equal([x[1] for x in r], os_), q


def test_rrf_hybrid_batch_bit_exact():
    rng = np.random.default_rng(5)
    nq, in_k, k = 200, 20, 10
    vi = np.stack([rng.choice(60, in_k, replace=False) for _ in range(nq)]).astype(np.uint32)
    ti = np.stack([rng.choice(60, in_k, replace=False) for _ in range(nq)]).astype(np.uint32)
    vc = rng.integers(0, in_k + 1, nq).astype(np.uint32)
    tc = rng.integers(0, in_k + 1, nq).astype(np.uint32)
    for w in (0.5, 0.3, 1.0, 0.0, 7.0):
        ids, sc, cnt = rrf_hybrid_batch(vi, vc, ti, tc, k, w)
        for q in range(nq):
            oi, os_ = vo.rrf_hybrid(vi[q, :vc[q]], ti[q, :tc[q]], k, w)
            assert cnt[q] == len(oi)
            assert np.array_equal(ids[q, :cnt[q]], oi.astype(np.uint32)), (w, q)
            assert bits_equal(sc[q, :cnt[q]], os_)


def sample_results():
    return [[(1, 0.95), (2, 0.85), (3, 0.75), (4, 0.65)], [(2, 0.90), (1, 0.80), (5, 0.70), (3, 0.60)],
            [(1, 0.92), (3, 0.82), (2, 0.72), (6, 0.62)]]


def test_fusion_strategies_bit_exact():
    # fusion/strategy_tests.rs:54-272 + random lists with in-list duplicates
    rng = np.random.default_rng(2)
    cases = [sample_results(), [[(1, 0.9), (2, 0.8)], [], [(1, 0.85), (3, 0.75)]], [[(1, 0.95), (2, 0.85), (3, 0.75)]]]
    for _ in range(5):
        cases.append([[(int(rng.integers(0, 40)), float(np.float32(rng.random()))) for _ in range(int(rng.integers(0, 50)))]
                      for _ in range(int(rng.integers(1, 10)))])
    strategies = [(FusionStrategy.Average(), (vo.AVERAGE, {})), (FusionStrategy.Maximum(), (vo.MAXIMUM, {})),
                  (FusionStrategy.RRF(60), (vo.RRF, {"rrf_k": 60})), (FusionStrategy.RRF(1), (vo.RRF, {"rrf_k": 1})),
                  (FusionStrategy.weighted(0.6, 0.3, 0.1), (vo.WEIGHTED, {"avg_w": 0.6, "max_w": 0.3, "hit_w": 0.1}))]
    for lists in cases:
        for st, (kind, kw) in strategies:
            got = st.fuse(lists)
            oi, os_ = vo.fuse(kind, lists, **kw)
            assert [g[0] for g in got] == oi.tolist(), (kind, lists)
            assert bits_equal([g[1] for g in got], os_)
    assert FusionStrategy.RRF().fuse([]) == [] and FusionStrategy.Average().fuse([[], []]) == []
    d1 = dict(FusionStrategy.RRF(60).fuse(sample_results()))[1]
    assert d1 > 0.04


def test_fusion_of_ten_lists_of_five_hundred_disjoint_hits():
    """multi_query_search allows 10 vectors x overfetch_k(top_k) = 500 hits each (collection/search/batch.rs:238,
    270-275): 5000 distinct documents in one request must fuse, not overflow."""
    rng = np.random.default_rng(9)
    lists = [[(l * 500 + i, float(np.float32(1.0 - i / 600.0))) for i in range(500)] for l in range(10)]
    lists[3] += [(7, 0.31), (1200, 0.5)]  # a few cross-list repeats
    for st, (kind, kw) in ((FusionStrategy.RRF(60), (vo.RRF, {"rrf_k": 60})), (FusionStrategy.Average(), (vo.AVERAGE, {})),
                           (FusionStrategy.weighted(0.6, 0.3, 0.1), (vo.WEIGHTED, {"avg_w": 0.6, "max_w": 0.3, "hit_w": 0.1}))):
        got = st.fuse(lists)
        oi, os_ = vo.fuse(kind, lists, **kw)
        assert len(got) == 5000 and [g[0] for g in got] == oi.tolist()
        assert bits_equal([g[1] for g in got], os_)


def test_hybrid_search_end_to_end():
    # Collection::hybrid_search (text.rs:113-203): vector top-2k + BM25 top-2k -> RRF -> top-k
    from velesdb_b200 import DistanceMetric, HnswIndex
    from tests.gpu_util import latent_data

    n, dim, k = 400, 32, 5
    x = latent_data(n, dim, seed=8)
    words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta"]
    rng = np.random.default_rng(3)
    hx, tx, ob = HnswIndex(dim, DistanceMetric.Cosine), Bm25Index(), vo.Bm25()
    for i in range(n):
        hx.insert(i, x[i])
        text = " ".join(rng.choice(words, size=int(rng.integers(2, 8))))
        tx.add_document(i, text)
        ob.add_document(i, text)
    q = x[7] + 0.05
    got = hybrid_search(hx, tx, q, "alpha gamma", k, 0.5)
    vres = hx.search(q, 2 * k)
    ti, _ = ob.search("alpha gamma", 2 * k)
    oi, os_ = vo.rrf_hybrid([r[0] for r in vres], ti, k, 0.5)
    assert [g[0] for g in got] == oi.tolist() and bits_equal([g[1] for g in got], os_)


def test_multi_query_search_and_search_with_filter_end_to_end():
    # Collection::multi_query_search (collection/search/batch.rs:238-330) and Collection::search_with_filter
    # (collection/search/vector.rs:164-239), minus the storage fetch
    from velesdb_b200 import DistanceMetric, HnswIndex, SearchQuality, multi_query_search, overfetch_k, search_with_filter
    from tests.gpu_util import latent_data

    assert [overfetch_k(k) for k in (1, 10, 11, 50, 51, 100, 101)] == [20, 200, 110, 500, 255, 500, 202]
    n, dim, k = 600, 32, 5
    x = latent_data(n, dim, seed=4)
    hx = HnswIndex(dim, DistanceMetric.Cosine)
    for i in range(n):
        hx.insert(i, x[i])
    qs = [x[3] + 0.02, x[77] - 0.03, x[400] + 0.01]
    even = lambda i: i % 2 == 0
    for st, (kind, kw) in ((FusionStrategy.RRF(60), (vo.RRF, {"rrf_k": 60})), (FusionStrategy.Average(), (vo.AVERAGE, {}))):
        for pred in (None, even):
            got = multi_query_search(hx, qs, k, st, pred)
            lists = hx.search_batch_parallel(qs, overfetch_k(k), SearchQuality.Balanced)
            if pred:
                lists = [[h for h in l if pred(h[0])] for l in lists]
            oi, os_ = vo.fuse(kind, lists, **kw)
            assert [g[0] for g in got] == oi[:k].tolist() and bits_equal([g[1] for g in got], os_[:k])
            assert pred is None or all(even(g[0]) for g in got)
    with pytest.raises(ValueError, match="at least one vector"):
        multi_query_search(hx, [], k, FusionStrategy.RRF())
    with pytest.raises(ValueError, match="at most 10 vectors"):
        multi_query_search(hx, [x[0]] * 11, k, FusionStrategy.RRF())
    # filtered search: first k matches among max(4k, k + 10) candidates, best first
    got = search_with_filter(hx, qs[0], k, even)
    cand = [h for h in hx.search(qs[0], max(4 * k, k + 10)) if even(h[0])][:k]
    assert got == sorted(cand, key=lambda h: -h[1]) and len(got) == k and all(even(h[0]) for h in got)
    assert search_with_filter(hx, qs[0], k, lambda i: False) == []


@pytest.mark.parametrize("k", [20, 50, 100])
def test_bm25_work_items_split_and_merge_bit_exact(k, monkeypatch):
    """bm25_sub_kernel (one warp per item) and bm25_flat_kernel (one CTA per item) cut a query into (query, part of the
    doc-id space) work items and the last one to finish merges the parts: any split -- one item per query, two, one per
    sub-range (what a single query gets) -- and round 1's walk kernel return the oracle's documents and score bits.
    30000 docs = 30 sub-ranges = 5 ranges; the frequent terms need several passes per sub-range / rounds per range."""
    docs, p = zipf_corpus(30000, 3000, seed=5)
    o, snap = build_both(docs)
    rng = np.random.default_rng(2)
    for nq in (1, 5):
        q_ptr, q_terms = [0], []
        for _ in range(nq):
            q_terms += rng.choice(3000, size=int(rng.integers(1, 7)), p=p).astype(np.uint32).tolist()
            q_ptr.append(len(q_terms))
        oi, os_, oc = o.search_batch_terms(q_ptr, q_terms, k, threads=4)
        for env in ({}, {"VELES_BM25_PARTS": "1"}, {"VELES_BM25_PARTS": "2"}, {"VELES_BM25_PARTS": "30"}, {"VELES_BM25_FLAT": "1"},
                    {"VELES_BM25_FLAT": "1", "VELES_BM25_PARTS": "1"}, {"VELES_BM25_FLAT": "1", "VELES_BM25_PARTS": "2"},
                    {"VELES_BM25_FLAT": "1", "VELES_BM25_FLAT_OCC": "5"}, {"VELES_BM25_WALK": "1"}):
            for key in ("VELES_BM25_PARTS", "VELES_BM25_FLAT", "VELES_BM25_FLAT_OCC", "VELES_BM25_WALK"):
                monkeypatch.delenv(key, raising=False)
            for key, val in env.items():
                monkeypatch.setenv(key, val)
            for rep in range(2):  # the per-query tickets must be back at zero for the second call
                docs_g, sc_g, cnt_g = snap.search_batch(q_ptr, q_terms, k)
                assert np.array_equal(cnt_g, oc), (env, rep)
                for i in range(nq):
                    c = int(cnt_g[i])
                    assert np.array_equal(docs_g[i, :c], oi[i, :c].astype(np.uint32)), (env, rep, i)
                    assert bits_equal(sc_g[i, :c], os_[i, :c]), (env, rep, i)


def test_bm25_snapshot_without_the_fine_skip_table(monkeypatch):
    """A snapshot built without the fine skip table (what a vocabulary too large for it gets) answers through
    bm25_flat_kernel by default: same documents and score bits as the oracle and as the snapshot with the table."""
    docs, p = zipf_corpus(20000, 1500, seed=9)
    monkeypatch.setenv("VELES_BM25_NO_FINE_TABLE", "1")
    o, snap_coarse = build_both(docs)
    monkeypatch.delenv("VELES_BM25_NO_FINE_TABLE")
    _, snap_fine = build_both(docs)
    rng = np.random.default_rng(3)
    nq, k = 33, 10
    q_ptr, q_terms = [0], []
    for _ in range(nq):
        q_terms += rng.choice(1500, size=int(rng.integers(1, 7)), p=p).astype(np.uint32).tolist()
        q_ptr.append(len(q_terms))
    oi, os_, oc = o.search_batch_terms(q_ptr, q_terms, k, threads=4)
    for snap in (snap_coarse, snap_fine):
        docs_g, sc_g, cnt_g = snap.search_batch(q_ptr, q_terms, k)
        assert np.array_equal(cnt_g, oc)
        for i in range(nq):
            c = int(cnt_g[i])
            assert np.array_equal(docs_g[i, :c], oi[i, :c].astype(np.uint32)), i
            assert bits_equal(sc_g[i, :c], os_[i, :c]), i


def test_hybrid_search_batch_one_call_equals_the_three_calls_and_the_oracle():
    """veles_hybrid_search_batch = Collection::hybrid_search (text.rs:113-203) for a batch: vector leg and text leg
    concurrently on the device, RRF over the device-resident lists.  Must equal, bit for bit, veles_search_batch +
    veles_bm25_search_batch + veles_rrf_hybrid called in sequence, and the oracle's search / Bm25 / rrf_hybrid."""
    rng = np.random.default_rng(21)
    n, dim, vocab = 3000, 64, 400
    z = rng.normal(size=(n, 8)).astype(np.float32) @ rng.normal(size=(8, dim)).astype(np.float32)
    x = (z + 0.3 * rng.normal(size=(n, dim))).astype(np.float32)
    g = vo.Hnsw(vo.COSINE, dim, M=16, ef_construction=100)
    g.insert_many(x)
    snap = DeviceSnapshot.from_arrays(x, DistanceMetric.Cosine, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    docs, p = zipf_corpus(n, vocab, seed=3)                      # document d describes vector d
    o, bm = build_both(docs)
    for nq, k, ef in ((1, 5, 32), (37, 10, 64), (300, 10, 128), (64, 40, 200)):
        q = (x[rng.integers(0, n, nq)] + 0.2 * rng.normal(size=(nq, dim))).astype(np.float32)
        q_ptr, q_terms = [0], []
        for i in range(nq):
            t = rng.choice(vocab, size=int(rng.integers(1, 7)), p=p).astype(np.uint32)
            if i % 5 == 0:
                t = np.array([0xFFFFFFFF], np.uint32)            # no known term: the text leg returns nothing
            if i % 9 == 0:
                t = np.concatenate([t, t[:1]])
            q_terms += t.tolist()
            q_ptr.append(len(q_terms))
        q_ptr, q_terms = np.array(q_ptr, np.uint32), np.array(q_terms, np.uint32)
        vi, vd, vc = snap.search_batch(q, 2 * k, ef)
        td, ts, tc = bm.search_batch(q_ptr, q_terms, 2 * k)
        for w in (None, 0.3, 1.0, 0.0):
            ids, sc, cnt = hybrid_search_batch(snap, bm, q, q_ptr, q_terms, k, ef, w)
            ri, rs, rc = rrf_hybrid_batch(vi, vc, td, tc, k, 0.5 if w is None else w)
            assert np.array_equal(cnt, rc), (nq, k, w)
            for i in range(nq):
                assert np.array_equal(ids[i, :cnt[i]], ri[i, :rc[i]]), (nq, k, w, i)
                assert bits_equal(sc[i, :cnt[i]], rs[i, :rc[i]])
        # against the oracle, end to end (canonical tie order on both sides)
        oi, od, oc, _ = g.search_batch(q, 2 * k, ef, order="canonical")
        ids, sc, cnt = hybrid_search_batch(snap, bm, q, q_ptr, q_terms, k, ef, 0.5)
        bi, _, bc = o.search_batch_terms(q_ptr, q_terms, 2 * k, threads=8)
        for i in range(nq):
            fi, fs = vo.rrf_hybrid(oi[i, :oc[i]].astype(np.uint32), bi[i, :bc[i]].astype(np.uint32), k, 0.5)
            assert cnt[i] == len(fi) and np.array_equal(ids[i, :cnt[i]], fi.astype(np.uint32)), (nq, k, i)
            assert bits_equal(sc[i, :cnt[i]], fs)
    with pytest.raises(DimensionMismatch):
        hybrid_search_batch(snap, bm, np.zeros((2, dim + 1), np.float32), np.zeros(3, np.uint32), np.zeros(0, np.uint32), 5, 32)
    # an empty text index: the fusion sees the vector list only (bm25.rs:275-278)
    empty = Bm25Snapshot(np.zeros(2, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(1, np.uint32),
                         np.zeros(1, np.uint32), 0, 0)
    q = x[:4].copy()
    ids, sc, cnt = hybrid_search_batch(snap, empty, q, np.array([0, 1, 2, 3, 4], np.uint32), np.zeros(4, np.uint32), 5, 64, 0.5)
    vi, vd, vc = snap.search_batch(q, 10, 64)
    ri, rs, rc = rrf_hybrid_batch(vi, vc, np.full((4, 10), 0xFFFFFFFF, np.uint32), np.zeros(4, np.uint32), 5, 0.5)
    assert np.array_equal(cnt, rc) and np.array_equal(ids, ri) and bits_equal(sc, rs)


def test_hybrid_search_reference_known_answers_on_gpu():
    """collection/tests.rs:298-336, 394-446 (the reference's own answers for Collection::hybrid_search) through the host
    mirror (HnswIndex + Bm25Index + hybrid_search) and through the one-call batch form on the same snapshots."""
    from tests.test_oracle_known_answers import HYBRID_KNOWN_ANSWERS
    from velesdb_b200 import HnswIndex, SearchQuality
    from velesdb_b200 import _native as nv
    from velesdb_b200.bm25 import tokenize

    for points, query, text, k, w, first in HYBRID_KNOWN_ANSWERS:
        hx, tx = HnswIndex(3, DistanceMetric.Cosine), Bm25Index()
        for pid, vec, txt in points:
            hx.insert(pid, np.asarray(vec, np.float32))
            tx.add_document(pid, txt)
        got = hybrid_search(hx, tx, np.asarray(query, np.float32), text, k, w)
        assert got and got[0][0] == first, (got, first)
        assert all(got[i][1] >= got[i + 1][1] for i in range(len(got) - 1))
        # the batch form: node i and BM25 slot i both stand for points[i] (ids ascending), so ids map back by position
        snap, bsnap = hx._ensure_snapshot(), tx._snapshot()
        terms = [tx._term(t, False) for t in tokenize(text)]
        q_terms = np.array([nv.INVALID_ID if t is None else t for t in terms], np.uint32)
        ef = SearchQuality.Balanced.ef_search(2 * k)
        ids, sc, cnt = hybrid_search_batch(snap, bsnap, np.asarray([query], np.float32), np.array([0, len(q_terms)], np.uint32),
                                           q_terms, k, ef, w)
        assert cnt[0] == len(got) and [points[int(i)][0] for i in ids[0, :cnt[0]]] == [g[0] for g in got]
        assert bits_equal(sc[0, :cnt[0]], [g[1] for g in got])
