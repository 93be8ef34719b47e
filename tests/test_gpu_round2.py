"""Round-2 parity and API tests on the GPU: the re-rank kernel against the oracle, the overflow flag of the
asynchronous entry points, the pipelined submit/wait ABI, and searches racing on one index."""
import threading

import numpy as np
import pytest

from oracle import oracle as vo
from tests.gpu_util import bits_equal, build_oracle, latent_data, queries_near
from velesdb_b200 import DeviceSnapshot, DistanceMetric
from velesdb_b200 import _native as nv

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric", [vo.COSINE, vo.EUCLIDEAN, vo.DOT, vo.HAMMING, vo.JACCARD])
@pytest.mark.parametrize("dim,store", [(768, "f32"), (100, "f32"), (13, "f32"), (96, "f16")])
def test_rerank_values_equal_the_oracle(metric, dim, store):
    """veles_rerank_batch = the exact-metric step of HnswIndex::search_with_rerank (index/hnsw/index/search.rs:118-160):
    compute_distance (search.rs:30-38) of the query against each candidate.  Values must have the oracle's bits."""
    rng = np.random.default_rng(5)
    n, nq, m = 300, 7, 24
    if metric in (vo.HAMMING, vo.JACCARD):
        x = (rng.random((n, dim)) > 0.5).astype(np.float32)
        q = (rng.random((nq, dim)) > 0.5).astype(np.float32)
    else:
        x = latent_data(n, dim, seed=2)
        q = queries_near(x, nq, seed=3)
    snap = DeviceSnapshot.from_vectors(x, metric, store_dtype=store)
    cand = rng.integers(0, n, size=(nq, m)).astype(np.uint32)
    cand[1, 3] = nv.INVALID_ID
    cand[4, :] = nv.INVALID_ID
    got = snap.rerank_batch(q, cand)
    xs = x.astype(np.float16).astype(np.float32) if store == "f16" else x
    want = np.full((nq, m), np.nan, np.float32)
    for i in range(nq):
        for j in range(m):
            if cand[i, j] != nv.INVALID_ID:
                want[i, j] = vo.metric_value(metric, q[i], xs[cand[i, j]])
    assert bits_equal(got[~np.isnan(want)], want[~np.isnan(want)])
    assert np.isnan(got[np.isnan(want)]).all()


TIE_EF = 5000


def _tie_index():
    """A graph on which one query's tie list outgrows its 4096 entries.  6000 identical `far` vectors in a 64-ary tree
    laid out in BFS order (so the beam fills its TIE_EF results with far nodes 0..4999, all at distance 1), and 5000
    identical `near` vectors (= the query, distance 0) in a second tree hanging off far node 100.  Every near node
    that enters the results evicts an unexpanded far node whose distance still equals the worst result's: by
    graph.rs:474 it stays poppable, so it goes to the tie list -- ~4900 of them."""
    F, N, dim, M = 6000, 5000, 16, 32
    x = np.zeros((F + N, dim), np.float32)
    x[:F, 0] = 1.0
    rows, cols = [0], []
    for i in range(F):
        if i == 100:
            cols.extend(range(F, F + 64))
        else:
            cols.extend(c for c in range(64 * i + 1, 64 * i + 65) if c < F)
        rows.append(len(cols))
    for j in range(N):
        cols.extend(F + c for c in range(64 * j + 1, 64 * j + 65) if c < N)
        rows.append(len(cols))
    layers = [(np.array(rows, np.uint64), np.array(cols, np.uint32))]
    return DeviceSnapshot.from_arrays(x, DistanceMetric.Euclidean, layers, M, 2 * M, 0, 0), x[F:F + 4].copy()


def test_tie_overflow_is_reported_by_host_and_device_entry_points():
    import torch

    snap, q = _tie_index()
    ids, dist, cnt = snap.search_batch(q, 5, 64)   # a small beam never gets there
    assert (dist == 1.0).all()
    with pytest.raises(nv.VelesError) as e:
        snap.search_batch(q, 5, TIE_EF)
    assert e.value.status == nv.ERR_OVERFLOW
    dev = torch.device("cuda", 0)
    q_d = torch.from_numpy(q).to(dev)
    ids = torch.empty((4, 5), dtype=torch.int32, device=dev)
    dist = torch.empty((4, 5), dtype=torch.float32, device=dev)
    cnt = torch.empty(4, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    snap.search_batch_device(q_d, 5, TIE_EF, ids, dist, cnt, None, st.cuda_stream)   # enqueues only: no error yet
    with pytest.raises(nv.VelesError) as e:
        snap.search_status(st.cuda_stream)
    assert e.value.status == nv.ERR_OVERFLOW
    snap.search_status(st.cuda_stream)  # reported once, then clear
    # the pipelined ABI reports it at wait
    out = [np.empty((4, 5), np.uint32), np.empty((4, 5), np.float32), np.empty(4, np.uint32)]
    t = snap.search_submit(q, 5, TIE_EF, *out)
    with pytest.raises(nv.VelesError) as e:
        snap.search_wait(t)
    assert e.value.status == nv.ERR_OVERFLOW


def test_submit_wait_equals_search_batch_with_tickets_in_flight():
    x = latent_data(5000, 64, seed=21)
    g = build_oracle(vo.COSINE, x, M=16, ef_c=100)
    snap = DeviceSnapshot.from_arrays(x, DistanceMetric.Cosine, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    batches = [queries_near(x, 300, seed=s) for s in range(6)]
    want = [snap.search_batch(b, 10, 64) for b in batches]
    outs = [(np.empty((300, 10), np.uint32), np.empty((300, 10), np.float32), np.empty(300, np.uint32)) for _ in batches]
    tickets = []
    for b, o in zip(batches, outs):          # three in flight at most
        tickets.append(snap.search_submit(b, 10, 64, *o))
        if len(tickets) >= 3:
            snap.search_wait(tickets.pop(0))
    for t in tickets:
        snap.search_wait(t)
    for w, o in zip(want, outs):
        assert np.array_equal(w[0], o[0]) and bits_equal(w[1], o[1]) and np.array_equal(w[2], o[2])
    with pytest.raises(nv.VelesError):
        snap.search_wait(12345)              # unknown ticket


def test_concurrent_searches_on_one_index_do_not_share_scratch():
    """HnswIndex is Send + Sync and searches take a read lock (index/hnsw/index/search.rs:59-94): host threads and
    streams must be able to search one snapshot at the same time."""
    import torch

    x = latent_data(8000, 48, seed=33)
    g = build_oracle(vo.COSINE, x, M=16, ef_c=100)
    snap = DeviceSnapshot.from_arrays(x, DistanceMetric.Cosine, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    qs = [queries_near(x, 700, seed=100 + s) for s in range(4)]
    want = [snap.search_batch(q, 10, 64) for q in qs]
    # (a) host threads, each with its own stream
    got = [None] * 4

    def work(i):
        st = torch.cuda.Stream()
        for _ in range(5):
            got[i] = snap.search_batch(qs[i], 10, 64, stream=st.cuda_stream)

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for w, o in zip(want, got):
        assert np.array_equal(w[0], o[0]) and bits_equal(w[1], o[1])
    # (b) device-pointer calls on four streams, all enqueued before any completes
    dev = torch.device("cuda", 0)
    streams = [torch.cuda.Stream() for _ in range(4)]
    q_d = [torch.from_numpy(q).to(dev) for q in qs]
    ids = [torch.empty((700, 10), dtype=torch.int32, device=dev) for _ in range(4)]
    dist = [torch.empty((700, 10), dtype=torch.float32, device=dev) for _ in range(4)]
    cnt = [torch.empty(700, dtype=torch.int32, device=dev) for _ in range(4)]
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(4):
            snap.search_batch_device(q_d[i], 10, 64, ids[i], dist[i], cnt[i], None, streams[i].cuda_stream)
    for i in range(4):
        snap.search_status(streams[i].cuda_stream)
        assert np.array_equal(want[i][0], ids[i].cpu().numpy().astype(np.uint32))
        assert bits_equal(want[i][1], dist[i].cpu().numpy())


def test_index_directory_round_trip_in_the_reference_layout(tmp_path):
    """HnswIndex::save / load (constructors.rs:190-287): graph + vectors in file format v1, mappings and meta as
    bincode.  After load the searches answer as before and ShardedVectors is empty (rerank finds nothing)."""
    import struct

    from velesdb_b200 import HnswIndex, SearchQuality

    x = latent_data(300, 24, seed=6)
    ix = HnswIndex(24, DistanceMetric.Euclidean)
    for i in range(300):
        ix.insert(5_000_000_000 + i, x[i])
    ix.remove(5_000_000_007)
    want = ix.search_batch_parallel(x[:20], 5, SearchQuality.Balanced)
    ix.save(str(tmp_path))
    meta = open(tmp_path / "native_meta.bin", "rb").read()
    assert meta == struct.pack("<QBB", 24, 1, 1)
    assert len(open(tmp_path / "native_mappings.bin", "rb").read()) == 8 + 299 * 16 + 8 + 299 * 16 + 8
    back = HnswIndex.load(str(tmp_path), 999, DistanceMetric.Cosine)   # both arguments are ignored, as in the reference
    assert back.dimension() == 24 and back.metric() == DistanceMetric.Euclidean and back.len() == 299
    assert back.search_batch_parallel(x[:20], 5, SearchQuality.Balanced) == want
    assert back.search_with_rerank(x[3], 3, 10) == []


@pytest.mark.parametrize("metric,store,n,dim,nq,k", [(vo.COSINE, "f32", 20000, 768, 300, 10), (vo.EUCLIDEAN, "f32", 5000, 100, 130, 5),
                                                    (vo.DOT, "f16", 9000, 64, 64, 20), (vo.COSINE, "f32", 300, 48, 7, 10),
                                                    (vo.COSINE, "f32", 6000, 64, 40, 40)])
def test_tensor_core_relaxed_brute_force(metric, store, n, dim, nq, k, monkeypatch):
    """veles_bruteforce_batch_relaxed: candidates from the tcgen05 fp16 GEMM, exact re-rank.  Recall-gated against the
    exact path (fp16 rounding may only move rows near rank k * oversample); every returned score must be the exact
    metric value of its row, bit for bit, and the order must be the exact path's order."""
    x = latent_data(n, dim, latent=12, noise=0.3, seed=41, normalize=(metric == vo.COSINE))
    q = queries_near(x, nq, jitter=0.2, seed=42)
    snap = DeviceSnapshot.from_vectors(x, metric, store_dtype=store)
    ei, es = snap.bruteforce_batch(q, k)
    ri, rs = snap.bruteforce_batch_relaxed(q, k, oversample=4)
    rec = np.mean([len(set(ei[i].tolist()) & set(ri[i].tolist())) / k for i in range(nq)])
    assert rec >= 0.99, rec
    same = ei == ri
    assert bits_equal(es[same], rs[same])                 # same row => same exact score
    exact_of = snap.rerank_batch(q, ri)                     # every score is the exact metric value of its row
    assert bits_equal(exact_of, rs)
    sign = -1.0 if metric in (vo.COSINE, vo.DOT) else 1.0
    assert (np.diff(sign * rs.astype(np.float64), axis=1) >= 0).all()
    # re-ranking EVERY candidate that passed the bound (VELES_TC_RERANK_ALL: CTA-per-query bound kernel + flattened
    # finish kernel) against the first versions of those two kernels: same bound, same candidates, same keys ->
    # identical output; and it can only agree with the exact path at least as often as the default, which re-ranks the
    # 2 * k * oversample best candidates by GEMM score
    monkeypatch.setenv("VELES_TC_RERANK_ALL", "1")
    ai, as_ = snap.bruteforce_batch_relaxed(q, k, oversample=4)
    monkeypatch.delenv("VELES_TC_RERANK_ALL")
    monkeypatch.setenv("VELES_TC_OLD_TAIL", "1")
    oi, os_ = snap.bruteforce_batch_relaxed(q, k, oversample=4)
    monkeypatch.delenv("VELES_TC_OLD_TAIL")
    assert np.array_equal(ai, oi) and bits_equal(as_, os_)
    rec_all = np.mean([len(set(ei[i].tolist()) & set(ai[i].tolist())) / k for i in range(nq)])
    assert rec_all >= rec - 1e-9, (rec_all, rec)
    # the CTA-wide selection (bound from per-thread minima, collect, rank) against the register-list scans it replaces
    # in the bound kernel and in the finish kernel: the same k-th key and the same kr keys -> identical output
    monkeypatch.setenv("VELES_TC_REGTOPK", "1")
    gi, gs = snap.bruteforce_batch_relaxed(q, k, oversample=4)
    monkeypatch.delenv("VELES_TC_REGTOPK")
    assert np.array_equal(ri, gi) and bits_equal(rs, gs)


@pytest.mark.parametrize("metric,dim,nq,k", [(vo.COSINE, 64, 40, 10), (vo.EUCLIDEAN, 96, 9, 5), (vo.DOT, 128, 33, 20),
                                            (vo.HAMMING, 64, 12, 10), (vo.COSINE, 64, 3, 10), (vo.EUCLIDEAN, 64, 8, 100),
                                            # rows beyond 64 MB, dim % 128 == 0: the tile kernel's dynamic chunk-major items
                                            (vo.COSINE, 256, 70, 10), (vo.EUCLIDEAN, 256, 33, 5)])
def test_fused_brute_force_equals_the_matrix_path_and_the_oracle(metric, dim, nq, k, monkeypatch):
    """The exact scan without the [nq, n] score matrix: <= 8 queries select inside the scan kernel, larger batches go
    sample -> bound -> filtered scan -> select.  Both must return exactly what the matrix path and the oracle do."""
    n = 70_000
    rng = np.random.default_rng(77)
    if metric == vo.HAMMING:
        x = (rng.random((n, dim)) > 0.5).astype(np.float32)
        q = (rng.random((nq, dim)) > 0.5).astype(np.float32)
    else:
        x = latent_data(n, dim, latent=10, noise=0.3, seed=5)
        q = queries_near(x, nq, jitter=0.2, seed=6)
    snap = DeviceSnapshot.from_vectors(x, metric)
    fi, fs = snap.bruteforce_batch(q, k)
    monkeypatch.setenv("VELES_BF_NO_FUSE", "1")
    mi, ms = snap.bruteforce_batch(q, k)
    monkeypatch.delenv("VELES_BF_NO_FUSE")
    assert np.array_equal(fi, mi) and bits_equal(fs, ms)
    if dim % 128 == 0:
        monkeypatch.setenv("VELES_BF_STATIC_TILES", "1")   # the static (tile, query group) mapping of the same kernel
        si, ss = snap.bruteforce_batch(q, k)
        monkeypatch.delenv("VELES_BF_STATIC_TILES")
        assert np.array_equal(fi, si) and bits_equal(fs, ss)
    pick = np.unique(np.linspace(0, nq - 1, 6).astype(int))  # queries of every 32-query group
    oi, os_ = vo.bruteforce_batch(metric, x, q[pick], k, threads=8)
    assert np.array_equal(fi[pick], oi.astype(np.uint32)) and bits_equal(fs[pick], os_)


def test_insert_after_load_and_sparse_bm25_ids(tmp_path):
    """(a) HnswIndex::load then insert (constructors.rs:190-253 + trait_impl.rs:10-36): the vectors come back from the
    device snapshot and the index keeps answering for old and new ids.  (b) BM25 documents keyed by arbitrary u32 ids
    (bm25.rs:134-140 only requires that they fit u32) cost one slot each."""
    from velesdb_b200 import Bm25Index, HnswIndex

    x = latent_data(260, 16, seed=3)
    ix = HnswIndex(16, DistanceMetric.Cosine)
    for i in range(200):
        ix.insert(1000 + i, x[i])
    ix.save(str(tmp_path))
    back = HnswIndex.load(str(tmp_path))
    for i in range(200, 260):
        back.insert(1000 + i, x[i])
    assert back.len() == 260
    assert back.search(x[7], 1)[0][0] == 1007 and back.search(x[250], 1)[0][0] == 1250
    bm, ob = Bm25Index(), vo.Bm25()
    for d, text in ((4_000_000_000, "alpha beta beta"), (7, "beta gamma"), (123_456_789, "alpha alpha gamma delta")):
        bm.add_document(d, text)
        ob.add_document(d, text)
    got = bm.search("alpha beta", 3)
    oi, os_ = ob.search("alpha beta", 3)
    assert [g[0] for g in got] == oi.tolist() and bits_equal([g[1] for g in got], os_)


def test_append_links_new_vectors_into_the_live_graph():
    """veles_index_append = HnswIndex::insert_batch_parallel on an index that already has a graph (batch.rs:82-108):
    no rebuild -- old adjacency stays, new nodes are linked in, degree bounds hold, recall stays at the level of a graph
    built at once, and the search on the grown graph still equals the oracle on the exported graph."""
    from velesdb_b200 import HnswIndex, SearchQuality

    n0, n1, dim, M = 6000, 3000, 48, 16
    x = latent_data(n0 + n1, dim, latent=10, noise=0.3, seed=91, normalize=True)
    snap = DeviceSnapshot.from_vectors(x[:n0], DistanceMetric.Cosine)
    snap.build_graph(M, 100)
    before = snap.export_graph()
    snap.append(x[n0:], 100)
    assert len(snap) == n0 + n1
    layers = snap.export_graph()
    deg = np.diff(layers[0][0].astype(np.int64))
    assert deg.max() <= 2 * M and (deg >= 1).all()
    lv = vo.levels(M, n0 + n1)
    assert snap.max_layer == int(lv.max()) and snap.entry_point == int(np.argmax(lv == lv.max()))
    # old rows only changed by reverse links: whatever left a row was replaced by a closer new node
    rp0, c0 = before[0]
    rp1, c1 = layers[0]
    changed = sum(1 for i in range(0, n0, 37) if not np.array_equal(c0[rp0[i]:rp0[i + 1]], c1[rp1[i]:rp1[i + 1]]))
    assert changed > 0
    q = queries_near(x, 300, jitter=0.05, seed=5)
    ids, dist, cnt, st = snap.search_batch(q, 10, 64, with_stats=True)
    bi, _ = snap.bruteforce_batch(q, 10)
    rec = np.mean([len(set(ids[i].tolist()) & set(bi[i].tolist())) / 10 for i in range(len(q))])
    assert rec >= 0.95, rec
    g = vo.Hnsw.from_arrays(vo.COSINE, x, layers, M, 2 * M, snap.entry_point, snap.max_layer)
    oi, od, oc, ost = g.search_batch(q, 10, 64, order="canonical", threads=8)
    keep = ost[:, 4] == 0
    assert np.array_equal(ids[keep], oi[keep].astype(np.uint32)) and bits_equal(dist, od)
    # the host mirror: a bulk insert into a live index goes through the append, ids stay mapped
    ix = HnswIndex(dim, DistanceMetric.Cosine)
    ix.insert_batch_parallel([(10_000 + i, x[i]) for i in range(n0)])
    assert ix.search(x[5], 1)[0][0] == 10_005
    added = ix.insert_batch_parallel([(10_000 + i, x[i]) for i in range(n0 - 50, n0 + n1)])   # 50 duplicates skipped
    assert added == n1 and ix.len() == n0 + n1
    res = ix.search_batch_parallel(x[[3, n0 + 7, n0 + n1 - 1]], 1, SearchQuality.Balanced)
    assert [r[0][0] for r in res] == [10_003, 10_000 + n0 + 7, 10_000 + n0 + n1 - 1]


def test_gather_window_of_a_single_rank_equals_the_plain_search():
    """veles_search_batch_gather_d with world = 1 (the degenerate box): the search writes straight into the gather
    window, two epochs alternate between its two buffers, and the window equals veles_search_batch_d's outputs."""
    import ctypes as C

    import torch

    x = latent_data(4000, 32, seed=8)
    g = build_oracle(vo.COSINE, x, M=16, ef_c=100)
    snap = DeviceSnapshot.from_arrays(x, DistanceMetric.Cosine, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    lib = nv.lib()
    nq, k, ef = 200, 10, 64
    blob = np.zeros(lib.veles_comm_handle_bytes(), np.uint8)
    h = C.c_void_p()
    nv.check(lib.veles_comm_create(0, 1, nq, k, C.byref(h), nv.ptr(blob)))
    dev = torch.device("cuda", 0)
    try:
        for seed in (1, 2, 3):
            q = queries_near(x, nq, seed=seed)
            want = snap.search_batch(q, k, ef)
            q_d = torch.from_numpy(q).to(dev)
            nv.check(lib.veles_search_batch_gather_d(snap.h, h, nv.ptr(q_d), nq, k, ef, None, None))
            nv.check(lib.veles_comm_status(h, None))
            pi, pd, pc = C.c_void_p(), C.c_void_p(), C.c_void_p()
            nv.check(lib.veles_comm_window(h, C.byref(pi), C.byref(pd), C.byref(pc)))
            from velesdb_b200.dist import _device_view

            ids = _device_view(pi.value, (nq, k), torch.int32, dev).cpu().numpy().astype(np.uint32)
            dd = _device_view(pd.value, (nq, k), torch.float32, dev).cpu().numpy()
            cnt = _device_view(pc.value, (nq,), torch.int32, dev).cpu().numpy().astype(np.uint32)
            assert np.array_equal(ids, want[0]) and bits_equal(dd, want[1]) and np.array_equal(cnt, want[2])
        with pytest.raises(nv.VelesError):
            lib_call = lib.veles_search_batch_gather_d(snap.h, h, nv.ptr(q_d), nq - 1, k, ef, None, None)
            nv.check(lib_call)   # the window was created for nq queries
    finally:
        lib.veles_comm_destroy(h)


@pytest.mark.parametrize("n,nq,k", [(5000, 1, 10), (5000, 2, 7), (9, 1, 10), (3000, 1, 32)])
def test_fused_brute_force_small_collection_tail_and_tied_scores(n, nq, k, monkeypatch):
    """Small collections select in the scan kernel's last CTA from staged keys under a bound taken from the CTA lists'
    heads (tail 2a).  (a) random rows: ids and score bits equal the oracle's and the general tail's; (b) a collection of
    identical rows -- every score ties, the candidates overflow the k * k table -- falls back to the general tail and
    still returns the k smallest ids."""
    rng = np.random.default_rng(n + k)
    x = latent_data(n, 64, latent=8, noise=0.3, seed=n)
    q = queries_near(x, nq, jitter=0.2, seed=3)
    for metric in (vo.COSINE, vo.EUCLIDEAN):
        snap = DeviceSnapshot.from_vectors(x, metric)
        fi, fs = snap.bruteforce_batch(q, k)
        oi, os_ = vo.bruteforce_batch(metric, x, q, k, threads=4)
        kk = min(k, n)
        assert np.array_equal(fi[:, :kk], oi.astype(np.uint32)[:, :kk]) and bits_equal(fs[:, :kk], os_[:, :kk])
        assert (fi[:, kk:] == 0xFFFFFFFF).all()
        monkeypatch.setenv("VELES_BF_NO_STAGED_TAIL", "1")
        gi, gs = snap.bruteforce_batch(q, k)
        monkeypatch.delenv("VELES_BF_NO_STAGED_TAIL")
        assert np.array_equal(fi, gi) and bits_equal(fs[:, :kk], gs[:, :kk])
    same = np.tile(rng.normal(size=(1, 64)).astype(np.float32), (n, 1))
    snap = DeviceSnapshot.from_vectors(same, vo.EUCLIDEAN)
    ti, ts = snap.bruteforce_batch(same[:nq] + 0.5, k)
    kk = min(k, n)
    assert np.array_equal(ti[:, :kk], np.tile(np.arange(kk, dtype=np.uint32), (nq, 1)))
    assert (ts[:, :kk].view(np.uint32) == ts[0, 0].view(np.uint32)).all()
