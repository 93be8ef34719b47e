"""GPU bulk graph builder (veles_index_build_graph): structural invariants that equal the
reference (levels, entry point, degree bounds), search parity *given the built graph*, and
recall against exact brute force."""
import numpy as np
import pytest

from oracle import oracle as vo
from tests.gpu_util import bits_equal, latent_data, queries_near
from velesdb_b200 import DeviceSnapshot, DistanceMetric

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("metric,dim,n,M", [(vo.COSINE, 64, 6000, 16), (vo.EUCLIDEAN, 100, 3000, 8),
                                           (vo.DOT, 64, 3000, 16)])
def test_built_graph_invariants_and_parity(metric, dim, n, M):
    x = latent_data(n, dim, latent=8, noise=0.3, seed=3, normalize=(metric != vo.EUCLIDEAN))
    snap = DeviceSnapshot.from_vectors(x, metric)
    snap.build_graph(M)
    # levels / entry point follow the reference PRNG in node-id order (graph.rs:368-403, 230-233)
    lv = vo.levels(M, n)
    assert snap.max_layer == int(lv.max())
    assert snap.entry_point == int(np.argmax(lv == lv.max()))
    layers = snap.export_graph()
    assert len(layers) == snap.max_layer + 1
    for l, (rp, cols) in enumerate(layers):
        deg = np.diff(rp.astype(np.int64))
        maxc = 2 * M if l == 0 else M
        assert deg.max() <= maxc
        members = np.nonzero(lv >= l)[0]
        assert (deg[lv < l] == 0).all()                     # only nodes of level >= l have links on layer l
        if len(members) > 1:
            assert (deg[members] >= 1).all()
        assert np.isin(cols, members).all()                  # links stay inside the layer
        owner = np.repeat(np.arange(n), deg)
        assert (cols != owner).all()                         # no self loops
        pairs = owner.astype(np.int64) * n + cols
        assert len(np.unique(pairs)) == len(pairs)           # no duplicate links
    # search parity given this graph: the oracle on the exported graph must agree bit for bit
    g = vo.Hnsw.from_arrays(metric, x, layers, M, 2 * M, snap.entry_point, snap.max_layer)
    q = queries_near(x, 200, jitter=0.05, seed=4)
    ids, dist, cnt, st = snap.search_batch(q, 10, 64, with_stats=True)
    oi, od, oc, ost = g.search_batch(q, 10, 64, order="canonical", threads=8)
    keep = ost[:, 4] == 0
    assert np.array_equal(ids[keep], oi[keep].astype(np.uint32)) and bits_equal(dist, od)
    assert np.array_equal(st[:, 0], ost[:, 0].astype(np.uint32))
    # quality: recall@10 against exact brute force (cosine/L2; MIPS graphs are not metric)
    if metric != vo.DOT:
        bi, _ = snap.bruteforce_batch(q, 10)
        rec = np.mean([len(set(ids[i].tolist()) & set(bi[i].tolist())) / 10 for i in range(len(q))])
        assert rec >= 0.95, rec


def test_build_edge_cases():
    e = DeviceSnapshot.from_vectors(np.zeros((0, 16), np.float32), DistanceMetric.Cosine)
    e.build_graph(16)
    ids, dist, cnt = e.search_batch(np.ones((2, 16), np.float32), 5, 32)
    assert (cnt == 0).all()
    one = DeviceSnapshot.from_vectors(np.ones((1, 16), np.float32), DistanceMetric.Cosine)
    one.build_graph(16)
    ids, dist, cnt = one.search_batch(np.ones((2, 16), np.float32), 5, 32)
    assert (cnt == 1).all() and (ids[:, 0] == 0).all()
    few = DeviceSnapshot.from_vectors(latent_data(5, 16, seed=1), DistanceMetric.Euclidean)
    few.build_graph(16)
    ids, dist, cnt = few.search_batch(latent_data(5, 16, seed=1), 5, 32)
    assert (cnt == 5).all() and (ids[:, 0] == np.arange(5)).all()


@pytest.mark.parametrize("metric,dim,n,M,ef_c", [(vo.COSINE, 48, 1500, 8, 60), (vo.EUCLIDEAN, 32, 1200, 16, 100),
                                                 (vo.DOT, 64, 800, 8, 40), (vo.COSINE, 768, 400, 16, 64)])
def test_exact_sequential_build_equals_reference_graph(metric, dim, n, M, ef_c):
    """Build parity: veles_index_build_graph_exact restates NativeHnsw::insert (graph.rs:158-237) for nodes in
    id order; every adjacency list, on every layer, must equal the oracle's (= the reference's deterministic
    sequential build), id for id and in the same order."""
    x = latent_data(n, dim, latent=8, noise=0.4, seed=11 + dim)
    g = vo.Hnsw(metric, dim, M=M, ef_construction=ef_c)
    g.insert_many(x)
    snap = DeviceSnapshot.from_vectors(x, metric)
    snap.build_graph_exact(M, ef_c)
    assert snap.max_layer == g.max_layer and snap.entry_point == g.entry_point
    ref_layers = g.export_graph()
    got_layers = snap.export_graph()
    assert len(got_layers) == len(ref_layers)
    for l, ((rp, cols), (orp, ocols)) in enumerate(zip(got_layers, ref_layers)):
        assert np.array_equal(rp, orp), f"layer {l}: degrees differ"
        assert np.array_equal(cols, ocols), f"layer {l}: neighbour lists differ"
    # and the search on it agrees as well
    q = queries_near(x, 64, seed=2)
    ids, dist, cnt = snap.search_batch(q, 10, 64)
    oi, od, oc, ost = g.search_batch(q, 10, 64, order="canonical", threads=8)
    keep = ost[:, 4] == 0
    assert np.array_equal(ids[keep], oi[keep].astype(np.uint32)) and bits_equal(dist, od)


def test_exact_build_edge_cases():
    for n in (0, 1, 2, 5):
        x = latent_data(max(n, 1), 16, seed=n)[:n]
        snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Euclidean)
        snap.build_graph_exact(8, 32)
        g = vo.Hnsw(vo.EUCLIDEAN, 16, M=8, ef_construction=32)
        if n:
            g.insert_many(x)
            for (rp, cols), (orp, ocols) in zip(snap.export_graph(), g.export_graph()):
                assert np.array_equal(rp, orp) and np.array_equal(cols, ocols)
        ids, dist, cnt = snap.search_batch(np.zeros((1, 16), np.float32), 3, 16)
        assert cnt[0] == min(n, 3)


def test_host_mirror_sequential_inserts_reproduce_the_reference_index():
    """index_tests.rs:1108-1159 through the host mirror: HnswIndex::insert one by one (HnswParams::auto(64) =
    M 24 / ef_construction 300), then Accurate search.  The mirror builds with the exact sequential builder,
    so ids and scores must equal the oracle's end to end (insert -> search -> id map -> transform_score)."""
    import math

    from velesdb_b200 import HnswIndex, SearchQuality

    dim, n, k = 64, 500, 10
    data = np.array([[math.sin(np.float32(i * dim + j) * np.float32(0.001)) for j in range(dim)] for i in range(n)], np.float32)
    ix = HnswIndex(dim, DistanceMetric.Cosine)
    g = vo.Hnsw(vo.COSINE, dim, M=24, ef_construction=300)
    for i in range(n):
        ix.insert(i, data[i])
        g.insert(data[i])
    q = np.array([math.sin(np.float32(j) * np.float32(0.001)) for j in range(dim)], np.float32)
    res = ix.search_with_quality(q, k, SearchQuality.Accurate)
    oi, od, st = g.search(q, k, vo.ef_search(vo.ACCURATE, k), order="canonical", with_stats=True)
    if not st["tie_at_k"]:
        assert [r[0] for r in res] == oi.tolist()
    assert bits_equal([r[1] for r in res], [vo.transform_score(vo.COSINE, d) for d in od])
    gt, _ = vo.bruteforce(vo.COSINE, data, q, k)
    assert len(set(r[0] for r in res) & set(gt.tolist())) / k >= 0.8   # the reference's own assertion
