"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bars: node ids and integer distances bit-exact; f32 distances bit-exact (the tests
assert 0 ulp, tighter than the 1e-4 relative tolerance BASELINE.json allows)."""
import os
import tempfile

import numpy as np
import pytest

from oracle import oracle as vo
from tests.gpu_util import bits_equal, build_oracle, latent_data, queries_near
from velesdb_b200 import DeviceSnapshot, DistanceMetric, distance_pairs

pytestmark = pytest.mark.gpu

METRICS = [vo.COSINE, vo.EUCLIDEAN, vo.DOT, vo.HAMMING, vo.JACCARD]


@pytest.mark.parametrize("dim", [1, 3, 7, 8, 15, 16, 17, 24, 31, 32, 33, 40, 47, 63, 64, 100, 128, 384, 768, 771, 1536])
def test_distance_kernels_bit_exact(dim):
    # simd_avx512.rs:87-352 / simd_explicit.rs:50-443 / native/distance.rs:75-85
    rng = np.random.default_rng(dim)
    a = rng.normal(size=(64, dim)).astype(np.float32)
    b = rng.normal(size=(64, dim)).astype(np.float32)
    b[0] = a[0]
    a[1] = 0
    b[2] = -a[2]
    for m in METRICS:
        if m in (vo.HAMMING, vo.JACCARD):
            aa, bb = (a > 0.3).astype(np.float32), (b > 0.1).astype(np.float32)
        else:
            aa, bb = a, b
        for as_value in (False, True):
            got = distance_pairs(m, aa, bb, as_value)
            f = vo.metric_value if as_value else vo.graph_distance
            ref = np.array([f(m, aa[i], bb[i]) for i in range(64)], np.float32)
            assert bits_equal(got, ref), (m, as_value, dim)


def test_known_answer_distances_on_gpu():
    # simd_avx512_tests.rs:45-135, native/distance.rs:224-259,453-483, simd_dispatch.rs:484-508
    one, two = np.ones((1, 16), np.float32), np.full((1, 16), 2.0, np.float32)
    assert distance_pairs(vo.DOT, one, two, True)[0] == 32.0
    z, b = np.zeros((1, 16), np.float32), np.zeros((1, 16), np.float32)
    b[0, 0], b[0, 1] = 3, 4
    assert distance_pairs(vo.EUCLIDEAN, z, b)[0] == 5.0
    assert distance_pairs(vo.COSINE, z, b)[0] == 1.0  # zero norm
    assert distance_pairs(vo.COSINE, b, b)[0] == 0.0
    assert distance_pairs(vo.COSINE, b, -b, True)[0] == -1.0
    o32, z32 = np.ones((1, 32), np.float32), np.zeros((1, 32), np.float32)
    assert distance_pairs(vo.HAMMING, o32, z32)[0] == 32.0 and distance_pairs(vo.HAMMING, o32, o32)[0] == 0.0
    assert distance_pairs(vo.JACCARD, z32, z32, True)[0] == 1.0
    assert distance_pairs(vo.DOT, np.array([[1, 2, 3]], np.float32), np.array([[4, 5, 6]], np.float32))[0] == -32.0


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dim", [12, 64, 100, 768])
def test_bruteforce_bit_exact(metric, dim):
    # index/hnsw/index/search.rs:176-219 with sort_results (core/distance.rs:95-103)
    n, nq, k = 1500, 37, 10
    x = latent_data(n, dim, seed=dim)
    if metric in (vo.HAMMING, vo.JACCARD):
        x = (x > 0.2).astype(np.float32)
    q = queries_near(x, nq, seed=dim + 1)
    if metric in (vo.HAMMING, vo.JACCARD):
        q = (q > 0.5).astype(np.float32)
    snap = DeviceSnapshot.from_vectors(x, metric)
    ids, sc = snap.bruteforce_batch(q, k)
    ri, rs = vo.bruteforce_batch(metric, x, q, k, threads=8)
    assert np.array_equal(ids, ri.astype(np.uint32))
    assert bits_equal(sc, rs)


@pytest.mark.parametrize("metric", [vo.COSINE, vo.EUCLIDEAN, vo.DOT])
@pytest.mark.parametrize("store", ["f32", "f16"])
def test_bruteforce_few_queries_scan_path(metric, store):
    # nq <= 8 takes the HBM-bound scan kernel (bf_scan_kernel); n >= 65536 the two-level top-k
    for n, dim in ((3001, 768), (70_001, 64)):
        x = latent_data(n, dim, seed=n % 97 + metric)
        x[n // 2] = x[7]  # an exact duplicate: equal scores, ordered by id
        snap = DeviceSnapshot.from_vectors(x, metric, store_dtype=store)
        xo = x.astype(np.float16).astype(np.float32) if store == "f16" else x
        for nq in (1, 2, 3, 5, 8, 20, 45):  # > 8: the tile kernels (shared-memory staged when dim % 128 == 0)
            q = queries_near(x, nq, seed=nq)
            q[0] = x[7]
            for k in (1, 10, 100):
                ids, sc = snap.bruteforce_batch(q, k)
                ri, rs = vo.bruteforce_batch(metric, xo, q, k, threads=8)
                assert np.array_equal(ids, ri.astype(np.uint32)), (n, nq, k)
                assert bits_equal(sc, rs), (n, nq, k)


def test_bruteforce_edges():
    x = latent_data(5, 32)
    snap = DeviceSnapshot.from_vectors(x, vo.EUCLIDEAN)
    ids, sc = snap.bruteforce_batch(x[:2], 10)  # k > n: padded
    assert (ids[:, 5:] == 0xFFFFFFFF).all() and np.isnan(sc[:, 5:]).all()
    assert ids[0, 0] == 0 and sc[0, 0] == 0.0 and ids[1, 0] == 1
    empty = DeviceSnapshot.from_vectors(np.zeros((0, 32), np.float32), vo.COSINE)
    ids, sc = empty.bruteforce_batch(x[:1], 3)
    assert (ids == 0xFFFFFFFF).all()
    ids, sc = snap.bruteforce_batch(np.zeros((0, 32), np.float32), 3)
    assert ids.shape == (0, 3)


GRAPHS = {}


def graph_case(metric, dim, n=2000, M=16, ef_c=100, binary=False):
    key = (metric, dim, n, M, ef_c, binary)
    if key not in GRAPHS:
        x = latent_data(n, dim, seed=metric * 100 + dim)
        if binary:
            x = (x > 0.0).astype(np.float32)
        g = build_oracle(metric, x, M, ef_c)
        snap = DeviceSnapshot.from_arrays(x, metric, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
        GRAPHS[key] = (x, g, snap)
    return GRAPHS[key]


def check_search(g, snap, q, k, ef):
    ids, dist, cnt, st = snap.search_batch(q, k, ef, with_stats=True)
    oi, od, oc, ost = g.search_batch(q, k, ef, order="canonical", threads=8)
    assert np.array_equal(cnt, oc)
    for r in range(q.shape[0]):
        c = int(cnt[r])
        if ost[r, 4]:  # a distance tie crosses the k boundary: the reference's own order is heap-internal
            assert bits_equal(dist[r, :c], od[r, :c])
            continue
        assert np.array_equal(ids[r, :c], oi[r, :c].astype(np.uint32)), (r, ids[r], oi[r])
        assert bits_equal(dist[r, :c], od[r, :c])
        assert (ids[r, c:] == 0xFFFFFFFF).all()
    # identical traversal: same number of distance evaluations and expansions
    assert np.array_equal(st[:, 0], ost[:, 0].astype(np.uint32))
    assert np.array_equal(st[:, 1], ost[:, 1].astype(np.uint32))
    assert np.array_equal(st[:, 2], ost[:, 2].astype(np.uint32))
    assert np.array_equal(st[:, 3], ost[:, 3].astype(np.uint32))


@pytest.mark.parametrize("metric", [vo.COSINE, vo.EUCLIDEAN, vo.DOT])
@pytest.mark.parametrize("dim", [20, 96, 768])
def test_hnsw_search_bit_exact(metric, dim):
    # native/graph.rs:251-270, 405-428, 438-520
    n = 2000 if dim < 768 else 1200
    x, g, snap = graph_case(metric, dim, n=n)
    q = queries_near(x, 64, seed=5)
    for k, ef in ((10, 64), (1, 1), (10, 16), (100, 256), (10, 5)):
        check_search(g, snap, q, k, max(ef, 1))


@pytest.mark.parametrize("metric", [vo.HAMMING, vo.JACCARD])
def test_hnsw_search_integer_metrics_with_ties(metric):
    # Hamming / Jaccard on {0,1} lanes: equal distances everywhere -> exercises the tie list
    x, g, snap = graph_case(metric, 64, n=1500, binary=True)
    q = (queries_near(x, 64, jitter=0.4, seed=9) > 0.5).astype(np.float32)
    for k, ef in ((10, 64), (10, 16), (50, 128)):
        check_search(g, snap, q, k, ef)


def test_hnsw_reference_known_answers_on_gpu():
    # native/graph_tests.rs:10-30 (ramp, first hit is node 0) and :33-41 (empty index)
    g = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100)
    x = np.array([[i * 32 + j for j in range(32)] for i in range(100)], np.float32)
    g.insert_many(x)
    snap = DeviceSnapshot.from_arrays(x, vo.EUCLIDEAN, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    ids, dist, cnt = snap.search_batch(np.arange(32, dtype=np.float32), 10, 50)
    assert cnt[0] == 10 and ids[0, 0] == 0
    e = vo.Hnsw(vo.COSINE, 3, M=16, ef_construction=100)
    esnap = DeviceSnapshot.from_arrays(np.zeros((0, 3), np.float32), vo.COSINE, e.export_graph(), 16, 32, 0, 0)
    ids, dist, cnt = esnap.search_batch(np.array([1, 2, 3], np.float32), 10, 50)
    assert cnt[0] == 0 and (ids == 0xFFFFFFFF).all()
    # single node
    s = vo.Hnsw(vo.COSINE, 3, M=16, ef_construction=100)
    s.insert(np.array([1, 0, 0], np.float32))
    ssnap = DeviceSnapshot.from_arrays(np.array([[1, 0, 0]], np.float32), vo.COSINE, s.export_graph(), 16, 32, 0, 0)
    ids, dist, cnt = ssnap.search_batch(np.array([1, 2, 3], np.float32), 10, 50)
    assert cnt[0] == 1 and ids[0, 0] == 0


def test_file_format_v1_interchange():
    # native/backend_adapter.rs:184-380: oracle dump -> GPU load; GPU dump -> oracle load
    x, g, snap = graph_case(vo.EUCLIDEAN, 96)
    q = queries_near(x, 32, seed=11)
    with tempfile.TemporaryDirectory() as d:
        g.dump(d, "native_hnsw")
        loaded = DeviceSnapshot.from_reference_files(d, DistanceMetric.Euclidean)
        assert len(loaded) == len(g) and loaded.max_layer == g.max_layer and loaded.entry_point == g.entry_point
        check_search(g, loaded, q, 10, 64)
        os.makedirs(os.path.join(d, "back"))
        loaded.dump(os.path.join(d, "back"), "native_hnsw")
        for name in ("native_hnsw.vectors", "native_hnsw.graph"):
            assert open(os.path.join(d, name), "rb").read() == open(os.path.join(d, "back", name), "rb").read()
        g2 = vo.Hnsw.load(os.path.join(d, "back"), vo.EUCLIDEAN)
        check_search(g2, snap, q, 10, 64)


def test_f16_storage_matches_oracle_on_rounded_vectors():
    # config C3 semantics: f16 store, candidates up-converted to f32 (core/half_precision.rs:97 rounding)
    x, g, _ = graph_case(vo.COSINE, 96)
    xh = x.astype(np.float16).astype(np.float32)
    gh = vo.Hnsw.from_arrays(vo.COSINE, xh, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    snap = DeviceSnapshot.from_arrays(x, vo.COSINE, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer,
                                      store_dtype="f16")
    q = queries_near(x, 48, seed=13)
    check_search(gh, snap, q, 10, 64)
    ids, sc = snap.bruteforce_batch(q, 10)
    ri, rs = vo.bruteforce_batch(vo.COSINE, xh, q, 10, threads=8)
    assert np.array_equal(ids, ri.astype(np.uint32)) and bits_equal(sc, rs)


def test_packed_binary_storage_matches_f32_threshold_form():
    # config C4 semantics: packed bits + popcount == f32 lanes thresholded at > 0.5 (simd_explicit.rs:256-360)
    x, g, _ = graph_case(vo.HAMMING, 64, n=1500, binary=True)
    snap = DeviceSnapshot.from_arrays(x, vo.HAMMING, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer,
                                      store_dtype="bin1")
    q = (queries_near(x, 64, jitter=0.4, seed=9) > 0.5).astype(np.float32)
    check_search(g, snap, q, 10, 64)
    ids, sc = snap.bruteforce_batch(q, 10)
    ri, rs = vo.bruteforce_batch(vo.HAMMING, x, q, 10, threads=8)
    assert np.array_equal(ids, ri.astype(np.uint32)) and bits_equal(sc, rs)
    # already-packed input (u64 words, LSB first)
    packed = np.packbits(x.astype(np.uint8), axis=1, bitorder="little").view(np.uint64)
    snap2 = DeviceSnapshot.from_arrays(packed, vo.HAMMING, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer,
                                       store_dtype="bin1", src_dtype="bin1", dim=64)
    check_search(g, snap2, q, 10, 64)


def test_persistent_grid_more_queries_than_slots():
    # nq far above the resident-warp count: the work counter must hand out every query exactly once,
    # and the visited bitmap must be clean for each new query of a slot
    x, g, snap = graph_case(vo.COSINE, 20)
    q = queries_near(x, 5000, seed=17)
    check_search(g, snap, q, 10, 32)


@pytest.mark.parametrize("dim", [128, 256])
def test_packed_binary_wide_rows(dim):
    # dim % 128 == 0 takes the four-rows-per-step path of the search kernel
    x, g, _ = graph_case(vo.HAMMING, dim, n=1200, binary=True)
    snap = DeviceSnapshot.from_arrays(x, vo.HAMMING, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer,
                                      store_dtype="bin1")
    q = (queries_near(x, 64, jitter=0.4, seed=9) > 0.5).astype(np.float32)
    check_search(g, snap, q, 10, 64)


@pytest.mark.parametrize("metric", [vo.COSINE, vo.EUCLIDEAN, vo.DOT])
def test_hnsw_search_f16_quad_path(metric):
    x, g, _ = graph_case(metric, 96)
    xh = x.astype(np.float16).astype(np.float32)
    gh = vo.Hnsw.from_arrays(metric, xh, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    snap = DeviceSnapshot.from_arrays(x, metric, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer,
                                      store_dtype="f16")
    check_search(gh, snap, queries_near(x, 48, seed=13), 10, 64)


@pytest.mark.parametrize("warps", ["1", "2", "4", "8"])
def test_hnsw_search_warps_per_query(warps, monkeypatch):
    # one warp per query vs the cooperative multi-warp kernel (hnsw_search_kernel<.., COOP>): identical results,
    # identical per-layer evaluation / expansion counts
    monkeypatch.setenv("VELES_SEARCH_WARPS", warps)
    for metric, dim in ((vo.COSINE, 768), (vo.EUCLIDEAN, 96), (vo.DOT, 96)):
        x, g, snap = graph_case(metric, dim, n=2000 if dim < 768 else 1200)
        q = queries_near(x, 64, seed=5)
        for k, ef in ((10, 64), (1, 1), (100, 256), (10, 300)):
            check_search(g, snap, q, k, ef)
    x, g, snap = graph_case(vo.HAMMING, 64, n=1500, binary=True)  # ties everywhere (generic path: always one warp)
    check_search(g, snap, (queries_near(x, 64, jitter=0.4, seed=9) > 0.5).astype(np.float32), 10, 64)
    x, g, _ = graph_case(vo.COSINE, 96)
    xh = x.astype(np.float16).astype(np.float32)
    gh = vo.Hnsw.from_arrays(vo.COSINE, xh, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    snap = DeviceSnapshot.from_arrays(x, vo.COSINE, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer, store_dtype="f16")
    check_search(gh, snap, queries_near(x, 48, seed=13), 10, 64)


def test_mapped_search_id_map_tombstones_and_filter():
    # veles_search_batch_mapped vs the host-side restatement of batch.rs:178-196 / search.rs:86-91 /
    # collection/search/vector.rs:182-211 on top of the raw search
    from velesdb_b200 import _native as nv
    for metric in (vo.COSINE, vo.EUCLIDEAN, vo.DOT):
        x, g, snap = graph_case(metric, 96)
        n = len(x)
        q = queries_near(x, 40, seed=21)
        rng = np.random.default_rng(5)
        ext = (rng.permutation(n).astype(np.uint64) * np.uint64(1_000_003) + np.uint64(1 << 40))  # ids beyond 32 bits
        live = rng.random(n) > 0.3
        allow = rng.random(n) > 0.5
        pack = lambda b: np.packbits(b, bitorder="little").view(np.uint8).tobytes()
        bits = lambda b: np.frombuffer(pack(np.concatenate([b, np.zeros((-len(b)) % 32, bool)])), np.uint32).copy()
        snap.set_id_map(ext, bits(live))
        k, kf, ef = 10, 50, 128
        raw_i, raw_d, raw_c = snap.search_batch(q, kf, ef)
        for use_allow in (False, True):
            ids, sc, cnt = snap.search_batch_mapped(q, k, ef, k_fetch=kf, allow_bits=bits(allow) if use_allow else None)
            for r in range(len(q)):
                keep = [j for j in range(int(raw_c[r])) if live[raw_i[r, j]] and (not use_allow or allow[raw_i[r, j]])][:k]
                assert cnt[r] == len(keep)
                assert ids[r, :len(keep)].tolist() == [int(ext[raw_i[r, j]]) for j in keep]
                want = np.array([nv.lib().veles_transform_score(int(metric), float(raw_d[r, j])) for j in keep], np.float32)
                assert bits_equal(sc[r, :len(keep)], want)
                assert (ids[r, len(keep):] == np.uint64(0xFFFFFFFFFFFFFFFF)).all() and np.isnan(sc[r, len(keep):]).all()
        snap.set_id_map(None, None)  # back to identity / all live for the other tests sharing this snapshot
        ids, sc, cnt = snap.search_batch_mapped(q, k, ef)
        ri, rd, rc = snap.search_batch(q, k, ef)
        assert np.array_equal(ids, ri.astype(np.uint64)) and np.array_equal(cnt, rc)


@pytest.mark.parametrize("warps", ["1", "4"])
def test_search_multi_entry_bit_exact(warps, monkeypatch):
    # NativeHnsw::search_multi_entry (graph.rs:288-348): the oracle draws the probes from its xorshift state, the
    # host helper reproduces the draws, the device starts layer 0 from the same entry points
    from velesdb_b200 import multi_entry_probes
    monkeypatch.setenv("VELES_SEARCH_WARPS", warps)
    for metric, dim in ((vo.COSINE, 96), (vo.EUCLIDEAN, 768)):
        x, g, snap = graph_case(metric, dim, n=2000 if dim < 768 else 1200)
        q = queries_near(x, 24, seed=31)
        n = len(x)
        for probes, k, ef in ((3, 10, 64), (4, 5, 16), (2, 10, 100), (1, 10, 64)):
            state = g.rng_state
            rows, want = [], []
            for r in range(len(q)):
                row, state = multi_entry_probes(state, n, probes)
                rows.append(row)
                ids, d, st, ent = g.search_multi_entry(q[r], k, ef, probes, order="canonical")
                assert set(ent[1:]) <= set(row)
                want.append((ids, d, st))
            assert state == g.rng_state
            ids, dist, cnt, st = snap.search_batch_multi_entry(q, k, ef, np.array(rows, np.uint32), with_stats=True)
            for r, (oi, od, ost) in enumerate(want):
                assert cnt[r] == len(oi)
                if not ost["tie_at_k"]:
                    assert ids[r, :len(oi)].tolist() == oi.tolist(), (metric, probes, r)
                assert bits_equal(dist[r, :len(oi)], od)
                assert (st[r, 0], st[r, 1]) == (ost["ndc0"], ost["hops0"])
    with pytest.raises(Exception, match="ef >= 4"):
        snap.search_batch_multi_entry(q, 2, 2, np.full((len(q), 3), 0xFFFFFFFF, np.uint32))


@pytest.mark.parametrize("dim", [16, 17, 40, 100, 771])
def test_hnsw_search_dims_with_tails_on_the_quad_path(dim, monkeypatch):
    # dims that are not multiples of 32: 32-element blocks, then the reference's 8-wide and scalar tails
    # (simd_avx512.rs:186-201) inside the four-rows-per-step evaluator, f32 and f16 rows, 1 and 4 warps per query
    for metric in (vo.COSINE, vo.EUCLIDEAN, vo.DOT):
        x, g, snap = graph_case(metric, dim, n=1200)
        q = queries_near(x, 32, seed=dim)
        for warps in ("1", "4"):
            monkeypatch.setenv("VELES_SEARCH_WARPS", warps)
            check_search(g, snap, q, 10, 64)
        monkeypatch.delenv("VELES_SEARCH_WARPS")
    xh = x.astype(np.float16).astype(np.float32)
    gh = vo.Hnsw.from_arrays(vo.DOT, xh, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    sh = DeviceSnapshot.from_arrays(x, vo.DOT, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer, store_dtype="f16")
    check_search(gh, sh, q, 10, 64)
