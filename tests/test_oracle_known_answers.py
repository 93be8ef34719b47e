"""Pins the CPU oracle against every known-answer test the reference holds for the
hot path.  Each test names the reference test it restates.
Paths are relative to /root/reference/crates/velesdb-core/src.
"""
import math
import os
import tempfile

import numpy as np
import pytest

from oracle import oracle as vo

F = np.float32


def sinvec(n, seed):
    # simd_avx512_tests.rs generate_test_vector / benches/simd_benchmark.rs:18-21
    return np.array([math.sin(seed + 0.1 * i) for i in range(n)], dtype=F)


@pytest.mark.parametrize("fma", [True, False])
def test_dot_product_auto_basic(fma):
    # simd_avx512_tests.rs:45-53
    assert vo.dot(np.ones(16, F), np.full(16, 2.0, F), fma) == 32.0


@pytest.mark.parametrize("fma", [True, False])
def test_dot_product_auto_768d_vs_scalar(fma):
    # simd_avx512_tests.rs:56-66 (1e-4 rel vs sequential scalar sum)
    a, b = sinvec(768, 0.0), sinvec(768, 1.0)
    s = F(0)
    for x, y in zip(a, b):
        s = F(s + F(x * y))
    r = vo.dot(a, b, fma)
    assert abs(r - s) / max(abs(s), 1.0) < 1e-4


def test_squared_l2_and_euclid_known():
    # simd_avx512_tests.rs:88-135
    a = np.zeros(16, F)
    b = np.zeros(16, F)
    b[0], b[1] = 3.0, 4.0
    assert vo.l2sq(a, b) == 25.0
    assert vo.metric_value(vo.EUCLIDEAN, a, b) == 5.0
    v = sinvec(768, 0.0)
    assert vo.l2sq(v, v) == 0.0


def test_cosine_known_answers():
    # simd_avx512_tests.rs cosine block; native/distance.rs:224-233
    v = sinvec(768, 0.0)
    assert abs(vo.metric_value(vo.COSINE, v, v) - 1.0) < 1e-6
    e0 = np.zeros(16, F)
    e1 = np.zeros(16, F)
    e0[0] = 1
    e1[1] = 1
    assert vo.metric_value(vo.COSINE, e0, e1) == 0.0
    assert vo.metric_value(vo.COSINE, e0, -e0) == -1.0
    # zero norm -> similarity 0 -> graph distance 1 (simd_avx512.rs:347-349, distance.rs:575-585)
    z = np.zeros(16, F)
    assert vo.metric_value(vo.COSINE, z, e0) == 0.0
    assert vo.graph_distance(vo.COSINE, z, e0) == 1.0
    small = np.array([1, 2, 3], F)
    assert abs(vo.graph_distance(vo.COSINE, small, small)) < 1e-5


def test_small_dim_paths():
    # native/distance.rs:236-243 (3-4-5), :453-461 (dot -32)
    assert abs(vo.graph_distance(vo.EUCLIDEAN, np.zeros(3, F), np.array([3, 4, 0], F)) - 5.0) < 1e-6
    assert vo.graph_distance(vo.DOT, np.array([1, 2, 3], F), np.array([4, 5, 6], F)) == -32.0


def test_simd_matches_scalar_768():
    # native/distance.rs:246-259
    a = np.array([math.sin(i * 0.01) for i in range(768)], F)
    b = np.array([math.cos(i * 0.02) for i in range(768)], F)
    dot = float(np.dot(a.astype(np.float64), b.astype(np.float64)))
    ref = 1.0 - dot / (np.linalg.norm(a.astype(np.float64)) * np.linalg.norm(b.astype(np.float64)))
    assert abs(vo.graph_distance(vo.COSINE, a, b) - ref) < 1e-4


def test_hamming_and_jaccard_known():
    # simd_dispatch.rs:484-508 (0 / 32 / 16); native/distance.rs:464-471 (2); :284-306 (1-32/48); :474-483 (2/3)
    one, zero = np.ones(32, F), np.zeros(32, F)
    assert vo.graph_distance(vo.HAMMING, one, one) == 0.0
    assert vo.graph_distance(vo.HAMMING, one, zero) == 32.0
    half = one.copy()
    half[:16] = 0
    assert vo.graph_distance(vo.HAMMING, one, half) == 16.0
    assert vo.graph_distance(vo.HAMMING, np.array([1, 0, 1, 0], F), np.array([1, 1, 0, 0], F)) == 2.0
    a = np.array([1.0 if i < 32 else 0.0 for i in range(64)], F)
    b = np.array([1.0 if i < 48 else 0.0 for i in range(64)], F)
    assert abs(vo.graph_distance(vo.JACCARD, a, b) - (1.0 - 32.0 / 48.0)) < 1e-6
    assert abs(vo.graph_distance(vo.JACCARD, np.array([1, 1, 0, 0], F), np.array([1, 0, 1, 0], F)) - 2.0 / 3.0) < 1e-6
    # empty sets: similarity 1 (simd_explicit.rs:431-436)
    assert vo.metric_value(vo.JACCARD, zero, zero) == 1.0


def test_more_distance_known_answers_from_the_simd_test_files():
    # simd_tests.rs:24-127, 186-322; simd_explicit_tests.rs:19-57, 178-249; simd_native_tests.rs:28-66, 134-262 --
    # exact-valued inputs, so the oracle must hit the expected numbers exactly (the reference tolerates 1e-5)
    A = lambda *v: np.array(v, F)
    v4 = A(1, 2, 3, 4)
    assert abs(vo.metric_value(vo.COSINE, v4, v4) - 1.0) < 1e-6 and abs(vo.metric_value(vo.COSINE, v4, -v4) + 1.0) < 1e-6
    assert vo.metric_value(vo.COSINE, A(1, 0, 0, 0), A(0, 1, 0, 0)) == 0.0
    assert vo.metric_value(vo.COSINE, A(1, 2, 3), A(0, 0, 0)) == 0.0          # zero vector -> 0
    assert vo.metric_value(vo.COSINE, A(1, 0), A(0, 1)) == 0.0
    assert vo.metric_value(vo.EUCLIDEAN, v4, v4) == 0.0
    assert vo.metric_value(vo.EUCLIDEAN, A(0, 0, 0), A(3, 4, 0)) == 5.0      # 3-4-5
    assert vo.metric_value(vo.EUCLIDEAN, np.zeros(8, F), A(3, 4, 0, 0, 0, 0, 0, 0)) == 5.0
    assert vo.metric_value(vo.EUCLIDEAN, A(3), A(4)) == 1.0
    assert vo.l2sq(A(0, 0), A(3, 4)) == 25.0 and vo.l2sq(A(1, 2, 3), A(1, 2, 3)) == 0.0
    assert vo.l2sq(v4, A(5, 6, 7, 8)) == 64.0
    assert vo.dot(v4, A(5, 6, 7, 8)) == 70.0
    assert vo.dot(A(1, 2, 3, 4, 5, 6, 7, 8), np.ones(8, F)) == 36.0
    assert vo.dot(A(1, 2, 3, 4, 5), A(5, 4, 3, 2, 1)) == 35.0                # odd dimension
    assert vo.dot(A(3), A(4)) == 12.0
    assert vo.dot(np.zeros(16, F), np.ones(16, F)) == 0.0 and vo.dot(np.ones(16, F), np.ones(16, F)) == 16.0
    assert vo.dot(np.ones(32, F), np.ones(32, F)) == 32.0
    assert vo.dot(np.arange(19, dtype=F), np.ones(19, F)) == 171.0           # 16 + remainder 3
    a = np.array([i * 0.001 for i in range(768)], F)
    b = np.array([(768 - i) * 0.001 for i in range(768)], F)
    assert abs(vo.dot(a, b) - float(np.dot(a.astype(np.float64), b.astype(np.float64)))) < 0.01
    assert abs(vo.l2sq(a, b) - float(((a.astype(np.float64) - b) ** 2).sum())) < 0.01
    x, y = sinvec(768, 0.0), sinvec(768, 1.0)
    naive = float(np.sqrt(((x.astype(np.float64) - y) ** 2).sum()))
    assert abs(vo.metric_value(vo.EUCLIDEAN, x, y) - naive) < 1e-5 * max(naive, 1.0)
    # Hamming on f32 lanes (threshold 0.5), Jaccard
    assert vo.metric_value(vo.HAMMING, A(1, 0, 1, 0), A(1, 0, 1, 0)) == 0.0
    assert vo.metric_value(vo.HAMMING, A(1, 0, 1, 0), A(0, 1, 0, 1)) == 4.0
    assert vo.metric_value(vo.HAMMING, A(1, 1, 0, 0), A(1, 0, 0, 1)) == 2.0
    assert vo.metric_value(vo.HAMMING, A(1, 0, 1, 0, 1), A(0, 0, 1, 1, 1)) == 2.0
    assert vo.metric_value(vo.HAMMING, A(1, 0, 1, 0, 1, 0, 1, 0), A(0, 1, 0, 1, 0, 1, 0, 1)) == 8.0
    assert vo.metric_value(vo.HAMMING, A(1, 1, 0, 0, 1, 1, 0, 0), A(1, 0, 0, 1, 1, 0, 0, 1)) == 4.0
    h3 = np.array([1.0 if i % 3 == 0 else 0.0 for i in range(768)], F)
    h2 = np.array([1.0 if i % 2 == 0 else 0.0 for i in range(768)], F)
    assert vo.metric_value(vo.HAMMING, h3, h2) == float(((h3 > 0.5) != (h2 > 0.5)).sum())
    assert vo.metric_value(vo.JACCARD, A(1, 0, 1, 0), A(1, 0, 1, 0)) == 1.0
    assert vo.metric_value(vo.JACCARD, A(1, 0, 0, 0), A(0, 1, 0, 0)) == 0.0
    assert abs(vo.metric_value(vo.JACCARD, A(1, 1, 0, 0), A(1, 0, 1, 0)) - 1.0 / 3.0) < 1e-6
    assert vo.metric_value(vo.JACCARD, np.zeros(4, F), np.zeros(4, F)) == 1.0
    # packed u64 Hamming (simd_explicit_tests.rs:229-249)
    ff = np.full(16, 0xFFFFFFFFFFFFFFFF, np.uint64)
    assert vo.hamming_binary(ff, ff) == 0
    assert vo.hamming_binary(np.zeros(1, np.uint64), ff[:1]) == 64
    assert vo.hamming_binary(np.array([0b10101010], np.uint64), np.array([0b01010101], np.uint64)) == 8
    assert vo.hamming_binary(np.zeros(16, np.uint64), ff) == 1024


def test_packed_hamming_equals_f32_threshold_form():
    # simd_explicit.rs:308-360 vs :256-287 on {0,1} inputs
    rng = np.random.default_rng(3)
    bits_a = rng.integers(0, 2, 1024).astype(np.uint8)
    bits_b = rng.integers(0, 2, 1024).astype(np.uint8)
    pa = np.packbits(bits_a, bitorder="little").view(np.uint64)
    pb = np.packbits(bits_b, bitorder="little").view(np.uint64)
    assert vo.hamming_binary(pa, pb) == int(vo.graph_distance(vo.HAMMING, bits_a.astype(F), bits_b.astype(F)))


@pytest.mark.parametrize("n", [16, 17, 24, 31, 32, 33, 40, 47, 100, 128, 768, 771, 1536])
@pytest.mark.parametrize("fma", [True, False])
def test_avx_path_equals_scalar_emulation(n, fma):
    # the AVX2 fast path of the oracle and its lane-by-lane emulation must agree bit for bit
    a, b = sinvec(n, 0.3), sinvec(n, 1.7)
    vo.force_scalar(False)
    fast = (vo.dot(a, b, fma), vo.l2sq(a, b, fma), vo.metric_value(vo.COSINE, a, b, fma))
    vo.force_scalar(True)
    slow = (vo.dot(a, b, fma), vo.l2sq(a, b, fma), vo.metric_value(vo.COSINE, a, b, fma))
    vo.force_scalar(False)
    assert fast == slow


def test_accumulation_tree_is_the_documented_one():
    # element i -> P[(i/8)%4][i%8]; C[j]=(P0+P1)+(P2+P3); hsum8 = ((l0+l4)+(l2+l6))+((l1+l5)+(l3+l7))
    a, b = sinvec(768, 0.0), sinvec(768, 1.0)
    P = np.zeros(32, F)
    for i in range(768):
        # fused multiply-add in float64 then round once == fmaf for these magnitudes
        P[i % 32] = F(np.float64(a[i]) * np.float64(b[i]) + np.float64(P[i % 32]))
    Cj = [F(F(P[j] + P[8 + j]) + F(P[16 + j] + P[24 + j])) for j in range(8)]
    r = F(F(F(Cj[0] + Cj[4]) + F(Cj[2] + Cj[6])) + F(F(Cj[1] + Cj[5]) + F(Cj[3] + Cj[7])))
    assert vo.dot(a, b, True) == float(r)


def test_transform_score():
    # native/backend_adapter_tests.rs:90-118
    assert abs(vo.transform_score(vo.EUCLIDEAN, 0.5) - 0.5) < 1.2e-7
    assert abs(vo.transform_score(vo.COSINE, 0.3) - 0.7) < 1.2e-7
    assert vo.transform_score(vo.COSINE, 1.5) == 0.0
    assert vo.transform_score(vo.DOT, 0.5) == -0.5


def test_ef_search_rules():
    # params.rs:309-319 and params_tests.rs
    assert vo.ef_search(vo.FAST, 10) == 64 and vo.ef_search(vo.FAST, 100) == 200
    assert vo.ef_search(vo.BALANCED, 10) == 128 and vo.ef_search(vo.BALANCED, 40) == 160
    assert vo.ef_search(vo.ACCURATE, 10) == 512 and vo.ef_search(vo.ACCURATE, 100) == 1600
    assert vo.ef_search(vo.PERFECT, 10) == 4096 and vo.ef_search(vo.PERFECT, 100) == 10000
    assert vo.ef_search(vo.CUSTOM, 10, 64) == 64 and vo.ef_search(vo.CUSTOM, 100, 64) == 100


def test_random_layer_sequence_properties():
    # graph.rs:368-403: xorshift64(13,7,17) seeded 0x5DEECE66D1A4B5B5, level = floor(-ln(u)/ln M), cap 15
    lv = vo.levels(32, 200000)
    assert lv.max() <= 15
    frac0 = float((lv == 0).mean())
    assert abs(frac0 - (1 - 1 / 32)) < 0.003
    # first value by hand
    s = 0x5DEECE66D1A4B5B5
    s ^= (s << 13) & 0xFFFFFFFFFFFFFFFF
    s ^= s >> 7
    s ^= (s << 17) & 0xFFFFFFFFFFFFFFFF
    u = s / 2.0**64
    assert lv[0] == min(15, math.floor(-math.log(u) / math.log(32)))


def test_insert_and_search_ramp():
    # native/graph_tests.rs:10-30
    g = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100)
    for i in range(100):
        g.insert(np.array([i * 32 + j for j in range(32)], F))
    assert len(g) == 100
    ids, d = g.search(np.arange(32, dtype=F), 10, 50)
    assert 0 < len(ids) <= 10 and ids[0] == 0


def test_native_hnsw_basic_and_alpha_known_answers():
    # native/tests.rs:10-30: 100 x 128 sin vectors, M 16, ef_c 100; search(query of node 0, 10, 50)
    g = vo.Hnsw(vo.COSINE, 128, M=16, ef_construction=100)
    for i in range(100):
        g.insert(np.array([math.sin((i + j) * 0.01) for j in range(128)], F))
    assert len(g) == 100
    ids, d = g.search(np.array([math.sin(j * 0.01) for j in range(128)], F), 10, 50)
    assert len(ids) == 10 and d[0] < 0.1 and int(ids[0]) == 0
    # native/tests.rs:136-178: alpha = 1.2 (VAMANA-style pruning), two clusters of 25, the search still answers
    ga = vo.Hnsw(vo.COSINE, 32, M=16, ef_construction=100, alpha=1.2)
    for hot in (0, 1):
        for i in range(25):
            ga.insert(np.array([1.0 if j == hot else (i + j) * 0.001 for j in range(32)], F))
    assert len(ga) == 50
    ids, d = ga.search(np.array([0.9 if j == 0 else 0.01 for j in range(32)], F), 5, 50)
    assert len(ids) == 5 and all(int(i) < 25 for i in ids)          # the first cluster
    # native/tests.rs:193-214: alpha changes the links, not the count; alpha = 1.0 is the default (:181-190)
    g1, g2 = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100), vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100, alpha=1.2)
    for i in range(30):
        v = np.array([(i + j) * 0.1 for j in range(32)], F)
        g1.insert(v)
        g2.insert(v)
    assert len(g1) == len(g2) == 30
    same = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100, alpha=1.0)
    for i in range(30):
        same.insert(np.array([(i + j) * 0.1 for j in range(32)], F))
    assert all(np.array_equal(a[1], b[1]) for a, b in zip(g1.export_graph(), same.export_graph()))


def test_empty_search():
    # native/graph_tests.rs:33-41
    g = vo.Hnsw(vo.COSINE, 3, M=16, ef_construction=100)
    ids, d = g.search(np.array([1, 2, 3], F), 10, 50)
    assert len(ids) == 0


def test_select_neighbors_quota():
    # native/graph_tests.rs:44-160: empty -> empty; <= max -> all; heuristic back-fills to max
    g = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100)
    for i in range(20):
        g.insert(np.full(32, float(i), F))
    assert len(g.select_neighbors([], [], 10)) == 0
    q = np.zeros(32, F)
    cand = [(i, vo.graph_distance(vo.EUCLIDEAN, q, np.full(32, float(i), F))) for i in range(1, 6)]
    assert list(g.select_neighbors([c[0] for c in cand], [c[1] for c in cand], 10)) == [1, 2, 3, 4, 5]
    cand = [(i, vo.graph_distance(vo.EUCLIDEAN, q, np.full(32, float(i), F))) for i in range(1, 20)]
    sel = g.select_neighbors([c[0] for c in cand], [c[1] for c in cand], 8)
    assert len(sel) == 8 and sel[0] == 1  # collinear points: only the first is diverse, rest back-filled closest-first
    assert list(sel) == list(range(1, 9))


def test_recall_with_heuristic_selection():
    # native/graph_tests.rs:169-199
    g = vo.Hnsw(vo.COSINE, 128, M=32, ef_construction=200)
    for i in range(500):
        g.insert(np.array([math.sin(F((i * 127 + j)) * F(0.01)) for j in range(128)], F))
    q = np.array([math.sin(F(j) * F(0.01)) for j in range(128)], F)
    ids, d = g.search(q, 10, 100)
    assert len(ids) >= 5
    assert all(d[i] >= d[i - 1] for i in range(1, len(d)))


def test_native_hnsw_recall():
    # native/tests.rs:32-91: recall@10 >= 0.8 on 200x128 sin data, ef=128
    vecs = np.array([[math.sin(F(i * 128 + j) * F(0.001)) for j in range(128)] for i in range(200)], F)
    g = vo.Hnsw(vo.COSINE, 128, M=16, ef_construction=100)
    g.insert_many(vecs)
    tot = 0.0
    for qi in range(5):
        q = vecs[qi * 40]
        ids, _ = g.search(q, 10, 128)
        sims = vecs @ q / (np.linalg.norm(vecs, axis=1) * np.linalg.norm(q))
        gt = set(np.argsort(-sims, kind="stable")[:10].tolist())
        tot += len(gt & set(int(x) for x in ids)) / 10
    assert tot / 5 >= 0.8


def test_recall_quality_minimum_threshold():
    # index_tests.rs:1108-1159: 500x64, HnswParams::auto(64) = M 24 / ef_c 300, Accurate -> ef 512
    dim, n, k = 64, 500, 10
    data = np.array([[math.sin(F(i * dim + j) * F(0.001)) for j in range(dim)] for i in range(n)], F)
    g = vo.Hnsw(vo.COSINE, dim, M=24, ef_construction=300)
    g.insert_many(data)
    q = np.array([math.sin(F(j) * F(0.001)) for j in range(dim)], F)
    ids, _ = g.search(q, k, vo.ef_search(vo.ACCURATE, k))
    gt, _ = vo.bruteforce(vo.COSINE, data, q, k)
    assert len(set(ids.tolist()) & set(gt.tolist())) / k >= 0.8


def test_cpu_vs_simd_top1():
    # native/tests.rs:105-130 -- SIMD engine finds node 0 first on the (i+j) ramp
    g = vo.Hnsw(vo.EUCLIDEAN, 64, M=16, ef_construction=100)
    for i in range(50):
        g.insert(np.array([i + j for j in range(64)], F))
    ids, _ = g.search(np.arange(64, dtype=F), 5, 30)
    assert ids[0] == 0


def test_bruteforce_top1_at_origin():
    # index_tests.rs:1691-1712
    rng = np.random.default_rng(0)
    data = rng.normal(size=(50, 16)).astype(F)
    data[0] = 0
    ids, sc = vo.bruteforce(vo.EUCLIDEAN, data, np.zeros(16, F), 5)
    assert ids[0] == 0 and sc[0] == 0.0


def test_file_dump_and_load_roundtrip():
    # native/backend_adapter_tests.rs:121-173 + format at backend_adapter.rs:184-261
    g = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100)
    vecs = np.array([[F((i * 32 + j)) * F(0.01) for j in range(32)] for i in range(30)], F)
    g.insert_many(vecs)
    with tempfile.TemporaryDirectory() as d:
        g.dump(d, "roundtrip")
        raw = open(os.path.join(d, "roundtrip.vectors"), "rb").read()
        assert raw[:4] == (1).to_bytes(4, "little") and raw[4:12] == (30).to_bytes(8, "little")
        assert raw[12:16] == (32).to_bytes(4, "little") and len(raw) == 16 + 30 * 32 * 4
        graw = open(os.path.join(d, "roundtrip.graph"), "rb").read()
        hdr = np.frombuffer(graw[:20], dtype="<u4")
        assert hdr[0] == 1 and hdr[2] == 16 and hdr[3] == 32 and hdr[4] == 100
        h = vo.Hnsw.load(d, vo.EUCLIDEAN, "roundtrip")
        assert len(h) == 30
        a = g.search(vecs[0], 5, 50)
        b = h.search(vecs[0], 5, 50)
        assert a[0].tolist() == b[0].tolist() and a[1].tolist() == b[1].tolist() and a[0][0] == 0


def test_tokenize():
    # bm25_tests.rs:75-102
    assert vo.tokenize("Hello World") == ["hello", "world"]
    assert vo.tokenize("Hello, World! How are you?") == ["hello", "world", "how", "are", "you"]
    t = vo.tokenize("I am a test")
    assert "i" not in t and "a" not in t and "am" in t and "test" in t
    assert vo.tokenize("") == []


def test_bm25_known_answers():
    # bm25_tests.rs:109-122, 160-179, 181-195, 201-236, 242-281
    ix = vo.Bm25()
    ix.add_document(1, "rust programming language")
    ix.add_document(2, "python programming language")
    ix.add_document(3, "rust is fast")
    ids, sc = ix.search("rust", 10)
    assert sorted(ids.tolist()) == [1, 3]
    assert len(vo.Bm25().search("rust", 10)[0]) == 0
    ix = vo.Bm25()
    for i in range(1, 101):
        ix.add_document(i, f"document number {i} about rust")
    assert len(ix.search("rust", 5)[0]) == 5
    ix = vo.Bm25()
    ix.add_document(1, "rust")
    ix.add_document(2, "rust rust")
    ix.add_document(3, "rust rust rust")
    _, sc = ix.search("rust", 10)
    assert all(sc[i] >= sc[i + 1] for i in range(len(sc) - 1))
    ix = vo.Bm25()
    ix.add_document(1, "rust programming")
    ix.add_document(2, "python programming")
    ix.add_document(3, "java programming")
    assert len(ix.search("rust", 10)[0]) == 1 and len(ix.search("programming", 10)[0]) == 3
    ix = vo.Bm25()
    ix.add_document(1, "rust")
    ix.add_document(2, "rust is a systems programming language that runs blazingly fast")
    ids, _ = ix.search("rust", 10)
    assert len(ids) == 2 and ids[0] == 1
    ix = vo.Bm25()
    ix.add_document(1, "hello@world.com is an email")
    assert len(ix.search("hello", 10)[0]) == 1 and len(ix.search("world", 10)[0]) == 1
    ix = vo.Bm25()
    ix.add_document(1, "version 2.0 released in 2024")
    assert len(ix.search("2024", 10)[0]) == 1
    ix = vo.Bm25()
    ix.add_document(1, "café résumé naïve")
    assert len(ix.search("café", 10)[0]) == 1
    ix = vo.Bm25()
    ix.add_document(1, "rust programming")
    ids, sc = ix.search("rust rust rust", 10)
    assert len(ids) == 1
    one = ix.search("rust", 10)[1][0]
    assert sc[0] == F(F(one + one) + one)  # duplicates are scored again, in query order
    ix = vo.Bm25()
    ix.add_document(1, "original text")
    ix.add_document(1, "updated text")
    assert len(ix) == 1


def test_bm25_bookkeeping_known_answers():
    # bm25_tests.rs:8-62, 123-160, 316-345: sizes, removal, no-match / empty queries, the u32 id limit -- on the oracle
    # and on the host mirror's CPU-side bookkeeping (no device call is made before the first non-empty search)
    from velesdb_b200 import Bm25Index

    for make in (vo.Bm25, Bm25Index):
        ix = make()
        assert len(ix) == 0 and ix.term_count() == 0
        ix.add_document(1, "hello world")
        assert len(ix) == 1 and ix.term_count() >= 2
        ix.add_document(2, "goodbye world")
        assert len(ix) == 2 and ix.remove_document(1) and len(ix) == 1 and not ix.remove_document(1)
        ix = make()
        for d, t in ((1, "rust programming language"), (2, "python programming language"), (3, "java programming")):
            ix.add_document(d, t)
        assert len(ix) == 3
        ix = make()
        ix.add_document(0xFFFFFFFF, "test document")                       # exactly u32::MAX is allowed
        assert len(ix) == 1
        with pytest.raises((AssertionError, ValueError), match="BM25 document ID"):   # the reference panics
            make().add_document(0xFFFFFFFF + 1, "test document")
    ix = vo.Bm25()
    for d, t in ((1, "rust programming language fast"), (2, "python programming language"), (3, "rust systems programming")):
        ix.add_document(d, t)
    ids, _ = ix.search("rust programming", 10)
    assert {1, 3} <= set(ids.tolist()) and len(ids) == 3
    ix = vo.Bm25()
    ix.add_document(1, "rust programming")
    ix.add_document(2, "python programming")
    assert len(ix.search("javascript", 10)[0]) == 0 and len(ix.search("", 10)[0]) == 0
    assert Bm25Index().search("rust", 10) == [] and Bm25Index().search("", 10) == []


def test_bm25_formula_by_hand():
    # bm25.rs:290-376 with k1=1.2, b=0.75
    ix = vo.Bm25()
    ix.add_document(1, "rust")
    ix.add_document(2, "rust is a systems programming language that runs blazingly fast")
    n, df = F(2), F(2)
    idf = F(math.log(F(F(F(n - df + F(0.5)) / F(df + F(0.5))) + F(1.0))))
    avgdl = F(F(1 + 9) / F(2))  # "a" is dropped: 9 tokens
    for doc, dl in ((1, 1), (2, 9)):
        len_norm = F(F(F(1.0) - F(0.75)) + F(F(F(0.75) * F(dl)) / avgdl))
        s = F(F(idf * F(F(1) * F(F(1.2) + F(1.0)))) / F(F(1) + F(F(1.2) * len_norm)))
        ids, sc = ix.search("rust", 10)
        got = sc[list(ids).index(doc)]
        assert abs(got - s) <= 1.2e-7 * abs(s) * 2


def sample_results():
    return [
        [(1, 0.95), (2, 0.85), (3, 0.75), (4, 0.65)],
        [(2, 0.90), (1, 0.80), (5, 0.70), (3, 0.60)],
        [(1, 0.92), (3, 0.82), (2, 0.72), (6, 0.62)],
    ]


def test_fusion_known_answers():
    # fusion/strategy_tests.rs:54-272
    ids, sc = vo.fuse(vo.AVERAGE, sample_results())
    assert ids[0] == 1 and abs(sc[0] - (0.95 + 0.80 + 0.92) / 3) < 1e-3
    ids, sc = vo.fuse(vo.MAXIMUM, sample_results())
    assert len(ids) == 6 and ids[0] == 1 and abs(sc[0] - 0.95) < 1e-3
    ids, sc = vo.fuse(vo.RRF, sample_results(), rrf_k=60)
    d1 = sc[list(ids).index(1)]
    assert d1 > 0.04 and abs(d1 - (1 / 61 + 1 / 62 + 1 / 61)) < 1e-6
    assert all(sc[i - 1] >= sc[i] for i in range(1, len(sc)))
    lo = vo.fuse(vo.RRF, sample_results(), rrf_k=1)
    hi = vo.fuse(vo.RRF, sample_results(), rrf_k=100)
    assert lo[1][list(lo[0]).index(1)] > hi[1][list(hi[0]).index(1)]
    ids, sc = vo.fuse(vo.RRF, [[(1, 0.95), (2, 0.85), (3, 0.75)]])
    assert len(ids) == 3 and sc[0] > sc[1]
    assert len(vo.fuse(vo.RRF, [])[0]) == 0
    assert len(vo.fuse(vo.AVERAGE, [[], []])[0]) == 0
    ids, sc = vo.fuse(vo.WEIGHTED, sample_results(), avg_w=0.6, max_w=0.3, hit_w=0.1)
    assert ids[0] == 1


def test_hybrid_rrf_formula():
    # collection/search/text.rs:133-180: w/(rank0+60) + (1-w)/(rank0+60); top-k keeps the largest (score, id)
    ids, sc = vo.rrf_hybrid([5, 7, 9], [7, 11], k=3, vector_weight=0.5)
    exp = {5: F(0.5) / F(60), 7: F(F(0.5) / F(61)) + F(F(0.5) / F(60)), 9: F(0.5) / F(62), 11: F(0.5) / F(61)}
    assert ids[0] == 7 and sc[0] == F(exp[7])
    assert ids.tolist() == [7, 5, 11] and sc[1] == exp[5] and sc[2] == exp[11]
    # weight is clamped to [0,1]
    ids, sc = vo.rrf_hybrid([1], [2], k=2, vector_weight=7.0)
    assert ids.tolist() == [1, 2] and sc[1] == 0.0


# collection/tests.rs:298-336, 394-446: the reference's own answers for Collection::hybrid_search.  (points, query
# vector, text query, k, vector_weight, id that must come first); a point's text is its payload strings joined by " "
# (collection/types.rs:169-191)
HYBRID_KNOWN_ANSWERS = [
    ([(1, [1.0, 0.0, 0.0], "Rust Programming"), (2, [0.9, 0.1, 0.0], "Python Programming"), (3, [0.0, 1.0, 0.0], "Rust Performance")],
     [1.0, 0.0, 0.0], "rust", 3, 0.5, 1),                                         # test_collection_hybrid_search
    ([(1, [1.0, 0.0, 0.0], "Rust"), (2, [0.9, 0.1, 0.0], "Python")], [0.9, 0.1, 0.0], "rust", 2, 1.0, 2),   # ..._text_weight_zero
    ([(1, [1.0, 0.0, 0.0], "Rust programming language"), (2, [0.99, 0.01, 0.0], "Python programming")],
     [0.99, 0.01, 0.0], "rust", 2, 0.0, 1),                                       # ..._vector_weight_zero
]


@pytest.mark.parametrize("points,query,text,k,w,first", HYBRID_KNOWN_ANSWERS)
def test_hybrid_search_reference_known_answers(points, query, text, k, w, first):
    # Collection::hybrid_search (collection/search/text.rs:113-180) composed from the oracle's parts: index.search(q, 2k)
    # with ef_search(Balanced, 2k), text_index.search(text, 2k), the RRF
    g, bm = vo.Hnsw(vo.COSINE, 3, M=16, ef_construction=100), vo.Bm25()
    for pid, vec, txt in points:
        g.insert(np.asarray(vec, np.float32))
        bm.add_document(pid, txt)
    ef = vo.ef_search(vo.BALANCED, 2 * k)
    nodes, _ = g.search(np.asarray(query, np.float32), 2 * k, ef)
    vec_ids = [points[int(i)][0] for i in nodes]
    txt_ids, _ = bm.search(text, 2 * k)
    ids, sc = vo.rrf_hybrid(vec_ids, txt_ids, k, w)
    assert len(ids) >= 1 and int(ids[0]) == first
    assert all(sc[i] >= sc[i + 1] for i in range(len(sc) - 1))


# ---- SQ8 dual precision (native/quantization_tests.rs, native/dual_precision_tests.rs) ----
def test_sq8_quantizer_known_answers():
    # quantization_tests.rs:13-29 train min / scale
    q = vo.ScalarQuantizer([[0.0, 10.0, -5.0], [5.0, 20.0, 5.0], [2.5, 15.0, 0.0]])
    assert q.dimension == 3
    assert np.allclose(q.min_vals, [0.0, 10.0, -5.0], atol=1e-6)
    assert np.allclose(q.scales, [255.0 / 5.0, 255.0 / 10.0, 255.0 / 10.0], atol=1e-4)
    # :32-41 constant dimensions get scale 1
    q = vo.ScalarQuantizer([[1.0, 5.0, 5.0], [2.0, 5.0, 5.0]])
    assert q.scales[1] == 1.0 and q.scales[2] == 1.0
    # :44-48 empty training set
    with pytest.raises(ValueError, match="Cannot train on empty vectors"):
        vo.ScalarQuantizer(np.zeros((0, 3), F))
    # :54-62, :66-85 range mapping, :88-100 clamping
    q = vo.ScalarQuantizer([[0.0, 100.0]])
    assert q.quantize([0.0, 100.0])[0] == 0
    q = vo.ScalarQuantizer([[0.0, 0.0], [10.0, 100.0]])
    assert q.quantize([0.0, 0.0]).tolist() == [0, 0] and q.quantize([10.0, 100.0]).tolist() == [255, 255]
    assert all(abs(int(v) - 127) <= 1 for v in q.quantize([5.0, 50.0]))
    q = vo.ScalarQuantizer([[0.0], [10.0]])
    assert q.quantize([-5.0])[0] == 0 and q.quantize([20.0])[0] == 255
    assert q.quantize([float("nan")])[0] == 0  # NaN survives clamp, `as u8` saturates it to 0
    # :103-124 dequantize within 1% of the range
    q = vo.ScalarQuantizer([[0.0, -10.0, 100.0], [10.0, 10.0, 200.0]])
    orig = np.array([5.0, 0.0, 150.0], F)
    rec = q.dequantize(q.quantize(orig))
    assert (np.abs(orig - rec) / np.array([10.0, 20.0, 100.0]) < 0.01).all()


def test_sq8_distances_known_answers():
    # quantization_tests.rs:129-147 identical -> 0, symmetry
    q = vo.ScalarQuantizer([[0.0, 0.0], [10.0, 10.0]])
    v = q.quantize([5.0, 5.0])
    assert q.distance_l2_quantized(v, v) == 0
    a, b = q.quantize([2.0, 3.0]), q.quantize([7.0, 8.0])
    assert q.distance_l2_quantized(a, b) == q.distance_l2_quantized(b, a) > 0
    # :150-177 asymmetric distance within 5% of exact
    q = vo.ScalarQuantizer([np.zeros(128, F), np.full(128, 10.0, F)])
    approx = q.distance_l2_asymmetric(np.full(128, 3.0, F), q.quantize(np.full(128, 7.0, F)))
    exact = math.sqrt(128 * 16.0)
    assert abs(approx - exact) / exact < 0.05
    # u32 sum against a plain integer sum, 768 codes, extreme values
    rng = np.random.default_rng(3)
    x, y = rng.integers(0, 256, 771, dtype=np.uint8), rng.integers(0, 256, 771, dtype=np.uint8)
    assert q.distance_l2_quantized(x, y) == int(((x.astype(np.int64) - y.astype(np.int64)) ** 2).sum())
    assert q.distance_l2_quantized(np.zeros(768, np.uint8), np.full(768, 255, np.uint8)) == 768 * 65025
    # :246-269 768-d reconstruction error
    v1 = np.array([math.sin(i * 0.01) for i in range(768)], F)
    v2 = np.array([math.cos(i * 0.01) for i in range(768)], F)
    q = vo.ScalarQuantizer([v1, v2])
    assert float(((v1 - q.dequantize(q.quantize(v1))) ** 2).mean()) < 1e-3


def _ramp(i, dim):
    return np.arange(i * dim, (i + 1) * dim, dtype=F)


def test_dual_precision_training_lifecycle():
    # dual_precision_tests.rs:12-77
    dp = vo.DualPrecisionHnsw(vo.EUCLIDEAN, 32, 16, 100, 1000)
    assert len(dp) == 0 and not dp.is_quantizer_trained()
    for i in range(10):
        dp.insert(_ramp(i, 32))
    assert len(dp) == 10 and not dp.is_quantizer_trained()
    dp.force_train_quantizer()
    assert dp.is_quantizer_trained() and len(dp.quantizer) == 10
    dp = vo.DualPrecisionHnsw(vo.EUCLIDEAN, 32, 16, 100, 100)  # training_sample_size = min(1000, 100)
    for i in range(100):
        dp.insert(np.array([math.sin((i * 32 + j) * 0.01) for j in range(32)], F))
    assert dp.is_quantizer_trained() and len(dp.quantizer) == 100
    dp.insert(np.zeros(32, F))  # :162-190 inserts after training are quantized too
    assert len(dp.quantizer) == 101


def test_dual_precision_search_known_answers():
    # dual_precision_tests.rs:80-120: ramp vectors, nearest to the first ramp is node 0 before and after training
    dp = vo.DualPrecisionHnsw(vo.EUCLIDEAN, 32, 16, 100, 1000)
    for i in range(100):
        dp.insert(_ramp(i, 32))
    q = _ramp(0, 32)
    ids, _ = dp.search(q, 10, 50)
    assert len(ids) > 0 and ids[0] == 0
    dp.force_train_quantizer()
    ids, d = dp.search(q, 10, 50)
    assert ids[0] == 0 and (np.diff(d) >= 0).all()
    # :260-288 int8 traversal returns sorted results
    dp = vo.DualPrecisionHnsw(vo.EUCLIDEAN, 64, 16, 100, 500)
    for i in range(200):
        dp.insert(np.array([math.sin((i * 64 + j) * 0.01) for j in range(64)], F))
    dp.force_train_quantizer()
    q = np.array([math.sin(j * 0.01) for j in range(64)], F)
    ids, d = dp.search_with_config(q, 10, 50, min_index_size=0)
    assert len(ids) > 0 and (np.diff(d) >= 0).all()
    # the size gate (:275-277): below min_index_size the f32 path answers
    a = dp.search_with_config(q, 10, 50)
    b = dp.inner.search(q, 10, 50)
    assert np.array_equal(a[0], b[0])


def test_dual_precision_int8_recall_vs_f32():
    # dual_precision_tests.rs:224-257, 291-334: 500 x 128 cos vectors, query = vector 0
    dp = vo.DualPrecisionHnsw(vo.EUCLIDEAN, 128, 32, 200, 1000)
    vs = np.array([[math.cos((i * 128 + j) * 0.001) for j in range(128)] for i in range(500)], F)
    for v in vs:
        dp.insert(v)
    dp.force_train_quantizer()
    f_ids, _ = dp.search(vs[0], 10, 100)
    assert 0 in f_ids.tolist()
    i_ids, i_d = dp.search_with_config(vs[0], 10, 100, oversampling_ratio=4, min_index_size=0)
    assert len(set(f_ids.tolist()) & set(i_ids.tolist())) / max(len(f_ids), 1) >= 0.9
    # canonical and reference coarse orders agree when no tie is flagged
    r = dp.search_with_config(vs[0], 10, 100, min_index_size=0, order="canonical", with_stats=True)
    if not r[2]["tie_at_k"]:
        assert np.array_equal(r[0], i_ids)


def test_search_multi_entry_known_answers():
    # native/tests.rs:217-256
    g = vo.Hnsw(vo.COSINE, 32, M=16, ef_construction=100)
    for i in range(50):
        g.insert(np.array([math.sin((i + j) * 0.01) for j in range(32)], F))
    q = np.array([math.sin(j * 0.01) for j in range(32)], F)
    s0 = g.rng_state
    ids, d, st, ent = g.search_multi_entry(q, 5, 50, 3)
    assert 0 < len(ids) <= 5 and (np.diff(d) >= 0).all()
    assert 1 <= len(ent) <= 3 and g.rng_state != s0          # two draws advanced the shared state
    from velesdb_b200 import multi_entry_probes
    row, s2 = multi_entry_probes(s0, 50, 3)
    assert s2 == g.rng_state and set(ent[1:]) <= set(row)      # the host helper follows the same stream
    g2 = vo.Hnsw(vo.EUCLIDEAN, 32, M=16, ef_construction=100)
    for i in range(30):
        g2.insert(np.array([(i + j) * 0.1 for j in range(32)], F))
    q2 = np.array([j * 0.05 for j in range(32)], F)
    std_ids, _ = g2.search(q2, 5, 50)
    mi, _, _, _ = g2.search_multi_entry(q2, 5, 50, 2)
    assert len(std_ids) > 0 and len(mi) > 0 and mi[0] == std_ids[0]   # ef covers the whole graph: same best hit
    # count <= 10 or a single probe: no draw, same as the standard search
    g3 = vo.Hnsw(vo.EUCLIDEAN, 4, M=8, ef_construction=20)
    for i in range(8):
        g3.insert(np.array([i, 0, 0, 1], F))
    s3 = g3.rng_state
    a, _, _, ent3 = g3.search_multi_entry(np.array([2.2, 0, 0, 1], F), 3, 16, 4)
    assert g3.rng_state == s3 and len(ent3) == 1 and a.tolist() == g3.search(np.array([2.2, 0, 0, 1], F), 3, 16)[0].tolist()
