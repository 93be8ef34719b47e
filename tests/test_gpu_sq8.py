"""GPU parity for the SQ8 dual-precision path (native/dual_precision.rs, native/quantization.rs): quantizer,
codes, int8 traversal and exact re-rank through the C ABI against the CPU oracle.  Integer work: bit-exact;
the re-ranked f32 distances: 0 ulp."""
import math

import numpy as np
import pytest

from oracle import oracle as vo
from tests.gpu_util import bits_equal, build_oracle, latent_data, queries_near
from velesdb_b200 import (DeviceSnapshot, DistanceMetric, DualPrecisionConfig, DualPrecisionHnsw, VelesError)

pytestmark = pytest.mark.gpu


def make_case(metric, dim, n=2000, train=None, M=16, ef_c=100, seed=0, normalize=False):
    x = latent_data(n, dim, seed=seed + metric * 100 + dim, normalize=normalize)
    g = build_oracle(metric, x, M, ef_c)
    snap = DeviceSnapshot.from_arrays(x, metric, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    dp = vo.DualPrecisionHnsw.from_graph(g, train_count=train)
    snap.attach_sq8(min(1000, n) if train is None else train)
    return x, g, dp, snap


@pytest.mark.parametrize("dim", [1, 3, 16, 100, 768, 771])
def test_quantizer_and_codes_bit_exact(dim):
    # quantization.rs:190-250
    rng = np.random.default_rng(dim)
    x = rng.normal(size=(1500, dim)).astype(np.float32)
    if dim >= 3:
        x[:, 1] = 0.25          # constant dimension -> scale 1.0
        x[1200:, 2] *= 40.0     # out of the trained range -> clamped
    x[1300, 0] = np.nan         # NaN after training -> code 0
    snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Euclidean)
    snap.attach_sq8(1000)
    assert snap.has_sq8
    mn, sc, inv, codes = snap.sq8_export()
    q = vo.ScalarQuantizer(x[:1000])
    assert bits_equal(mn, q.min_vals) and bits_equal(sc, q.scales) and bits_equal(inv, q.inv_scales)
    assert np.array_equal(codes, q.quantize(x))
    if dim >= 3:
        assert sc[1] == 1.0 and codes[:, 1].max() == 0
        assert codes[1200:, 2].min() == 0 and codes[1200:, 2].max() == 255
    assert codes[1300, 0] == 0


def check_sq8(dp, snap, q, k, ef, over):
    ids, dist, cnt, st = snap.search_batch_sq8(q, k, ef, over, with_stats=True)
    oi, od, oc, ost = dp.search_int8_batch(q, k, ef, over, order="canonical", threads=8)
    assert np.array_equal(cnt, oc)
    for r in range(q.shape[0]):
        c = int(cnt[r])
        assert np.array_equal(ids[r, :c], oi[r, :c].astype(np.uint32)), (r, ids[r], oi[r])
        assert bits_equal(dist[r, :c], od[r, :c])
        assert (ids[r, c:] == 0xFFFFFFFF).all()
    # identical int8 traversal: same distance evaluations and expansions on every layer
    for j in range(4):
        assert np.array_equal(st[:, j], ost[:, j].astype(np.uint32)), j
    # the reference's own order (heap-array order among equal int8 distances) where no tie is flagged
    ri, rd, rc, rst = dp.search_int8_batch(q, k, ef, over, order="reference", threads=8)
    free = rst[:, 4] == 0
    assert free.sum() > 0
    assert np.array_equal(ids[free], ri[free].astype(np.uint32))
    assert bits_equal(dist[free], rd[free])


@pytest.mark.parametrize("metric", [vo.COSINE, vo.EUCLIDEAN, vo.DOT])
@pytest.mark.parametrize("dim", [20, 96, 768])
def test_sq8_search_bit_exact(metric, dim):
    # dual_precision.rs:284-441
    n = 2000 if dim < 768 else 1200
    x, g, dp, snap = make_case(metric, dim, n=n)
    q = queries_near(x, 64, seed=5)
    for k, ef, over in ((10, 64, 4), (1, 1, 1), (10, 16, 4), (50, 100, 4), (10, 300, 2), (3, 600, 8)):
        check_sq8(dp, snap, q, k, ef, over)


def test_sq8_search_many_ties():
    # coarse codes (few distinct values per dimension) -> equal int8 distances everywhere: the tie list
    rng = np.random.default_rng(11)
    x = rng.integers(0, 3, size=(1500, 24)).astype(np.float32)
    g = build_oracle(vo.EUCLIDEAN, x, 16, 100)
    snap = DeviceSnapshot.from_arrays(x, vo.EUCLIDEAN, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    dp = vo.DualPrecisionHnsw.from_graph(g)
    snap.attach_sq8(1000)
    q = rng.integers(0, 3, size=(64, 24)).astype(np.float32)
    ids, dist, cnt, st = snap.search_batch_sq8(q, 10, 32, 2, with_stats=True)
    oi, od, oc, ost = dp.search_int8_batch(q, 10, 32, 2, order="canonical", threads=8)
    assert np.array_equal(cnt, oc) and np.array_equal(ids, oi.astype(np.uint32)) and bits_equal(dist, od)
    for j in range(4):
        assert np.array_equal(st[:, j], ost[:, j].astype(np.uint32))


def test_sq8_train_count_and_errors():
    x, g, dp, snap = make_case(vo.EUCLIDEAN, 48, n=600, train=600)  # force_train_quantizer on 600 inserts
    q = queries_near(x, 16, seed=2)
    check_sq8(dp, snap, q, 5, 40, 4)
    plain = DeviceSnapshot.from_arrays(x, vo.EUCLIDEAN, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
    assert not plain.has_sq8
    with pytest.raises(VelesError, match="no SQ8 store"):
        plain.search_batch_sq8(q, 5, 40)
    with pytest.raises(VelesError, match="Cannot train on empty vectors"):
        plain.attach_sq8(0)
    with pytest.raises(VelesError, match="train_count"):
        plain.attach_sq8(601)
    with pytest.raises(VelesError, match="oversampling"):
        snap.search_batch_sq8(q, 2048, 40, 4)
    half = DeviceSnapshot.from_arrays(x, vo.EUCLIDEAN, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer,
                                      store_dtype="f16")
    with pytest.raises(VelesError, match="f32 snapshot"):
        half.attach_sq8(100)


def test_dual_precision_host_mirror_reference_tests():
    # dual_precision_tests.rs:12-120 through the host mirror (graph by the exact GPU builder)
    dp = DualPrecisionHnsw.new(DistanceMetric.Euclidean, 32, 16, 100, 1000)
    assert dp.is_empty() and not dp.is_quantizer_trained()
    assert dp.search(np.zeros(32, np.float32), 10, 50) == []
    for i in range(100):
        dp.insert(np.arange(i * 32, (i + 1) * 32, dtype=np.float32))
    assert dp.len() == 100 and not dp.is_quantizer_trained()
    q = np.arange(32, dtype=np.float32)
    r = dp.search(q, 10, 50)
    assert r and r[0][0] == 0
    dp.force_train_quantizer()
    assert dp.is_quantizer_trained() and dp.quantizer() is not None
    r = dp.search(q, 10, 50)
    assert r[0][0] == 0 and all(r[i][1] >= r[i - 1][1] for i in range(1, len(r)))
    # trains itself at the threshold (:36-54)
    dp2 = DualPrecisionHnsw.new(DistanceMetric.Euclidean, 32, 16, 100, 100)
    for i in range(100):
        dp2.insert(np.array([math.sin((i * 32 + j) * 0.01) for j in range(32)], np.float32))
    assert dp2.is_quantizer_trained()
    # defaults (:337-342)
    c = DualPrecisionConfig()
    assert (c.oversampling_ratio, c.use_int8_traversal, c.min_index_size) == (4, True, 10_000)


def test_dual_precision_host_mirror_int8_vs_oracle():
    # dual_precision_tests.rs:260-334: 500 x 128 cos vectors; the mirror builds the same graph as the oracle
    vs = np.array([[math.cos((i * 128 + j) * 0.001) for j in range(128)] for i in range(500)], np.float32)
    dp = DualPrecisionHnsw.new(DistanceMetric.Euclidean, 128, 32, 200, 1000)
    odp = vo.DualPrecisionHnsw(vo.EUCLIDEAN, 128, 32, 200, 1000)
    for v in vs:
        dp.insert(v)
        odp.insert(v)
    dp.force_train_quantizer()
    odp.force_train_quantizer()
    cfg = DualPrecisionConfig(oversampling_ratio=4, use_int8_traversal=True, min_index_size=0)
    f32 = dp.search(vs[0], 10, 100)
    i8 = dp.search_with_config(vs[0], 10, 100, cfg)
    assert any(i == 0 for i, _ in f32)
    assert len({i for i, _ in f32} & {i for i, _ in i8}) / max(len(f32), 1) >= 0.9
    assert all(i8[j][1] >= i8[j - 1][1] for j in range(1, len(i8)))
    oi, od, ost = odp.search_with_config(vs[0], 10, 100, min_index_size=0, order="canonical", with_stats=True)
    assert [i for i, _ in i8] == oi.tolist()
    assert bits_equal(np.array([d for _, d in i8], np.float32), od)
    # size gate: the default config (min_index_size 10_000) answers with the f32 path
    assert dp.search_with_config(vs[0], 10, 100) == f32
    # batched form
    ids, dist, cnt = dp.search_batch_with_config(vs[:8], 10, 100, cfg)
    for r in range(8):
        o = odp.search_with_config(vs[r], 10, 100, min_index_size=0, order="canonical")
        assert ids[r, :cnt[r]].tolist() == o[0].tolist()


def test_sq8_recall_close_to_f32():
    # normalised 128-d data, cosine index: int8 traversal + 4x oversampling keeps the f32 path's recall
    x, g, dp, snap = make_case(vo.COSINE, 128, n=4000, M=16, ef_c=100, normalize=True)
    q = queries_near(x, 128, jitter=0.05, seed=3)
    gt, _ = snap.bruteforce_batch(q, 10)
    f_ids, _, _ = snap.search_batch(q, 10, 64)
    s_ids, _, _ = snap.search_batch_sq8(q, 10, 64, 4)
    rec = lambda a: np.mean([len(set(a[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(len(q))])
    assert rec(s_ids) >= rec(f_ids) - 0.03, (rec(s_ids), rec(f_ids))


@pytest.mark.parametrize("warps", ["1", "2", "4", "8"])
def test_sq8_warps_per_query(warps, monkeypatch):
    monkeypatch.setenv("VELES_SEARCH_WARPS", warps)
    for metric, dim in ((vo.COSINE, 768), (vo.EUCLIDEAN, 20)):
        x, g, dp, snap = make_case(metric, dim, n=1200 if dim == 768 else 2000)
        q = queries_near(x, 64, seed=5)
        for k, ef, over in ((10, 64, 4), (50, 100, 4), (3, 600, 8)):
            check_sq8(dp, snap, q, k, ef, over)
