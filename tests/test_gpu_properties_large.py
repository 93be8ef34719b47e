"""Size-independent properties of the search path on a larger index (200K x 128, built by the GPU bulk
builder): the checks the domain offers when an oracle build would be too slow -- sortedness, determinism,
batch == single, self-queries, recall against exact brute force -- plus oracle parity on a query sample
over the exported graph."""
import numpy as np
import pytest

from oracle import oracle as vo
from tests.gpu_util import bits_equal
from velesdb_b200 import DeviceSnapshot, DistanceMetric

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big():
    import torch

    from bench import gen_data

    dev = torch.device("cuda", 0)
    x = gen_data(torch, 200_000, 128, 16, 5, dev).cpu().numpy()
    q = gen_data(torch, 512, 128, 16, 99, dev).cpu().numpy()
    snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
    snap.build_graph(16)
    return x, q, snap


def test_results_sorted_padded_and_deterministic(big):
    x, q, snap = big
    ids, dist, cnt, st = snap.search_batch(q, 10, 64, with_stats=True)
    assert (cnt == 10).all()
    dd = np.diff(dist, axis=1)
    assert (dd >= 0).all()                                # ascending by distance ...
    assert (np.diff(ids.astype(np.int64), axis=1)[dd == 0] > 0).all()   # ... ties by ascending node id
    ids2, dist2, cnt2, st2 = snap.search_batch(q, 10, 64, with_stats=True)
    assert np.array_equal(ids, ids2) and bits_equal(dist, dist2) and np.array_equal(st, st2)
    # a batch is the concatenation of its single-query searches
    for i in (0, 17, 511):
        i1, d1, c1 = snap.search_batch(q[i], 10, 64)
        assert np.array_equal(i1[0], ids[i]) and bits_equal(d1[0], dist[i])
    # k > ef: at most ef results (graph.rs:266-269)
    ids3, _, cnt3 = snap.search_batch(q[:8], 100, 32)
    assert (cnt3 <= 32).all() and (ids3[:, 32:] == 0xFFFFFFFF).all()


def test_self_queries_and_recall(big):
    x, q, snap = big
    ids, dist, cnt = snap.search_batch(x[:256], 1, 64)
    assert (ids[:, 0] == np.arange(256)).mean() >= 0.99 and (dist[:, 0] < 1e-6).all()
    ids, dist, cnt = snap.search_batch(q, 10, 128)
    bi, bs = snap.bruteforce_batch(q, 10)
    rec = np.mean([len(set(ids[i].tolist()) & set(bi[i].tolist())) / 10 for i in range(len(q))])
    assert rec >= 0.95, rec
    # brute force is exact: its scores are the sorted maxima of the full similarity matrix
    sims = q[:16] @ x.T
    top = -np.sort(-sims, axis=1)[:, :10]
    assert np.allclose(bs[:16], top, atol=2e-6)


def test_oracle_parity_on_a_sample_over_the_exported_graph(big):
    x, q, snap = big
    g = vo.Hnsw.from_arrays(vo.COSINE, x, snap.export_graph(), 16, 32, snap.entry_point, snap.max_layer)
    ids, dist, cnt, st = snap.search_batch(q[:96], 10, 64, with_stats=True)
    oi, od, oc, ost = g.search_batch(q[:96], 10, 64, order="canonical", threads=8)
    keep = ost[:, 4] == 0
    assert np.array_equal(ids[keep], oi[keep].astype(np.uint32)) and bits_equal(dist, od)
    assert np.array_equal(st[:, 0], ost[:, 0].astype(np.uint32)) and np.array_equal(st[:, 1], ost[:, 1].astype(np.uint32))
