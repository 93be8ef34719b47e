"""The committed fixtures of tests/golden/ through the C ABI on the GPU (tests/golden/make_golden.py documents them)."""
import json
import os

import numpy as np
import pytest

from velesdb_b200 import DeviceSnapshot, DistanceMetric

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_spec_v1_files_on_gpu():
    exp = json.load(open(os.path.join(GOLD, "spec_v1_expected.json")))
    q = np.array(exp["queries"], np.float32)
    for name, metric in (("euclidean", DistanceMetric.Euclidean), ("cosine", DistanceMetric.Cosine)):
        snap = DeviceSnapshot.from_reference_files(GOLD, metric, basename="spec_v1")
        assert len(snap) == exp["n"] and snap.dim == exp["dim"] and snap.entry_point == exp["entry_point"]
        ids, dist, cnt = snap.search_batch(q, exp["k"], exp["ef"])
        assert (cnt == exp["k"]).all() and ids.tolist() == exp["expected_ids"][name]
        bi, _ = snap.bruteforce_batch(q, exp["k"])
        assert bi.tolist() == exp["expected_ids"][name]


def test_regression_fixture_on_gpu():
    z = np.load(os.path.join(GOLD, "regress_cos24.npz"))
    snap = DeviceSnapshot.from_reference_files(GOLD, DistanceMetric.Cosine, basename="regress_cos24")
    ids, dist, cnt, st = snap.search_batch(z["queries"], 5, 32, with_stats=True)
    assert np.array_equal(ids, z["ids"].astype(np.uint32)) and np.array_equal(dist.view(np.uint32), z["dist_bits"])
    assert np.array_equal(cnt, z["counts"]) and np.array_equal(st, z["stats"].astype(np.uint32))
    snap.attach_sq8(200)
    mn, sc, _, codes = snap.sq8_export()
    assert np.array_equal(mn, z["sq8_min"]) and np.array_equal(sc, z["sq8_scale"])
    assert int(codes.astype(np.uint64).sum()) == int(z["sq8_codes_crc"][0])
    sids, sd, scnt, sst = snap.search_batch_sq8(z["queries"], 5, 32, 4, with_stats=True)
    assert np.array_equal(sids, z["sq8_ids"].astype(np.uint32)) and np.array_equal(sd.view(np.uint32), z["sq8_dist_bits"])
    assert np.array_equal(sst, z["sq8_stats"].astype(np.uint32))
    bi, bs = snap.bruteforce_batch(z["queries"], 5)
    assert np.array_equal(bi, z["bf_ids"].astype(np.uint32)) and np.array_equal(bs.view(np.uint32), z["bf_score_bits"])
