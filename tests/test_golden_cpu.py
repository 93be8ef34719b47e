"""The committed fixtures of tests/golden/ against the CPU oracle (see tests/golden/make_golden.py for what each
file is and where it comes from)."""
import json
import os

import numpy as np

from oracle import oracle as vo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_spec_v1_files_load_and_answer():
    # format v1 written from the spec (backend_adapter.rs:184-261); expected ids from a numpy brute force
    exp = json.load(open(os.path.join(GOLD, "spec_v1_expected.json")))
    for name, metric in (("euclidean", vo.EUCLIDEAN), ("cosine", vo.COSINE)):
        g = vo.Hnsw.load(GOLD, metric, basename="spec_v1")
        assert len(g) == exp["n"] and g.dim == exp["dim"] and g.entry_point == exp["entry_point"]
        assert g.max_layer == exp["max_layer"] and g.M == 4 and g.M0 == 8
        for q, want in zip(exp["queries"], exp["expected_ids"][name]):
            ids, d = g.search(np.array(q, np.float32), exp["k"], exp["ef"])
            assert ids.tolist() == want and (np.diff(d) >= 0).all()


def test_regression_fixture_matches_the_oracle():
    z = np.load(os.path.join(GOLD, "regress_cos24.npz"))
    g = vo.Hnsw.load(GOLD, vo.COSINE, basename="regress_cos24")
    assert len(g) == 300 and g.dim == 24
    ids, d, cnt, st = g.search_batch(z["queries"], 5, 32, order="canonical")
    assert np.array_equal(ids, z["ids"]) and np.array_equal(d.view(np.uint32), z["dist_bits"])
    assert np.array_equal(cnt, z["counts"]) and np.array_equal(st[:, :4], z["stats"])
    dp = vo.DualPrecisionHnsw.from_graph(g, train_count=200)
    assert np.array_equal(dp.quantizer.min_vals, z["sq8_min"]) and np.array_equal(dp.quantizer.scales, z["sq8_scale"])
    assert int(dp.quantizer.codes().astype(np.uint64).sum()) == int(z["sq8_codes_crc"][0])
    sids, sd, _, sst = dp.search_int8_batch(z["queries"], 5, 32, 4, order="canonical")
    assert np.array_equal(sids, z["sq8_ids"]) and np.array_equal(sd.view(np.uint32), z["sq8_dist_bits"])
    assert np.array_equal(sst[:, :4], z["sq8_stats"])
    bi, bs = vo.bruteforce_batch(vo.COSINE, g.vectors(), z["queries"], 5)
    assert np.array_equal(bi, z["bf_ids"]) and np.array_equal(bs.view(np.uint32), z["bf_score_bits"])
