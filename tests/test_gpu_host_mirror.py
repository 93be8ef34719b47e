"""Runs the C++ host mirror's test program (velesdb_b200/csrc/host/host_mirror_test.cpp), which drives
`veles::host::HnswIndex` through the C ABI the way the reference's index_tests.rs drives HnswIndex."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "velesdb_b200", "csrc", "host", "host_mirror_test")


def test_host_mirror_builds():
    assert os.path.exists(EXE), "run __graft_entry__.build() first"


@pytest.mark.gpu
def test_host_mirror_cpp(tmp_path):
    r = subprocess.run([EXE, str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout


@pytest.mark.gpu
def test_vacuum_rebuilds_a_searchable_index():
    # index/hnsw/index/vacuum.rs:110-190 end to end: tombstones gone, ids kept, searches still find their vectors
    import numpy as np
    from tests.gpu_util import latent_data
    from velesdb_b200 import DistanceMetric, HnswIndex, SearchQuality

    x = latent_data(400, 32, seed=12)
    ix = HnswIndex(32, DistanceMetric.Cosine)
    for i in range(400):
        ix.insert(10_000 + i, x[i])
    assert ix.search(x[5], 1)[0][0] == 10_005
    for i in range(0, 400, 3):
        ix.remove(10_000 + i)
    assert all((h[0] - 10_000) % 3 != 0 for h in ix.search(x[6], 10))  # removed ids never come back
    assert all((h[0] - 10_000) % 3 != 0 for r in ix.search_batch_parallel(x[:20], 10, SearchQuality.Balanced) for h in r)
    kept = ix.vacuum()
    assert kept == ix.len() == 400 - len(range(0, 400, 3)) and ix.tombstone_count() == 0
    res = ix.search_batch_parallel(x[[1, 2, 4, 5]], 3, SearchQuality.Accurate)
    assert [r[0][0] for r in res] == [10_001, 10_002, 10_004, 10_005]
    assert all((h[0] - 10_000) % 3 != 0 for r in res for h in r)
