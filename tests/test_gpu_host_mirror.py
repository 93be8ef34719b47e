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
