"""CPU-side checks of the C ABI library: it loads, exports every symbol include/veles_b200.h
declares, its pure host helpers agree with the oracle, and it fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest

from oracle import oracle as vo
from velesdb_b200 import _native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "veles_b200.h")).read()
    declared = set(re.findall(r"\b(veles_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    lib = nv.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in veles_b200.h but not exported"
    assert declared == {s[0] for s in nv.SYMBOLS}, "ctypes table and header disagree"


def test_ef_search_and_transform_score_match_oracle():
    lib = nv.lib()
    for q in range(5):
        for k in (1, 10, 40, 100, 500):
            assert lib.veles_ef_search(q, k, 64) == vo.ef_search(q, k, 64)
    for m in range(5):
        for raw in (-0.5, 0.0, 0.3, 1.0, 1.5, 7.25):
            assert lib.veles_transform_score(m, raw) == vo.transform_score(m, raw)


def test_no_cuda_device_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = nv.lib()
    assert lib.veles_init(0) == nv.ERR_CUDA
    assert b"no CPU path" in lib.veles_last_error()
    with pytest.raises(nv.VelesError):
        nv.init()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "velesdb_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "veles_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f
