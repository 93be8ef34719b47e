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


def test_committed_bench_lines_follow_the_contract():
    # the JSON line bench.py printed on the B200 (profiles/), checked for the keys the driver reads
    import glob
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = sorted(glob.glob(os.path.join(root, "profiles", "r1_bench_n*_v[3-9].json")))
    assert lines
    for path in lines:
        j = json.load(open(path))
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
            assert key in j, (path, key)
        assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None
        assert j["warmup"] >= 3 and j["gpu_launches"] == j["steps"] and "workload" in j["config"]
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(j["e2e"])
        assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["value"] <= j["value"] * 1.02
        r = j["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
        assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"] * 0.98
        c = j["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and c["ids_match_gpu"] is True
        assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_round2_bench_lines_follow_the_contract():
    # every config's line printed on the B200 in round 2 (profiles/r2*_bench_*.json)
    import glob
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lines = sorted(glob.glob(os.path.join(root, "profiles", "r2*_bench_*.json")))
    assert len(lines) >= 10
    seen_ref = False
    for path in lines:
        j = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"):
            assert key in j, (path, key)
        assert j["higher_is_better"] is True and j["vs_baseline"] is None and "workload" in j["config"]
        assert "build_s" not in j["config"] and "datagen_s" not in j["config"]   # keep timings out of `config`
        assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(j["e2e"])
        if j.get("impl") == "reference":
            seen_ref = True
            assert j["gpu_launches"] == 0 and j["cpu_baseline"]["kind"] == "port" and j["steps"] >= 20
            assert all(len(v) == 64 for v in j["frozen_files"]["sha256"].values())
            continue
        assert j["warmup"] >= 3 and j["gpu_launches"] >= j["steps"]
        r = j["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] <= 1.0
        assert r["kernel_ms"] <= j["ms_per_step"] * 1.001
        assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes_per_launch"] * 0.98
        assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if "cpu_baseline" in j:
            assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["ids_match_gpu"] is True
        if "parity" in j:
            assert j["parity"]["ids_and_distance_bits_equal_oracle"] is True and j["parity"]["queries"] >= 256
    assert seen_ref


def test_integration_md_ffi_block_matches_the_header():
    # every `pub fn veles_*` INTEGRATION.md shows a maintainer exists in include/veles_b200.h with the same arity
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(root, "include", "veles_b200.h")).read(), flags=re.S)
    rust = re.findall(r"pub fn (veles_\w+)\s*\((.*?)\)\s*->", md, flags=re.S)
    assert len(rust) >= 20
    for name, args in rust:
        m = re.search(r"\b%s\s*\((.*?)\)\s*;" % name, hdr, flags=re.S)
        assert m, f"{name} is in INTEGRATION.md but not in the header"
        c_args = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        r_args = [a for a in args.split(",") if a.strip()]
        assert len(c_args) == len(r_args), (name, len(c_args), len(r_args))
