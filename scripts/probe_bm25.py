"""BM25 batch search through the host API: ms per 1024-query batch for each kernel variant (env switch), 1M docs."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, bm25_corpus, measured_peak
from velesdb_b200 import Bm25Snapshot
from velesdb_b200 import _native as nv

nv.init(0)
cfg = dict(CONFIGS["c5"])
nq, k = 1024, 20
corp = bm25_corpus(cfg, nq, 11)
bm = Bm25Snapshot(corp["term_ptr"], corp["post_doc"], corp["tf"], corp["df"], corp["lens"], corp["n_docs"], corp["total"])
alg = int(corp["df"][corp["q_terms"]].astype(np.int64).sum() * 12 + nq * k * 8)
peak, _ = measured_peak()
ref = None
KEYS = ("VELES_BM25_WALK", "VELES_BM25_SLICE", "VELES_BM25_PREFETCH", "VELES_BM25_HASH", "VELES_BM25_PARTS", "VELES_BM25_FLAT_OCC",
        "VELES_BM25_FLAT", "VELES_BM25_SUB_OCC")
VARIANTS = (("sub (default)", {}), ("sub, 8 CTAs per SM", {"VELES_BM25_SUB_OCC": "8"}), ("sub, 12 CTAs per SM", {"VELES_BM25_SUB_OCC": "12"}),
            ("sub, 12 parts", {"VELES_BM25_PARTS": "12"}), ("sub, 40 parts", {"VELES_BM25_PARTS": "40"}),
            ("flat", {"VELES_BM25_FLAT": "1"}), ("walk (round 1, precomputed postings)", {"VELES_BM25_WALK": "1"}))
if len(sys.argv) > 1 and sys.argv[1] == "all":
    VARIANTS += (("slice", {"VELES_BM25_SLICE": "1"}), ("prefetch", {"VELES_BM25_PREFETCH": "1"}), ("hash", {"VELES_BM25_HASH": "1"}))
if len(sys.argv) > 1 and sys.argv[1] == "default":
    VARIANTS = VARIANTS[:1]
for name, env in VARIANTS:
    for key in KEYS:
        os.environ.pop(key, None)
    os.environ.update(env)
    for _ in range(3):
        out = bm.search_batch(corp["q_ptr"], corp["q_terms"], k)
    ts = []
    for _ in range(10):
        t = time.perf_counter()
        out = bm.search_batch(corp["q_ptr"], corp["q_terms"], k)
        ts.append(time.perf_counter() - t)
    ms = float(np.median(ts)) * 1e3
    same = True if ref is None else bool(np.array_equal(ref[0], out[0]) and np.array_equal(ref[1].view(np.uint32), out[1].view(np.uint32)))
    ref = ref or out
    print(json.dumps({"kernel": name, "ms_per_batch_host_api": ms, "queries_per_s": nq / ms * 1e3, "alg_bytes": alg,
                      "alg_GBps_incl_copies": alg / ms / 1e6, "frac_hbm_incl_copies": alg / ms / 1e6 / peak,
                      "same_as_first": same}), flush=True)
