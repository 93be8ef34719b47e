"""Exact brute force after the top-k fusion: C1 at Q = 1 and Q = 1024, and 1M x 768 at Q = 1 / 1024, device-timed."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import CONFIGS, build_snapshot, make_queries, measured_peak
from velesdb_b200 import _native as nv

nv.init(0)
dev = torch.device("cuda", 0)
peak, _ = measured_peak()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in (10_000, 1_000_000):
    cfg = dict(CONFIGS["c1"], n=n, name="probe")
    snap, _, _, _ = build_snapshot(torch, cfg, dev)
    for nq in (1, 8, 1024):
        q_d = make_queries(torch, cfg, nq, 99, dev)
        k = 10
        ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            snap.bruteforce_batch_device(q_d, k, ids, sc, stream)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            snap.bruteforce_batch_device(q_d, k, ids, sc, stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        alg = n * 768 * 4 + nq * 768 * 4 + nq * k * 8
        print(json.dumps({"n": n, "nq": nq, "ms": ms, "us": ms * 1e3, "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / peak,
                          "pair_evals_per_s": n * nq / ms * 1e3}), flush=True)
