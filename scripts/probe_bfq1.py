"""Single-query exact brute force over a small collection (C1: 10K x 768): device-timed, cold L2; with
VELES_BF_DEBUG_TIMING=1 the library prints per-phase globaltimer stamps of the fused scan kernel to stderr."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import CONFIGS, build_snapshot, make_queries
from velesdb_b200 import _native as nv

nv.init(0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cfg = dict(CONFIGS["c1"], n=10_000, name="probe")
snap, _, _, _ = build_snapshot(torch, cfg, dev)
k = 10
stream = torch.cuda.current_stream().cuda_stream
for nq in (1, 2):
    q_d = make_queries(torch, cfg, nq, 99, dev)
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(3):
        snap.bruteforce_batch_device(q_d, k, ids, sc, stream)
    torch.cuda.synchronize()
    for cold in (True, False):
        ts = []
        for _ in range(20):
            if cold:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            snap.bruteforce_batch_device(q_d, k, ids, sc, stream)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(json.dumps({"n": 10_000, "nq": nq, "cold_l2": cold, "us_median": float(np.median(ts)) * 1e3, "us_min": float(np.min(ts)) * 1e3}), flush=True)
