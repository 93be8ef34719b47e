"""Tensor-core relaxed brute force vs the exact kernels: time, TFLOP/s of the GEMM passes, recall.
  python scripts/probe_gemm.py --n 1000000 --dim 768 --nq 1024 --k 10"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--over", type=int, default=4)
    ap.add_argument("--store", default="f32")
    ap.add_argument("--skip-exact", action="store_true")
    a = ap.parse_args()
    import torch

    from bench import CONFIGS, build_snapshot, make_queries
    from velesdb_b200 import _native as nv

    nv.init(0)
    dev = torch.device("cuda", 0)
    cfg = dict(CONFIGS["c1"], n=a.n, dim=a.dim, nq=a.nq, k=a.k, store=a.store, name="probe")
    snap, _, t_gen, _ = build_snapshot(torch, cfg, dev)
    q_d = make_queries(torch, cfg, a.nq, 99, dev)
    stream = torch.cuda.current_stream().cuda_stream
    ids = torch.empty((a.nq, a.k), dtype=torch.int32, device=dev)
    sc = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
    out = {"n": a.n, "dim": a.dim, "nq": a.nq, "k": a.k, "oversample": a.over, "store": a.store}
    snap.bruteforce_batch_relaxed_device(q_d, a.k, a.over, ids, sc, stream)   # builds the fp16 copy
    torch.cuda.synchronize()
    reps = 5
    t = time.time()
    gemm = 0.0
    for _ in range(reps):
        gemm += snap.bruteforce_batch_relaxed_device(q_d, a.k, a.over, ids, sc, stream, want_gemm_ms=True)
    torch.cuda.synchronize()
    out["relaxed_ms"] = (time.time() - t) / reps * 1e3
    out["gemm_ms"] = gemm / reps
    dpad = (a.dim + 63) // 64 * 64
    flops = 2.0 * a.n * dpad * a.nq
    out["gemm_tflops"] = flops / (out["gemm_ms"] / 1e3) / 1e12
    out["queries_per_s"] = a.nq / (out["relaxed_ms"] / 1e3)
    if not a.skip_exact:
        ei = torch.empty((a.nq, a.k), dtype=torch.int32, device=dev)
        es = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
        snap.bruteforce_batch_device(q_d, a.k, ei, es, stream)
        torch.cuda.synchronize()
        t = time.time()
        snap.bruteforce_batch_device(q_d, a.k, ei, es, stream)
        torch.cuda.synchronize()
        out["exact_ms"] = (time.time() - t) * 1e3
        a_, b_ = ids.cpu().numpy(), ei.cpu().numpy()
        out["recall_vs_exact"] = float(np.mean([len(set(a_[i].tolist()) & set(b_[i].tolist())) / a.k for i in range(a.nq)]))
        out["ids_identical_fraction"] = float((a_ == b_).all(axis=1).mean())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
