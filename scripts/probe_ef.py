"""Recall / throughput of one config's index over a sweep of ef_search: prints one JSON line per ef and, last, the
smallest ef whose recall reaches --target.   python scripts/probe_ef.py --config c4 --n 4000000 --efs 64,128,256,512"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--efs", default="64,128,256,512")
    ap.add_argument("--latent", type=int, default=None)
    ap.add_argument("--target", type=float, default=0.95)
    a = ap.parse_args()
    import torch

    from bench import CONFIGS, build_snapshot, make_queries
    from velesdb_b200 import _native as nv

    nv.init(0)
    dev = torch.device("cuda", 0)
    cfg = dict(CONFIGS[a.config], name=a.config)
    if a.n:
        cfg["n"] = a.n
    if a.latent:
        cfg["latent"] = a.latent
    k = cfg["k"]
    snap, _, t_gen, t_build = build_snapshot(torch, cfg, dev)
    q_d = make_queries(torch, cfg, a.nq, 1_000_003, dev)
    stream = torch.cuda.current_stream().cuda_stream
    gi = torch.empty((a.nq, k), dtype=torch.int32, device=dev)
    gs = torch.empty((a.nq, k), dtype=torch.float32, device=dev)
    snap.bruteforce_batch_device(q_d, k, gi, gs, stream)
    torch.cuda.synchronize()
    gin, gsn = gi.cpu().numpy(), gs.cpu().numpy()
    best = None
    for ef in [int(x) for x in a.efs.split(",")]:
        ids = torch.empty((a.nq, k), dtype=torch.int32, device=dev)
        dist = torch.empty((a.nq, k), dtype=torch.float32, device=dev)
        cnt = torch.empty(a.nq, dtype=torch.int32, device=dev)
        st = torch.zeros((a.nq, 4), dtype=torch.int32, device=dev)
        snap.search_batch_device(q_d, k, max(ef, k), ids, dist, cnt, st, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            snap.search_batch_device(q_d, k, max(ef, k), ids, dist, cnt, None, stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        if cfg["store"] == "bin1":
            rec = float(np.mean(dist.cpu().numpy() <= gsn[:, k - 1:k]))
        else:
            got = ids.cpu().numpy()
            rec = float(np.mean([len(set(got[i].tolist()) & set(gin[i].tolist())) / k for i in range(a.nq)]))
        s = st.cpu().numpy()
        print(json.dumps({"config": a.config, "n": cfg["n"], "latent": cfg["latent"], "ef": ef, "recall": round(rec, 4),
                          "ndc": float((s[:, 0] + s[:, 2]).mean()), "ms_per_batch": ms, "qps": a.nq / ms * 1e3,
                          "build_s": round(t_build, 2)}), flush=True)
        if best is None and rec >= a.target:
            best = ef
    print(json.dumps({"chosen_ef": best}))


if __name__ == "__main__":
    main()
