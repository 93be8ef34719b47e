"""Bulk builder probe: build time, recall@k and evaluations per query of the block-insertion builder, optionally next
to an oracle-built (reference sequential insert) graph over the same vectors.

  python scripts/probe_build.py --n 100000 --dim 128 --M 16 --efc 200 [--oracle] [--dtype f32|f16|bin1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--ef", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--latent", type=int, default=24)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--oracle", action="store_true", help="also build with the oracle's sequential insert (slow)")
    ap.add_argument("--oracle-efc", type=int, default=0)
    ap.add_argument("--parity", type=int, default=0, help="compare this many queries with the oracle on the exported graph")
    a = ap.parse_args()
    import torch

    from bench import gen_data
    from velesdb_b200 import DeviceSnapshot, DistanceMetric

    dev = torch.device("cuda", 0)
    out = {"n": a.n, "dim": a.dim, "M": a.M, "efc": a.efc, "ef": a.ef, "k": a.k, "dtype": a.dtype}
    if a.dtype == "bin1":
        g = torch.Generator(device=dev)
        g.manual_seed(7)
        # clustered bits: centres + 12% flips, so that Hamming neighbours exist
        nc = max(16, a.n // 2000)
        cen = torch.randint(0, 2, (nc, a.dim), generator=g, device=dev, dtype=torch.int8)

        def bits(m, seed):
            g.manual_seed(seed)
            c = torch.randint(0, nc, (m,), generator=g, device=dev)
            flip = (torch.rand(m, a.dim, generator=g, device=dev) < 0.12).to(torch.int8)
            return (cen[c] ^ flip).to(torch.float32)

        x = bits(a.n, 11).cpu().numpy()
        q = bits(a.nq, 99).cpu().numpy()
        metric = DistanceMetric.Hamming
    else:
        x = gen_data(torch, a.n, a.dim, a.latent, 7, dev).cpu().numpy()
        q = gen_data(torch, a.nq, a.dim, a.latent, 1_000_003, dev).cpu().numpy()
        metric = DistanceMetric.Cosine
    snap = DeviceSnapshot.from_vectors(x, metric, store_dtype=a.dtype)
    torch.cuda.synchronize()
    t = time.time()
    snap.build_graph(a.M, a.efc)
    torch.cuda.synchronize()
    out["build_s"] = round(time.time() - t, 3)
    out["inserts_per_s"] = round(a.n / out["build_s"])

    def quality(s, tag):
        ids, dist, cnt, st = s.search_batch(q, a.k, a.ef, with_stats=True)
        bi, _ = s.bruteforce_batch(q, a.k)
        rec = float(np.mean([len(set(ids[i].tolist()) & set(bi[i].tolist())) / a.k for i in range(len(q))]))
        out[tag] = {"recall": round(rec, 4), "ndc": float((st[:, 0] + st[:, 2]).mean()), "hops": float(st[:, 1].mean())}
        t0 = time.time()
        for _ in range(5):
            s.search_batch(q, a.k, a.ef)
        out[tag]["qps_host_api"] = round(5 * len(q) / (time.time() - t0))
        return ids, dist, st

    ids, dist, st = quality(snap, "bulk")
    layers = snap.export_graph()
    deg0 = np.diff(layers[0][0].astype(np.int64))
    out["bulk"]["deg0_mean"] = float(deg0.mean())
    out["bulk"]["deg0_min"] = int(deg0.min())
    if a.parity or a.oracle:
        from oracle import oracle as vo
    if a.parity and a.dtype == "f32":
        g = vo.Hnsw.from_arrays(int(metric), x, layers, a.M, 2 * a.M, snap.entry_point, snap.max_layer)
        m = a.parity
        oi, od, oc, ost = g.search_batch(q[:m], a.k, a.ef, order="canonical", threads=os.cpu_count())
        keep = ost[:, 4] == 0
        out["parity"] = {"queries": m, "ids_equal": bool(np.array_equal(ids[:m][keep], oi[keep].astype(np.uint32))),
                         "dist_bits_equal": bool(np.array_equal(dist[:m].view(np.uint32), od.view(np.uint32))),
                         "ndc_equal": bool(np.array_equal(st[:m, 0], ost[:, 0].astype(np.uint32))),
                         "tie_at_k": int((~keep).sum())}
    if a.oracle:
        efc = a.oracle_efc or a.efc
        t = time.time()
        g = vo.Hnsw(int(metric), a.dim, M=a.M, ef_construction=efc)
        g.insert_many(x)
        out["oracle_build_s"] = round(time.time() - t, 2)
        s2 = DeviceSnapshot.from_arrays(x, metric, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
        quality(s2, "oracle_built")
        out["oracle_built"]["efc"] = efc
    print(json.dumps(out))


if __name__ == "__main__":
    main()
