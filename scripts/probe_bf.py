"""Device-timing probe of the brute-force scan (distance kernel + top-k) at several N and batch sizes (dev tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_data
from velesdb_b200 import DeviceSnapshot, DistanceMetric
from velesdb_b200 import _native as nv

nv.init(0)
dev = torch.device("cuda", 0)
dim, k = 768, 10
s = torch.cuda.current_stream().cuda_stream
for n in [int(v) for v in os.environ.get("NS", "10000,1000000").split(",")]:
    x = gen_data(torch, n, dim, 24, 7, dev).cpu().numpy()
    snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
    for nq in [int(v) for v in os.environ.get("NQS", "1,8,64,1024").split(",")]:
        q = gen_data(torch, nq, dim, 24, 1_000_003, dev).contiguous()
        ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
        sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
        for _ in range(3):
            snap.bruteforce_batch_device(q, k, ids, sc, s)
        torch.cuda.synchronize()
        reps = 20 if n * nq < 5e8 else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            snap.bruteforce_batch_device(q, k, ids, sc, s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        alg = n * dim * 4 + nq * dim * 4 + nq * k * 8
        print(f"n={n} nq={nq}: {ms:.4f} ms  {alg / ms / 1e6:.0f} GB/s alg  {n * nq * dim * 2 / ms / 1e9:.2f} TFLOP/s", flush=True)
    del snap
