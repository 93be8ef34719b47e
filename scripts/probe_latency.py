"""Small-batch latency of the HNSW search (f32 traversal and SQ8 dual precision) on the bench index (dev tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_data
from velesdb_b200 import DeviceSnapshot, DistanceMetric
from velesdb_b200 import _native as nv

n, dim, k, ef = int(os.environ.get("N", 1_000_000)), 768, 10, 64
nv.init(0)
dev = torch.device("cuda", 0)
x = gen_data(torch, n, dim, 24, 7, dev).cpu().numpy()
snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
snap.build_graph(32)
snap.attach_sq8(1000)
s = torch.cuda.current_stream().cuda_stream
for nq in (1, 8, 64, 256, 1024):
    q = gen_data(torch, nq, dim, 24, 1_000_003, dev).contiguous()
    ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    out = []
    for name, fn in (("f32", lambda: snap.search_batch_device(q, k, ef, ids, dist, cnt, None, s)),
                     ("sq8", lambda: snap.search_batch_sq8_device(q, k, ef, 4, ids, dist, cnt, None, s))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 20 * 1e3:.0f} us")
    print(f"nq={nq}: " + "  ".join(out), flush=True)
