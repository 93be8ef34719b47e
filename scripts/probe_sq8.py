"""Device-timing probe of the SQ8 traversal on the bench workload (dev tool; also the ncu target)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_data
from velesdb_b200 import DeviceSnapshot, DistanceMetric
from velesdb_b200 import _native as nv

n = int(os.environ.get("N", 1_000_000))
nq = int(os.environ.get("NQ", 1024))
reps = int(os.environ.get("REPS", 20))
dim, k, ef, over = 768, 10, 64, 4
nv.init(0)
dev = torch.device("cuda", 0)
x = gen_data(torch, n, dim, 24, 7, dev).cpu().numpy()
q = gen_data(torch, nq, dim, 24, 1_000_003, dev).contiguous()
snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
snap.build_graph(32)
snap.attach_sq8(1000)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    snap.search_batch_sq8_device(q, k, ef, over, ids, dist, cnt, None, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    snap.search_batch_sq8_device(q, k, ef, over, ids, dist, cnt, None, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"sq8 search: {ms:.3f} ms/batch {nq / ms * 1e3:.0f} q/s  checksum {int(ids.sum())}")
