"""ncu target: the cooperative (multi-warp-per-query) search at a small batch (dev tool)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_data
from velesdb_b200 import DeviceSnapshot, DistanceMetric
from velesdb_b200 import _native as nv

nq = int(os.environ.get("NQ", 64))
nv.init(0)
dev = torch.device("cuda", 0)
x = gen_data(torch, 1_000_000, 768, 24, 7, dev).cpu().numpy()
q = gen_data(torch, nq, 768, 24, 1_000_003, dev).contiguous()
snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
snap.build_graph(32)
ids = torch.empty((nq, 10), dtype=torch.int32, device=dev)
dist = torch.empty((nq, 10), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    snap.search_batch_device(q, 10, 64, ids, dist, cnt, None, s)
torch.cuda.synchronize()
print("done")
