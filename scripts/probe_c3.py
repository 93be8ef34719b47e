"""Device-timing probe of the f16 / k=100 / ef=256 / batch 8192 search (config C3 shape) on 1M rows (dev tool, ncu target)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_data
from velesdb_b200 import DeviceSnapshot, DistanceMetric
from velesdb_b200 import _native as nv

n = int(os.environ.get("N", 1_000_000))
nq = int(os.environ.get("NQ", 8192))
reps = int(os.environ.get("REPS", 5))
k, ef = int(os.environ.get("K", 100)), int(os.environ.get("EF", 256))
store = os.environ.get("STORE", "f16")
dim = 768
nv.init(0)
dev = torch.device("cuda", 0)
x = gen_data(torch, n, dim, 24, 7, dev).cpu().numpy()
q = gen_data(torch, nq, dim, 24, 1_000_003, dev).contiguous()
b = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
b.build_graph(32)
snap = b if store == "f32" else DeviceSnapshot.from_arrays(x, DistanceMetric.Cosine, b.export_graph(), 32, 64, b.entry_point, b.max_layer, store_dtype=store)
ids = torch.empty((nq, k), dtype=torch.int32, device=dev)
dist = torch.empty((nq, k), dtype=torch.float32, device=dev)
cnt = torch.empty(nq, dtype=torch.int32, device=dev)
st = torch.empty((nq, 4), dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    snap.search_batch_device(q, k, ef, ids, dist, cnt, st, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    snap.search_batch_device(q, k, ef, ids, dist, cnt, None, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
sc = st.cpu().numpy().astype(np.int64)
elt = 2 if store == "f16" else 4
alg = int(((sc[:, 0] + sc[:, 2]) * dim * elt + sc[:, 1] * 256 + sc[:, 3] * 128 + dim * 4 + k * 8).sum())
print(f"{store} k={k} ef={ef} nq={nq}: {ms:.3f} ms/batch {nq / ms * 1e3:.0f} q/s  ndc {float((sc[:,0]+sc[:,2]).mean()):.0f} hops {float(sc[:,1].mean()):.0f}  {alg / ms / 1e6:.0f} GB/s alg")
