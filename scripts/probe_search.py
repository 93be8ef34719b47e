"""Quick device-timing probe of the HNSW search kernel on an oracle-built graph (dev tool)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from oracle import oracle as vo
from tests.gpu_util import latent_data, queries_near
from velesdb_b200 import DeviceSnapshot

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
dim, nq, k, ef = 768, 1024, 10, 64
x = latent_data(n, dim, latent=16, noise=0.3, seed=0, normalize=True)
t = time.time()
g = vo.Hnsw(vo.COSINE, dim, M=16, ef_construction=48)
g.insert_many(x)
print(f"oracle build n={n}: {time.time()-t:.1f}s", flush=True)
q = queries_near(x, nq, jitter=0.02, seed=3)
snap = DeviceSnapshot.from_arrays(x, vo.COSINE, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
qt = torch.from_numpy(q).cuda()
ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
dist = torch.empty((nq, k), dtype=torch.float32, device="cuda")
cnt = torch.empty(nq, dtype=torch.int32, device="cuda")
st = torch.empty((nq, 4), dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    snap.search_batch_device(qt, k, ef, ids, dist, cnt, st, s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    snap.search_batch_device(qt, k, ef, ids, dist, cnt, st, s)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
stc = st.cpu().numpy().astype(np.int64)
ndc = stc[:, 0] + stc[:, 2]
bytes_q = ndc * (dim * 4 + 16) + stc[:, 1] * 32 * 4 * 2 + stc[:, 3] * 32 * 4
print(f"search: {ms:.3f} ms/batch  {nq/ms*1e3:.0f} QPS  ndc mean {ndc.mean():.0f}  hops0 {stc[:,1].mean():.1f} "
      f"alg GB/s {bytes_q.sum()/ms/1e6:.1f}")
gi, _ = snap.bruteforce_batch(q, k)
rec = np.mean([len(set(ids[i].cpu().numpy().tolist()) & set(gi[i].astype(np.int32).tolist())) / k for i in range(nq)])
print("recall@10 vs GPU brute force:", rec)
t = time.time()
oi, od, oc, ost = g.search_batch(q[:256], k, ef, threads=8)
dt = time.time() - t
print(f"oracle 8 threads: {256/dt:.0f} QPS")
assert np.array_equal(oi.astype(np.int32), ids[:256].cpu().numpy()), "parity"
print("parity ok")
