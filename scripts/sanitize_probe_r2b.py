"""Kernels of the last round-2 session under compute-sanitizer (memcheck / racecheck / synccheck): the hybrid call (two
streams + fusion), the tile kernel's dynamic chunk-major items, and the relaxed brute force's bound / finish kernels
(CTA-wide selection, register-list overflow path, re-rank-everything path).  Small sizes; correctness is tests/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as vo  # builds the small graph (checker-side tool, not measured)
from velesdb_b200 import Bm25Snapshot, DeviceSnapshot, DistanceMetric, hybrid_search_batch

rng = np.random.default_rng(0)

# ---- hybrid call
n, dim, vocab = 1500, 32, 120
x = rng.normal(size=(n, dim)).astype(np.float32)
g = vo.Hnsw(vo.COSINE, dim, M=8, ef_construction=40)
g.insert_many(x)
snap = DeviceSnapshot.from_arrays(x, DistanceMetric.Cosine, g.export_graph(), g.M, g.M0, g.entry_point, g.max_layer)
p = 1.0 / np.arange(1, vocab + 1) ** 1.07
p /= p.sum()
lens = rng.integers(8, 40, n)
post, total = {}, 0
for d in range(n):
    t, c = np.unique(rng.choice(vocab, size=int(lens[d]), p=p), return_counts=True)
    total += int(lens[d])
    for a, b in zip(t, c):
        post.setdefault(int(a), []).append((d, int(b)))
term_ptr = np.zeros(vocab + 1, np.uint64)
pd, pt, df = [], [], np.zeros(vocab, np.uint32)
for t in range(vocab):
    l = post.get(t, [])
    df[t] = len(l)
    pd += [a for a, _ in l]
    pt += [b for _, b in l]
    term_ptr[t + 1] = len(pd)
bm = Bm25Snapshot(term_ptr, np.array(pd, np.uint32), np.array(pt, np.uint32), df, lens.astype(np.uint32), n, total)
q_ptr, q_terms = [0], []
nq = 9
for i in range(nq):
    q_terms += rng.choice(vocab, size=int(rng.integers(1, 5)), p=p).tolist() if i % 4 else [0xFFFFFFFF]
    q_ptr.append(len(q_terms))
for k in (3, 10):
    hybrid_search_batch(snap, bm, x[:nq] + 0.01, np.array(q_ptr, np.uint32), np.array(q_terms, np.uint32), k, 32, 0.5)

# ---- tile kernel, dynamic items (VELES_BF_DYNAMIC_MIN_BYTES lowers the 64 MB switch for this probe)
os.environ["VELES_BF_DYNAMIC_MIN_BYTES"] = "1"
x2 = rng.normal(size=(20_000, 128)).astype(np.float32)
for metric in (DistanceMetric.Cosine, DistanceMetric.Euclidean):
    s2 = DeviceSnapshot.from_vectors(x2, metric)
    s2.bruteforce_batch(x2[:40] + 0.01, 5)
os.environ.pop("VELES_BF_DYNAMIC_MIN_BYTES", None)

# ---- relaxed brute force: selection primitive, register-list path, re-rank-everything path, large kp (one-warp bound kernel)
x3 = rng.normal(size=(6000, 64)).astype(np.float32)
x3 /= np.linalg.norm(x3, axis=1, keepdims=True)
s3 = DeviceSnapshot.from_vectors(x3, DistanceMetric.Cosine)
for env in ({}, {"VELES_TC_REGTOPK": "1"}, {"VELES_TC_RERANK_ALL": "1"}, {"VELES_TC_OLD_TAIL": "1"}):
    for key in ("VELES_TC_REGTOPK", "VELES_TC_RERANK_ALL", "VELES_TC_OLD_TAIL"):
        os.environ.pop(key, None)
    os.environ.update(env)
    for k, over in ((10, 4), (25, 4), (40, 4)):
        s3.bruteforce_batch_relaxed(x3[:33] + 0.01, k, oversample=over)
print("probe done")
