"""Round-2 kernels under compute-sanitizer (memcheck / racecheck / synccheck): BM25 query kernels (sub, flat, walk), the
fused brute-force scan with both tails, the block-insertion builder and append.  Small sizes; correctness is tests/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from velesdb_b200 import Bm25Snapshot, DeviceSnapshot, DistanceMetric

rng = np.random.default_rng(0)

# ---- BM25: 9000 docs = 9 sub-ranges = 2 ranges; frequent terms span several steps / rounds
n_docs, vocab = 9000, 300
p = 1.0 / np.arange(1, vocab + 1) ** 1.07
p /= p.sum()
lens = rng.integers(8, 60, n_docs)
post = {}
total = 0
for d in range(n_docs):
    t, c = np.unique(rng.choice(vocab, size=int(lens[d]), p=p), return_counts=True)
    total += int(lens[d])
    for a, b in zip(t, c):
        post.setdefault(int(a), []).append((d, int(b)))
term_ptr = np.zeros(vocab + 1, np.uint64)
pd, pt, df = [], [], np.zeros(vocab, np.uint32)
for t in range(vocab):
    l = post.get(t, [])
    df[t] = len(l)
    pd += [x[0] for x in l]
    pt += [x[1] for x in l]
    term_ptr[t + 1] = len(pd)


def make(no_fine):
    if no_fine:
        os.environ["VELES_BM25_NO_FINE_TABLE"] = "1"
    s = Bm25Snapshot(term_ptr, np.array(pd, np.uint32), np.array(pt, np.uint32), df, lens.astype(np.uint32), n_docs, total)
    os.environ.pop("VELES_BM25_NO_FINE_TABLE", None)
    return s


q_ptr, q_terms = [0], []
for i in range(12):
    t = rng.choice(vocab, size=int(rng.integers(1, 7)), p=p).tolist()
    if i % 4 == 0:
        t += t[:1] + [0xFFFFFFFF]
    q_terms += t
    q_ptr.append(len(q_terms))
KEYS = ("VELES_BM25_PARTS", "VELES_BM25_FLAT", "VELES_BM25_WALK")
for snap in (make(False), make(True)):
    for env in ({}, {"VELES_BM25_PARTS": "1"}, {"VELES_BM25_PARTS": "3"}, {"VELES_BM25_FLAT": "1"}, {"VELES_BM25_FLAT": "1", "VELES_BM25_PARTS": "2"},
                {"VELES_BM25_WALK": "1"}):
        for key in KEYS:
            os.environ.pop(key, None)
        os.environ.update(env)
        for k in (5, 40, 100):
            snap.search_batch(q_ptr, q_terms, k)
        snap.search_batch(q_ptr[:2], q_terms[:q_ptr[1]], 10)
for key in KEYS:
    os.environ.pop(key, None)

# ---- fused brute-force scan: 1, 2 and 5 queries, both tails, small and tiny collections
for n, dim in ((3000, 96), (7, 64)):
    x = rng.normal(size=(n, dim)).astype(np.float32)
    for metric in (DistanceMetric.Cosine, DistanceMetric.Euclidean, DistanceMetric.DotProduct):
        for store in ("f32", "f16"):
            s = DeviceSnapshot.from_vectors(x, metric, store_dtype=store)
            for env in ({}, {"VELES_BF_NO_STAGED_TAIL": "1"}):
                os.environ.pop("VELES_BF_NO_STAGED_TAIL", None)
                os.environ.update(env)
                for nq in (1, 2, 5):
                    for k in (3, 10, 40):
                        s.bruteforce_batch(x[:nq] + 0.01, k)
os.environ.pop("VELES_BF_NO_STAGED_TAIL", None)

# ---- block-insertion builder and append
x = rng.normal(size=(2500, 64)).astype(np.float32)
b = DeviceSnapshot.from_vectors(x[:2000], DistanceMetric.Cosine)
b.build_graph(16, 100)
b.search_batch(x[:16], 10, 64)
b.append(x[2000:], 100)
b.search_batch(x[2000:2016], 10, 64)
print("sanitize probe (round 2) done")
