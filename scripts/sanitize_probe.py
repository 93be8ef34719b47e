"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck); dev tool.
No oracle here: it only has to touch the code paths; correctness is what tests/ checks."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from velesdb_b200 import Bm25Index, DeviceSnapshot, DistanceMetric, FusionStrategy

rng = np.random.default_rng(0)
for dim, n in ((96, 1500), (100, 700)):
    z = rng.normal(size=(n, 8)).astype(np.float32) @ rng.normal(size=(8, dim)).astype(np.float32)
    x = (z + 0.2 * rng.normal(size=(n, dim))).astype(np.float32)
    q = (x[:40] + 0.1).astype(np.float32)
    for metric in (DistanceMetric.Cosine, DistanceMetric.Euclidean):
        snap = DeviceSnapshot.from_vectors(x, metric)
        snap.build_graph(16)
        for w in ("1", "2", "4", "8"):
            os.environ["VELES_SEARCH_WARPS"] = w
            snap.search_batch(q, 10, 64)
            snap.search_batch(q, 50, 200)
        os.environ.pop("VELES_SEARCH_WARPS")
        snap.search_batch(q, 10, 600)
        snap.attach_sq8(min(1000, n))
        for w in ("1", "4"):
            os.environ["VELES_SEARCH_WARPS"] = w
            snap.search_batch_sq8(q, 10, 64, 4)
        os.environ.pop("VELES_SEARCH_WARPS")
        for nq in (1, 3, 8, 20):
            snap.bruteforce_batch(q[:nq], 10)
    h = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine, store_dtype="f16")
    h.build_graph(16)
    h.search_batch(q, 10, 64)
    h.bruteforce_batch(q[:2], 5)
xs = rng.normal(size=(900, 128)).astype(np.float32)   # dim % 128 == 0: shared-memory staged brute-force tile
for store in ("f32", "f16"):
    t = DeviceSnapshot.from_vectors(xs, DistanceMetric.Cosine, store_dtype=store)
    t.bruteforce_batch(xs[:37], 10)
t = DeviceSnapshot.from_vectors(xs, DistanceMetric.Euclidean)
t.build_graph(16)
t.set_id_map(np.arange(900, dtype=np.uint64) + 7, np.full(29, 0xAAAAAAAA, np.uint32))
t.search_batch_mapped(xs[:20], 5, 64, k_fetch=30, allow_bits=np.full(29, 0xF0F0F0F0, np.uint32))
t.search_batch_multi_entry(xs[:20], 5, 32, np.tile(np.array([[3, 3, 800]], np.uint32), (20, 1)))
small = DeviceSnapshot.from_vectors(x[:300], DistanceMetric.Euclidean)
small.build_graph_exact(8, 40)
small.search_batch(q, 5, 32)
bits = (x[:, :64] > 0).astype(np.float32)
xb = np.ascontiguousarray(np.tile(bits, (1, 2)))
fb = DeviceSnapshot.from_vectors(xb, DistanceMetric.Hamming)
fb.build_graph(16)
b = DeviceSnapshot.from_arrays(xb, DistanceMetric.Hamming, fb.export_graph(), 16, 32, fb.entry_point, fb.max_layer, store_dtype="bin1")
b.search_batch(xb[:16], 10, 64)
b.bruteforce_batch(xb[:3], 5)
tx = Bm25Index()
words = ["alpha", "beta", "gamma", "delta", "epsilon", "zeta"]
for i in range(500):
    tx.add_document(i, " ".join(rng.choice(words, size=int(rng.integers(2, 9)))))
tx.search("alpha gamma gamma", 10)
FusionStrategy.RRF().fuse([[(1, 0.9), (2, 0.8)], [(2, 0.7), (3, 0.6)]])
print("sanitize probe done")
