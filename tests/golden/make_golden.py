#!/usr/bin/env python
"""Generates the fixtures in this directory.  Run from the repo root: python tests/golden/make_golden.py

The reference (Rust) cannot be built or run in this image, so none of these files are outputs of VelesDB
itself.  What they pin instead:

spec_v1.{vectors,graph} + spec_v1_expected.json
    A 6-node, 2-layer index written byte by byte from the reference's on-disk format v1
    (velesdb-core/src/index/hnsw/native/backend_adapter.rs:184-261), with struct.pack only -- no code of this
    repo touches it.  Layer 0 is fully connected, so a search with ef >= 6 must return the exact nearest
    neighbours; the expected ids come from a numpy brute force.  Pins the loaders (oracle and GPU) against the
    format, and the search against an independent answer.

regress_cos24.{vectors,graph} + regress_cos24.npz
    Regression fixtures written by the CPU oracle (oracle/veles_oracle.cpp) at the time of generation: a seeded
    300 x 24 cosine index built with the reference's sequential insert, 16 queries, and the oracle's answers
    (ids, distance bits, evaluation counts) for the f32 traversal, the SQ8 dual-precision path and brute force.
    They guard the oracle and the CUDA path against drifting apart from what round 1 validated.
"""
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def write_v1(base, vectors, layers, M, M0, ef_c, entry, max_layer):
    n, dim = vectors.shape
    with open(base + ".vectors", "wb") as f:  # u32 version, u64 count, u32 dim, count*dim f32 (LE)
        f.write(struct.pack("<IQI", 1, n, dim))
        f.write(vectors.astype("<f4").tobytes())
    with open(base + ".graph", "wb") as f:
        # u32 version, u32 num_layers, u32 M, u32 M0, u32 ef_construction, u64 entry_point, u32 max_layer, u64 count
        f.write(struct.pack("<IIIIIQIQ", 1, len(layers), M, M0, ef_c, entry, max_layer, n))
        for layer in layers:  # u64 num_nodes, then per node: u32 degree, degree * u32 neighbour
            f.write(struct.pack("<Q", len(layer)))
            for nb in layer:
                f.write(struct.pack("<I", len(nb)))
                f.write(struct.pack("<%dI" % len(nb), *nb))


def spec_fixture():
    x = np.array([[0, 0, 0, 1], [1, 0, 0, 0], [0.9, 0.1, 0, 0], [0, 1, 0, 0], [0, 0.8, 0.6, 0], [0.5, 0.5, 0.5, 0.5]], np.float32)
    n = len(x)
    layer0 = [[j for j in range(n) if j != i] for i in range(n)]
    layer1 = [[], [4], [], [], [1], []]          # nodes 1 and 4 live on layer 1; entry point 4
    write_v1(os.path.join(HERE, "spec_v1"), x, [layer0, layer1], 4, 8, 16, 4, 1)
    qs = np.array([[1, 0.05, 0, 0], [0, 0.9, 0.5, 0], [0.4, 0.4, 0.4, 0.6]], np.float32)
    exp = {}
    for name, metric in (("euclidean", 1), ("cosine", 0)):
        ids = []
        for q in qs:
            if metric == 1:
                d = np.sqrt(((x - q) ** 2).sum(1).astype(np.float64))
            else:
                d = 1.0 - (x @ q).astype(np.float64) / (np.linalg.norm(x, axis=1) * np.linalg.norm(q))
            ids.append(np.argsort(d, kind="stable")[:3].tolist())
        exp[name] = ids
    json.dump({"queries": qs.tolist(), "k": 3, "ef": 16, "expected_ids": exp, "n": n, "dim": 4, "entry_point": 4,
               "max_layer": 1}, open(os.path.join(HERE, "spec_v1_expected.json"), "w"), indent=1)


def regression_fixture():
    from oracle import oracle as vo
    rng = np.random.default_rng(20261017)
    n, dim, nq = 300, 24, 16
    z = rng.normal(size=(n, 6)).astype(np.float32) @ rng.normal(size=(6, dim)).astype(np.float32)
    x = (z + 0.2 * rng.normal(size=(n, dim))).astype(np.float32)
    q = (x[rng.integers(0, n, nq)] + 0.1 * rng.normal(size=(nq, dim))).astype(np.float32)
    g = vo.Hnsw(vo.COSINE, dim, M=8, ef_construction=40)
    g.insert_many(x)
    g.dump(HERE, "regress_cos24")
    ids, d, cnt, st = g.search_batch(q, 5, 32, order="canonical")
    dp = vo.DualPrecisionHnsw.from_graph(g, train_count=200)
    sids, sd, scnt, sst = dp.search_int8_batch(q, 5, 32, 4, order="canonical")
    bids, bsc = vo.bruteforce_batch(vo.COSINE, x, q, 5)
    np.savez_compressed(os.path.join(HERE, "regress_cos24.npz"), queries=q, ids=ids, dist_bits=d.view(np.uint32), counts=cnt,
                        stats=st[:, :4], sq8_ids=sids, sq8_dist_bits=sd.view(np.uint32), sq8_stats=sst[:, :4],
                        sq8_codes_crc=np.array([int(dp.quantizer.codes().astype(np.uint64).sum())], np.uint64),
                        sq8_min=dp.quantizer.min_vals, sq8_scale=dp.quantizer.scales, bf_ids=bids,
                        bf_score_bits=bsc.view(np.uint32))


if __name__ == "__main__":
    spec_fixture()
    regression_fixture()
    print("golden fixtures written to", HERE)
