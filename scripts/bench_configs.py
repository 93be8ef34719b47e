#!/usr/bin/env python
"""Measures the BASELINE.json configs other than the headline one (which bench.py owns) on one B200:

  C1  brute-force cosine top-10, 10K x 768 f32, Q = 1 and Q = 1024 (the reference's CPU-runnable case)
  C3s HNSW search, f16 storage, k=100, ef=256, batch 8192 (N scaled to what the O(N^2) builder does in ~1 min)
  C4s Hamming HNSW on packed 1024-bit vectors, k=10, ef=64, batch 4096 (N scaled likewise)
  C5  hybrid: 1M x 768 vector index + 1M-doc BM25 index, top-10 RRF, batch 1024

Each config prints one JSON line with device-timed throughput, the algorithmic-bytes roofline fraction,
recall where it applies, parity against the CPU oracle on a sample, and the oracle's CPU throughput.
Dev/evidence tool: results are copied into profiles/.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import gen_data, measured_peak  # noqa: E402
from oracle import oracle as vo  # noqa: E402  (checker + CPU baseline only)
from velesdb_b200 import Bm25Snapshot, DeviceSnapshot, DistanceMetric, rrf_hybrid_batch  # noqa: E402
from velesdb_b200 import _native as nv  # noqa: E402

DEV = torch.device("cuda", 0)
PEAK, PEAK_SRC = measured_peak()
NCORES = os.cpu_count() or 1


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def recall(a, b, k):
    return float(np.mean([len(set(a[i].tolist()) & set(b[i].tolist())) / k for i in range(a.shape[0])]))


def c1():
    n, dim, k = 10_000, 768, 10
    x = gen_data(torch, n, dim, 24, 7, DEV).cpu().numpy()
    snap = DeviceSnapshot.from_vectors(x, DistanceMetric.Cosine)
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    for nq in (1, 1024):
        q_d = gen_data(torch, nq, dim, 24, 99, DEV).contiguous()
        ids = torch.empty((nq, k), dtype=torch.int32, device=DEV)
        sc = torch.empty((nq, k), dtype=torch.float32, device=DEV)
        ms = timed(lambda: snap.bruteforce_batch_device(q_d, k, ids, sc, stream), 20)
        alg = n * dim * 4 + nq * dim * 4 + nq * k * 8
        ri, rs = vo.bruteforce_batch(vo.COSINE, x, q_d.cpu().numpy(), k, threads=NCORES)
        par = bool(np.array_equal(ids.cpu().numpy(), ri.astype(np.int32)) and
                   np.array_equal(sc.cpu().numpy().view(np.uint32), rs.view(np.uint32)))
        t = time.time()
        reps = 3 if nq > 1 else 20
        for _ in range(reps):
            vo.bruteforce_batch(vo.COSINE, x, q_d.cpu().numpy(), k, threads=NCORES if nq > 1 else 1)
        cpu_qps = nq * reps / (time.time() - t)
        out[f"Q{nq}"] = {"ms": ms, "queries_per_s": nq / ms * 1e3, "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / PEAK,
                         "pair_evals_per_s": n * nq / ms * 1e3, "bit_exact_vs_oracle": par,
                         "cpu_queries_per_s": cpu_qps, "cpu_threads": NCORES if nq > 1 else 1}
    print(json.dumps({"config": "C1 brute-force cosine top-10, 10K x 768 f32", "peak_GBps": PEAK, **out}), flush=True)


def hnsw_case(name, n, dim, store, metric, k, ef, nq, latent, binary=False, M=32, sample=128):
    t0 = time.time()
    x = gen_data(torch, n, dim, latent, 7, DEV)
    q = gen_data(torch, nq, dim, latent, 1_000_003, DEV)
    if binary:
        x = (x > 0).float()
        q = (q > 0).float()
    x_h = x.cpu().numpy()
    q_d = q.contiguous()
    del x
    torch.cuda.empty_cache()
    build = DeviceSnapshot.from_vectors(x_h, metric)           # f32 snapshot for the builder
    build.build_graph(M)
    torch.cuda.synchronize()
    t_build = time.time() - t0
    layers = build.export_graph()
    if store == "f32":
        snap = build
    else:
        snap = DeviceSnapshot.from_arrays(x_h, metric, layers, M, 2 * M, build.entry_point, build.max_layer,
                                          store_dtype=store)
        del build
    stream = torch.cuda.current_stream().cuda_stream
    ids = torch.empty((nq, k), dtype=torch.int32, device=DEV)
    dist = torch.empty((nq, k), dtype=torch.float32, device=DEV)
    cnt = torch.empty(nq, dtype=torch.int32, device=DEV)
    st = torch.empty((nq, 4), dtype=torch.int32, device=DEV)
    snap.search_batch_device(q_d, k, ef, ids, dist, cnt, st, stream)
    torch.cuda.synchronize()
    ms = timed(lambda: snap.search_batch_device(q_d, k, ef, ids, dist, cnt, None, stream), 10)
    gi = torch.empty((nq, k), dtype=torch.int32, device=DEV)
    gs = torch.empty((nq, k), dtype=torch.float32, device=DEV)
    snap.bruteforce_batch_device(q_d, k, gi, gs, stream)
    torch.cuda.synchronize()
    got = ids.cpu().numpy()
    # recall on distances for integer metrics (ties make id sets ambiguous), on ids otherwise
    if binary:
        rec = float(np.mean(dist.cpu().numpy() <= gs.cpu().numpy()[:, -1:]))
    else:
        rec = recall(got, gi.cpu().numpy(), k)
    s = st.cpu().numpy().astype(np.int64)
    elt = {"f32": dim * 4, "f16": dim * 2, "bin1": dim // 8}[store]
    alg = int(((s[:, 0] + s[:, 2]) * elt + s[:, 1] * 2 * M * 4 + s[:, 3] * M * 4 + dim * 4 + k * 8).sum())
    # parity + CPU baseline on a sample, same graph
    xo = x_h.astype(np.float16).astype(np.float32) if store == "f16" else x_h
    g = vo.Hnsw.from_arrays(int(metric), xo, layers, M, 2 * M, snap.entry_point, snap.max_layer)
    qs = q_d[:sample].cpu().numpy()
    t = time.time()
    oi, od, oc, ost = g.search_batch(qs, k, ef, order="canonical", threads=NCORES)
    cpu_qps = sample / (time.time() - t)
    keep = ost[:, 4] == 0
    par_ids = bool(np.array_equal(got[:sample][keep], oi[keep].astype(np.int32)))
    par_d = bool(np.array_equal(dist[:sample].cpu().numpy().view(np.uint32), od.view(np.uint32)))
    print(json.dumps({"config": name, "n": n, "dim": dim, "store": store, "k": k, "ef": ef, "batch": nq, "M": M,
                      "build_s": round(t_build, 1), "ms_per_batch": ms, "queries_per_s": nq / ms * 1e3,
                      "recall_at_k": rec, "ndc_per_query": float((s[:, 0] + s[:, 2]).mean()),
                      "alg_GBps": alg / ms / 1e6, "frac_hbm": alg / ms / 1e6 / PEAK, "peak_GBps": PEAK,
                      "parity_ids_sample": par_ids, "parity_dist_bits_sample": par_d, "ties_crossing_k": int((~keep).sum()),
                      "cpu_queries_per_s": cpu_qps, "cpu_threads": NCORES, "cpu_sample": sample}), flush=True)
    return snap, x_h, q_d


def c5(n_docs=1_000_000, vocab=100_000, nq=1024, k=10):
    snap, x_h, q_d = hnsw_case("C2 (for C5) HNSW 1M x 768 f32", 1_000_000, 768, "f32", DistanceMetric.Cosine, 20, 128, nq, 24,
                               sample=64)
    # synthetic corpus: Zipf(1.07) over `vocab` terms, lognormal doc lengths
    t0 = time.time()
    rng = np.random.default_rng(7)
    p = 1.0 / np.arange(1, vocab + 1) ** 1.07
    p /= p.sum()
    lens = np.clip(np.exp(rng.normal(np.log(120), 0.5, n_docs)).astype(np.int64), 8, 1024)
    total = int(lens.sum())
    cdf = np.cumsum(p)
    toks = np.searchsorted(cdf, rng.random(total)).astype(np.uint32)
    np.minimum(toks, vocab - 1, out=toks)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.uint64), lens)
    key = (toks.astype(np.uint64) << np.uint64(32)) | doc_of          # sort by (term, doc)
    uk, tf = np.unique(key, return_counts=True)
    post_doc = (uk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    terms = (uk >> np.uint64(32)).astype(np.uint32)
    df = np.bincount(terms, minlength=vocab).astype(np.uint32)
    term_ptr = np.zeros(vocab + 1, np.uint64)
    term_ptr[1:] = np.cumsum(df)
    bm = Bm25Snapshot(term_ptr, post_doc, tf.astype(np.uint32), df, lens.astype(np.uint32), n_docs, total)
    t_corpus = time.time() - t0
    # queries: 3-6 terms from the same Zipf, skipping the 50 most frequent ranks
    p2 = p.copy()
    p2[:50] = 0
    p2 /= p2.sum()
    q_ptr, q_terms = [0], []
    for _ in range(nq):
        q_terms += rng.choice(vocab, size=int(rng.integers(3, 7)), p=p2).tolist()
        q_ptr.append(len(q_terms))
    q_ptr, q_terms = np.array(q_ptr, np.uint32), np.array(q_terms, np.uint32)
    in_k = 2 * k
    stream = torch.cuda.current_stream().cuda_stream
    vi = torch.empty((nq, in_k), dtype=torch.int32, device=DEV)
    vd = torch.empty((nq, in_k), dtype=torch.float32, device=DEV)
    vc = torch.empty(nq, dtype=torch.int32, device=DEV)
    ef = int(nv.lib().veles_ef_search(nv.BALANCED, in_k, 0))  # index.search(vec, 2k) is Balanced (text.rs:137)

    def step():
        snap.search_batch_device(q_d, in_k, ef, vi, vd, vc, None, stream)
        torch.cuda.synchronize()
        td, ts, tc = bm.search_batch(q_ptr, q_terms, in_k)
        return rrf_hybrid_batch(vi.cpu().numpy().astype(np.uint32), vc.cpu().numpy().astype(np.uint32), td, tc, k, 0.5), (td, ts, tc)

    for _ in range(2):
        (fi, fs, fc), (td, ts, tc) = step()
    t = time.time()
    reps = 5
    for _ in range(reps):
        step()
    ms = (time.time() - t) / reps * 1e3
    t = time.time()
    for _ in range(reps):
        bm.search_batch(q_ptr, q_terms, in_k)
    ms_bm = (time.time() - t) / reps * 1e3
    df_q = df[q_terms].astype(np.int64)
    alg_bm = int(df_q.sum() * 8 + nq * in_k * 8)
    if os.environ.get("C5_SKIP_ORACLE"):
        print(json.dumps({"config": "C5 (no oracle leg)", "hybrid_ms_per_batch_host_api": ms, "hybrid_queries_per_s": nq / ms * 1e3,
                          "bm25_ms_per_batch_host_api": ms_bm, "bm25_queries_per_s": nq / ms_bm * 1e3,
                          "bm25_alg_GBps_incl_copies": alg_bm / ms_bm / 1e6}), flush=True)
        return
    # oracle on a sample of queries (full corpus)
    t0 = time.time()
    ob = vo.Bm25()
    starts = np.concatenate([[0], np.cumsum(lens)])
    for d in range(n_docs):
        ob.add_document_terms(d, toks[starts[d]:starts[d + 1]])
    t_oracle_build = time.time() - t0
    sample = 64
    t = time.time()
    oi, os_, oc = ob.search_batch_terms(q_ptr[:sample + 1], q_terms[:q_ptr[sample]], in_k, threads=NCORES)
    cpu_bm_qps = sample / (time.time() - t)
    par = bool(np.array_equal(td[:sample], oi.astype(np.uint32)) and np.array_equal(ts[:sample].view(np.uint32), os_.view(np.uint32)))
    vin = vi.cpu().numpy()
    ok_rrf = True
    for qi in range(sample):
        ri, rs = vo.rrf_hybrid(vin[qi, :int(vc[qi])], td[qi, :int(tc[qi])], k, 0.5)
        ok_rrf &= bool(np.array_equal(fi[qi, :fc[qi]], ri.astype(np.uint32)) and np.array_equal(fs[qi, :fc[qi]].view(np.uint32), rs.view(np.uint32)))
    print(json.dumps({"config": "C5 hybrid vector + BM25, top-10 RRF, batch 1024", "n_docs": n_docs, "vocab": vocab,
                      "postings": int(post_doc.size), "corpus_build_s": round(t_corpus, 1),
                      "hybrid_ms_per_batch_host_api": ms, "hybrid_queries_per_s": nq / ms * 1e3,
                      "bm25_ms_per_batch_host_api": ms_bm, "bm25_queries_per_s": nq / ms_bm * 1e3,
                      "bm25_alg_bytes_per_batch": alg_bm, "bm25_alg_GBps_incl_copies": alg_bm / ms_bm / 1e6,
                      "bm25_bit_exact_vs_oracle_sample": par, "rrf_bit_exact_vs_oracle_sample": bool(ok_rrf),
                      "cpu_bm25_queries_per_s": cpu_bm_qps, "cpu_threads": NCORES, "oracle_corpus_build_s": round(t_oracle_build, 1)}),
          flush=True)


def sq8(n=1_000_000, dim=768, nq=1024, k=10, ef=64, over=4, sample=128):
    """C2 through DualPrecisionHnsw::search_with_config (dual_precision.rs:284-325): int8 traversal + exact re-rank."""
    snap, x_h, q_d = hnsw_case("C2 f32 traversal (for comparison)", n, dim, "f32", DistanceMetric.Cosine, k, ef, nq, 24, sample=sample)
    t0 = time.time()
    snap.attach_sq8(1000)
    torch.cuda.synchronize()
    t_attach = time.time() - t0
    stream = torch.cuda.current_stream().cuda_stream
    out = {}
    for batch in sorted({nq, 4 * nq}):
        qb = q_d if batch == nq else gen_data(torch, batch, dim, 24, 1_000_003, DEV).contiguous()
        ids = torch.empty((batch, k), dtype=torch.int32, device=DEV)
        dist = torch.empty((batch, k), dtype=torch.float32, device=DEV)
        cnt = torch.empty(batch, dtype=torch.int32, device=DEV)
        st = torch.empty((batch, 4), dtype=torch.int32, device=DEV)
        snap.search_batch_sq8_device(qb, k, ef, over, ids, dist, cnt, st, stream)
        torch.cuda.synchronize()
        ms = timed(lambda: snap.search_batch_sq8_device(qb, k, ef, over, ids, dist, cnt, None, stream), 20)
        gi = torch.empty((batch, k), dtype=torch.int32, device=DEV)
        gs = torch.empty((batch, k), dtype=torch.float32, device=DEV)
        snap.bruteforce_batch_device(qb, k, gi, gs, stream)
        torch.cuda.synchronize()
        s = st.cpu().numpy().astype(np.int64)
        ck = k * over
        alg = int(((s[:, 0] + s[:, 2]) * dim + s[:, 1] * 64 * 4 + s[:, 3] * 32 * 4 + dim * 4 + ck * dim * 4 + k * 8).sum())
        out[f"batch{batch}"] = {"ms_per_batch": ms, "queries_per_s": batch / ms * 1e3,
                                "recall_at_k": recall(ids.cpu().numpy(), gi.cpu().numpy(), k),
                                "ndc_per_query": float((s[:, 0] + s[:, 2]).mean()), "alg_GBps": alg / ms / 1e6,
                                "frac_hbm": alg / ms / 1e6 / PEAK}
        if batch == nq:
            got_i, got_d = ids.cpu().numpy(), dist.cpu().numpy()
    g = vo.Hnsw.from_arrays(int(DistanceMetric.Cosine), x_h, snap.export_graph(), 32, 64, snap.entry_point, snap.max_layer)
    dp = vo.DualPrecisionHnsw.from_graph(g, train_count=1000)
    qs = q_d[:sample].cpu().numpy()
    t = time.time()
    oi, od, oc, ost = dp.search_int8_batch(qs, k, ef, over, order="canonical", threads=NCORES)
    cpu_qps = sample / (time.time() - t)
    print(json.dumps({"config": "C2-sq8 DualPrecisionHnsw int8 traversal + exact re-rank", "n": n, "dim": dim, "k": k,
                      "ef_search": ef, "oversampling": over, "attach_s": round(t_attach, 2), **out,
                      "parity_ids_sample": bool(np.array_equal(got_i[:sample], oi.astype(np.int32))),
                      "parity_dist_bits_sample": bool(np.array_equal(got_d[:sample].view(np.uint32), od.view(np.uint32))),
                      "cpu_queries_per_s": cpu_qps, "cpu_threads": NCORES, "cpu_sample": sample, "peak_GBps": PEAK}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c3", "c4", "c5"])
    ap.add_argument("--n3", type=int, default=2_000_000)
    ap.add_argument("--n4", type=int, default=2_000_000)
    a = ap.parse_args()
    nv.init(0)
    if "c1" in a.which:
        c1()
    if "c3" in a.which:
        hnsw_case("C3s HNSW f16 store, k=100, ef=256, batch 8192", a.n3, 768, "f16", DistanceMetric.Cosine, 100, 256, 8192, 24)
        torch.cuda.empty_cache()
    if "c4" in a.which:
        hnsw_case("C4s Hamming HNSW packed 1024-bit, k=10, ef=64, batch 4096", a.n4, 1024, "bin1", DistanceMetric.Hamming, 10, 64,
                  4096, 48, binary=True)
        torch.cuda.empty_cache()
    if "c5" in a.which:
        c5()
    if "sq8" in a.which:
        sq8()
