#!/usr/bin/env python
"""bench.py -- the vector-search hot path on BASELINE.json's configs, one JSON line per run (rank 0).

  python bench.py [--config c2] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

  c2    (default, BASELINE configs[1], the headline) HNSW search 1M x 768 f32 cosine, k=10, ef_search=64, batch 1024
  c2sq8 the same workload through DualPrecisionHnsw (int8 traversal + exact re-rank, dual_precision.rs:284-325)
  c3    HNSW search 10M x 768, f16 storage, k=100, ef_search=256, batch 8192 per GPU
  c4    Hamming HNSW 50M x 1024-bit packed vectors, k=10, batch 4096 per GPU
  c5    hybrid: c2's index + BM25 over 1M documents, RRF top-10, batch 1024
  c1    brute-force cosine top-10 on 10K x 768 f32 (configs[0], the reference's CPU-runnable case), batch 1024
  (--n / --nq / --ef / --k scale a config down for smoke runs; the metric name carries the sizes.)

One "step" = one pass of the hot path over one batch of synthetic queries.  Fields:

  value         whole-job queries/s with queries and index resident in HBM; CUDA events on the launching stream, max
                over ranks; at N > 1 the step includes the library's own gather of every rank's top-k
  e2e           same metric through the host-pointer C ABI (pinned host buffers): every step copies its queries H2D
                and its results D2H inside the timed region.  Pipelined with veles_search_submit / _wait, two
                batches in flight (a serving loop); `sync_value` is the one-call-at-a-time veles_search_batch number
  roofline      algorithmic bytes of one launch (the kernel's own NDC / expansion counters) / its mean duration
                (CUDA events around every launch inside the timed region) vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (C++ restatement of the reference algorithm; Rust cannot be built here) on the box's host
                cores over the same graph and queries, bounded sample
  --impl reference   times that CPU arm alone.  It never loads libveles_b200.so: graph, vectors and queries come from
                frozen files (SHA-256 in the line) written once by `bench.py --prepare` in a child process.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # kind, n, dim, store, metric, k, ef, nq, M, efc, latent
    "c2": dict(kind="hnsw", n=1_000_000, dim=768, store="f32", metric="cosine", k=10, ef=64, nq=1024, M=32, efc=200, latent=24),
    "c2sq8": dict(kind="hnsw", n=1_000_000, dim=768, store="f32", metric="cosine", k=10, ef=64, nq=1024, M=32, efc=200,
                  latent=24, sq8=4),
    "c3": dict(kind="hnsw", n=10_000_000, dim=768, store="f16", metric="cosine", k=100, ef=256, nq=8192, M=32, efc=200,
               latent=24),
    # binary codes of the same latent-24 data as c2 / c3; ef_search 128 is the smallest beam that holds recall@10 >= 0.95
    # at 50M (profiles/README.md has the sweep, and the latent-48 rows where the same recall needs ef ~ 1500)
    "c4": dict(kind="hnsw", n=50_000_000, dim=1024, store="bin1", metric="hamming", k=10, ef=128, nq=4096, M=32, efc=200,
               latent=24),
    "c5": dict(kind="hybrid", n=1_000_000, dim=768, store="f32", metric="cosine", k=10, ef=128, nq=1024, M=32, efc=200,
               latent=24, docs=1_000_000, vocab=100_000),
    "c1": dict(kind="brute", n=10_000, dim=768, store="f32", metric="cosine", k=10, ef=0, nq=1024, M=32, efc=200, latent=24),
}
METRIC_ID = {"cosine": 0, "euclidean": 1, "dot": 2, "hamming": 3}
ELT_BYTES = {"f32": 4.0, "f16": 2.0, "bin1": 0.125}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--prepare", action="store_true", help="write the frozen files of this config and exit (child of --impl reference)")
    for name in ("n", "dim", "nq", "k", "ef", "M", "efc", "latent"):
        ap.add_argument("--" + name, type=int, default=None)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true", help="N > 1: split ONE batch of nq queries over the GPUs (strong scaling)")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N > 1: the library's peer-store gather, or torch NCCL")
    ap.add_argument("--pipe-depth", type=int, default=3, help="e2e: batches in flight through veles_search_submit / _wait")
    ap.add_argument("--hybrid-three-calls", action="store_true",
                    help="c5: time search, BM25 and RRF as three host-API calls (round 2's first form) instead of veles_hybrid_search_batch")
    ap.add_argument("--cache", default=os.environ.get("VELES_BENCH_CACHE", "/tmp/veles_bench_cache"))
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    for name in ("n", "dim", "nq", "k", "ef", "M", "efc", "latent"):
        if getattr(a, name) is not None:
            cfg[name] = getattr(a, name)
    cfg["name"] = a.config
    return a, cfg


def metric_name(cfg):
    def sz(n):
        return f"{n // 1_000_000}M" if n % 1_000_000 == 0 else (f"{n // 1000}K" if n % 1000 == 0 else str(n))

    n, d = sz(cfg["n"]), cfg["dim"]
    if cfg["kind"] == "brute":
        return f"bruteforce_qps_{n}x{d}_k{cfg['k']}"
    if cfg["kind"] == "hybrid":
        return f"hybrid_rrf_qps_{n}x{d}_bm25_{sz(cfg['docs'])}docs_k{cfg['k']}"
    if cfg.get("sq8"):
        return f"hnsw_sq8_search_qps_{n}x{d}_k{cfg['k']}_ef{cfg['ef']}"
    if cfg["store"] == "bin1":
        return f"hnsw_search_qps_{n}x{d}bit_hamming_k{cfg['k']}_ef{cfg['ef']}"
    if cfg["store"] == "f16":
        return f"hnsw_search_qps_{n}x{d}_f16_k{cfg['k']}_ef{cfg['ef']}"
    return f"hnsw_search_qps_{n}x{d}_k{cfg['k']}_ef{cfg['ef']}"


# --------------------------------------------------------------------------------------------------------------
# synthetic data
# --------------------------------------------------------------------------------------------------------------
def gen_data(torch, n, dim, latent, seed, device, out_dtype=None):
    """Embedding-like vectors: a `latent`-dim Gaussian pushed through a fixed random linear map plus isotropic noise,
    L2-normalised (why not the reference's hash generator: DESIGN.md section 6).  Deterministic per (seed, torch
    version, device type)."""
    g = torch.Generator(device=device)
    g.manual_seed(1234)
    w = torch.randn(latent, dim, generator=g, device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dim, device=device, dtype=out_dtype or torch.float32)
    step = 65536
    for i in range(0, n, step):
        m = min(step, n - i)
        z = torch.randn(m, latent, generator=g, device=device)
        e = torch.randn(m, dim, generator=g, device=device)
        x = z @ w + 0.5 * e
        out[i:i + m] = x / x.norm(dim=1, keepdim=True)
    return out


def data_chunks(torch, cfg, seed, device, count=None, chunk=262144):
    """Yields (first_row, rows) of the config's collection in the form set_rows_device takes: f32 rows (f32 / f16
    stores) or packed bits as uint8 [m, dim/8] (bit j of a vector = sign of coordinate j of the latent data: a
    SimHash-style binary code, LSB first)."""
    n = cfg["n"] if count is None else count
    g = torch.Generator(device=device)
    g.manual_seed(1234)
    w = torch.randn(cfg["latent"], cfg["dim"], generator=g, device=device)
    g.manual_seed(seed)
    weights = (2 ** torch.arange(8, device=device)).to(torch.int32)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        rows = torch.empty(m, cfg["dim"], device=device)
        for j in range(0, m, 65536):
            mm = min(65536, m - j)
            z = torch.randn(mm, cfg["latent"], generator=g, device=device)
            e = torch.randn(mm, cfg["dim"], generator=g, device=device)
            x = z @ w + 0.5 * e
            rows[j:j + mm] = x / x.norm(dim=1, keepdim=True)
        if cfg["store"] == "bin1":
            bits = (rows > 0).to(torch.int32).view(m, cfg["dim"] // 8, 8)
            yield i, (bits * weights).sum(dim=2).to(torch.uint8).contiguous()
        else:
            yield i, rows


def make_queries(torch, cfg, nq, seed, device):
    """f32 queries [nq, dim] as the ABI takes them (binary configs: {0, 1} lanes, thresholded at > 0.5 on the device)."""
    q = gen_data(torch, nq, cfg["dim"], cfg["latent"], seed, device)
    return (q > 0).float().contiguous() if cfg["store"] == "bin1" else q.contiguous()


def build_snapshot(torch, cfg, device, keep_host=False):
    """Collection generated on the device chunk by chunk, uploaded through veles_index_set_rows_d, graph built by the
    block-insertion builder.  Returns (snapshot, host vectors in store form or None, datagen_s, build_s)."""
    from velesdb_b200 import DeviceSnapshot

    t0 = time.time()
    snap = DeviceSnapshot.create(cfg["n"], cfg["dim"], METRIC_ID[cfg["metric"]], cfg["store"])
    host = None
    if keep_host:
        dt, width = {"f32": (np.float32, cfg["dim"]), "f16": (np.float16, cfg["dim"]), "bin1": (np.uint8, cfg["dim"] // 8)}[cfg["store"]]
        host = np.empty((cfg["n"], width), dtype=dt)
    for first, rows in data_chunks(torch, cfg, 7, device):
        snap.set_rows_device(first, rows, "bin1" if cfg["store"] == "bin1" else "f32")
        if keep_host:
            r = rows if cfg["store"] != "f16" else rows.to(torch.float16)
            host[first:first + rows.shape[0]] = r.cpu().numpy()
    torch.cuda.synchronize()
    t_gen = time.time() - t0
    t0 = time.time()
    if cfg["kind"] != "brute":
        snap.build_graph(cfg["M"], cfg["efc"])
    torch.cuda.synchronize()
    return snap, host, t_gen, time.time() - t0


def sha256_file(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        while True:
            b = f.read(1 << 24)
            if not b:
                break
            h.update(b)
    return h.hexdigest()


def graph_digest(layers):
    h = hashlib.sha256()
    for rp, cols in layers:
        h.update(np.ascontiguousarray(rp).tobytes())
        h.update(np.ascontiguousarray(cols).tobytes())
    return h.hexdigest()


def cache_dir(a, cfg):
    key = "_".join(str(cfg[k]) for k in ("name", "n", "dim", "store", "M", "efc", "latent")) + f"_nq{cfg['nq']}_v2"
    return os.path.join(a.cache, key)


def prepare(a, cfg):
    """Writes the frozen files of a config: `native_hnsw.graph` (format v1), the vectors (`native_hnsw.vectors` for f32,
    `vectors.npy` in store form otherwise), rank 0's queries, and meta.json with their SHA-256."""
    import torch

    from velesdb_b200 import _native as nv

    nv.init(0)
    dev = torch.device("cuda", 0)
    d = cache_dir(a, cfg)
    os.makedirs(d, exist_ok=True)
    snap, host, t_gen, t_build = build_snapshot(torch, cfg, dev, keep_host=cfg["store"] != "f32" or cfg["kind"] == "brute")
    if cfg["kind"] == "brute":
        np.save(os.path.join(d, "vectors.npy"), host)
        vec_file = "vectors.npy"
    elif cfg["store"] == "f32":
        snap.dump(d)
        vec_file = "native_hnsw.vectors"
    else:
        snap.dump_graph(d)
        np.save(os.path.join(d, "vectors.npy"), host)
        vec_file = "vectors.npy"
    q = make_queries(torch, cfg, cfg["nq"], 1_000_003, dev).cpu().numpy()
    np.save(os.path.join(d, "queries.npy"), q)
    meta = {"config": {k: v for k, v in cfg.items()}, "datagen_s": round(t_gen, 2), "build_s": round(t_build, 2),
            "entry_point": snap.entry_point, "max_layer": snap.max_layer,
            "sha256": {f: sha256_file(os.path.join(d, f))
                       for f in (["native_hnsw.graph"] if cfg["kind"] != "brute" else []) + [vec_file, "queries.npy"]}}
    with open(os.path.join(d, "meta.json"), "w") as f:
        json.dump(meta, f)
    print(json.dumps({"prepared": d, **meta}))


# --------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(cfg):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if it is this workload."""
    try:
        for j in json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))):
            if all(j["workload"].get(k) == cfg.get(k) for k in ("name", "n", "dim", "k", "ef", "nq")):
                return int(j["traffic_bytes"])
    except Exception:
        pass
    return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_text(cfg, world, strong):
    per = f"batch={cfg['nq']} queries" + (" split over the GPUs" if strong else " per GPU")
    if cfg["kind"] == "brute":
        return f"brute-force {cfg['metric']} top-{cfg['k']} on {cfg['n']}x{cfg['dim']} f32, {per}"
    base = (f"HNSW search {cfg['n']}x{cfg['dim']} {cfg['store']} {cfg['metric']}, k={cfg['k']}, ef_search={cfg['ef']}, {per}, "
            f"M={cfg['M']} M0={2 * cfg['M']}, graph built by the GPU block-insertion builder (ef_construction={cfg['efc']})")
    if cfg.get("sq8"):
        base += f", DualPrecisionHnsw int8 traversal + exact re-rank, oversampling {cfg['sq8']}"
    if cfg["kind"] == "hybrid":
        base += f" + BM25 top-{2 * cfg['k']} over {cfg['docs']} docs (vocab {cfg['vocab']}) + RRF top-{cfg['k']}"
    return base


def base_config(cfg, world, strong):
    gb = cfg["n"] * cfg["dim"] * ELT_BYTES[cfg["store"]] / 1e9
    return {"workload": workload_text(cfg, world, strong), "name": cfg["name"], "n": cfg["n"], "dim": cfg["dim"], "k": cfg["k"],
            "ef_search": cfg["ef"], "batch": cfg["nq"], "store": cfg["store"],
            "l2_policy": f"the collection ({gb:.1f} GB) is far larger than the 126 MB L2; no flush between steps"
                         if gb > 1.0 else "L2 flushed between steps by writing a 256 MB buffer",
            "parallelism": f"queries sharded over {world} GPU(s), index replicated",
            "data_generator": f"latent{cfg['latent']}-gaussian+0.5*noise, normalised, seed 7"
                              + ("; bits = sign of each coordinate" if cfg["store"] == "bin1" else "")}


# --------------------------------------------------------------------------------------------------------------
# BM25 corpus of config c5: Zipf(1.07) terms, lognormal document lengths
# --------------------------------------------------------------------------------------------------------------
def bm25_corpus(cfg, nq, seed_q):
    n_docs, vocab = cfg["docs"], cfg["vocab"]
    rng = np.random.default_rng(7)
    p = 1.0 / np.arange(1, vocab + 1) ** 1.07
    p /= p.sum()
    lens = np.clip(np.exp(rng.normal(np.log(120), 0.5, n_docs)).astype(np.int64), 8, 1024)
    total = int(lens.sum())
    toks = np.searchsorted(np.cumsum(p), rng.random(total)).astype(np.uint32)
    np.minimum(toks, vocab - 1, out=toks)
    doc_of = np.repeat(np.arange(n_docs, dtype=np.uint64), lens)
    uk, tf = np.unique((toks.astype(np.uint64) << np.uint64(32)) | doc_of, return_counts=True)
    post_doc = (uk & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    df = np.bincount((uk >> np.uint64(32)).astype(np.uint32), minlength=vocab).astype(np.uint32)
    term_ptr = np.zeros(vocab + 1, np.uint64)
    term_ptr[1:] = np.cumsum(df)
    rq = np.random.default_rng(seed_q)
    p2 = p.copy()
    p2[:50] = 0  # queries skip the 50 most frequent ranks
    p2 /= p2.sum()
    q_ptr, q_terms = [0], []
    for _ in range(nq):
        q_terms += rq.choice(vocab, size=int(rq.integers(3, 7)), p=p2).tolist()
        q_ptr.append(len(q_terms))
    return dict(term_ptr=term_ptr, post_doc=post_doc, tf=tf.astype(np.uint32), df=df, lens=lens.astype(np.uint32),
                n_docs=n_docs, total=total, toks=toks, q_ptr=np.array(q_ptr, np.uint32), q_terms=np.array(q_terms, np.uint32))


# --------------------------------------------------------------------------------------------------------------
# reference arm: the oracle on host cores, from frozen files; libveles_b200.so is never loaded in this process
# --------------------------------------------------------------------------------------------------------------
def run_reference(a, cfg):
    from oracle import oracle as vo  # the CPU arm: the one other place bench.py may execute oracle/

    ncores = os.cpu_count() or 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    d = cache_dir(a, cfg)
    if not os.path.exists(os.path.join(d, "meta.json")):
        cmd = [sys.executable, os.path.abspath(__file__), "--prepare", "--config", cfg["name"], "--cache", a.cache]
        for name in ("n", "dim", "nq", "k", "ef", "M", "efc", "latent"):
            cmd += ["--" + name, str(cfg[name])]
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, env=env)
    meta = json.load(open(os.path.join(d, "meta.json")))
    q = np.load(os.path.join(d, "queries.npy"))
    mid = METRIC_ID[cfg["metric"]]
    config = base_config(cfg, world, False)
    steps, warm = a.steps, max(1, a.warmup)

    if cfg["kind"] == "brute":
        xs = np.load(os.path.join(d, "vectors.npy"))
        sample = q[:min(len(q), 128)]
        fn = lambda: vo.bruteforce_batch(mid, xs, sample, cfg["k"], threads=ncores)
        nsamp, what = len(sample), f"{len(sample)} of the {cfg['nq']} queries per step, full scan, all host threads"
    else:
        if cfg["store"] == "f32":
            g = vo.open_index(d, mid)
        else:
            v = np.load(os.path.join(d, "vectors.npy"))
            g = vo.open_index(d, mid, vectors=v.view(np.uint64) if cfg["store"] == "bin1" else v, dim=cfg["dim"])
        # bounded sample per step: the whole batch for c2, the first 1024 queries of the larger batches
        nsamp = min(len(q), 1024)
        sample = np.ascontiguousarray(q[:nsamp])
        if cfg.get("sq8"):
            dp = vo.DualPrecisionHnsw.from_graph(g, train_count=1000)
            fn = lambda: dp.search_int8_batch(sample, cfg["k"], cfg["ef"], cfg["sq8"], order="canonical", threads=ncores)
        elif cfg["kind"] == "hybrid":
            corp = bm25_corpus(cfg, cfg["nq"], 11)
            ob = vo.Bm25()
            starts = np.concatenate([[0], np.cumsum(corp["lens"].astype(np.int64))])
            for doc in range(corp["n_docs"]):
                ob.add_document_terms(doc, corp["toks"][starts[doc]:starts[doc + 1]])
            nsamp = min(nsamp, 256)
            sample = np.ascontiguousarray(q[:nsamp])
            qp, qt = corp["q_ptr"][:nsamp + 1], corp["q_terms"][:corp["q_ptr"][nsamp]]

            def fn():
                vi, vd, vc, _ = g.search_batch(sample, 2 * cfg["k"], cfg["ef"], order="canonical", threads=ncores)
                ti, ts, tc = ob.search_batch_terms(qp, qt, 2 * cfg["k"], threads=ncores)
                for i in range(nsamp):
                    vo.rrf_hybrid(vi[i, :vc[i]], ti[i, :tc[i]], cfg["k"], 0.5)
        else:
            fn = lambda: g.search_batch(sample, cfg["k"], cfg["ef"], order="canonical", threads=ncores)
        what = f"the first {nsamp} of the {cfg['nq']} queries per step, all host threads, same frozen graph"
    for _ in range(warm):
        fn()
    t = time.time()
    for _ in range(steps):
        fn()
    el = time.time() - t
    qps = nsamp * steps / el
    line = {"metric": metric_name(cfg), "value": qps, "unit": "queries/s", "n_gpus": a.gpus, "steps": steps, "warmup": warm,
            "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": cfg["store"].replace("bin1", "u64"), "data": "synthetic", "impl": "reference", "config": config,
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": ncores, "kind": "port", "sample": what},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "frozen_files": {"dir": d, "sha256": meta["sha256"], "build_s": meta["build_s"]}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
def main():
    a, cfg = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.prepare:
        return prepare(a, cfg)
    if a.impl == "reference":
        if rank == 0:
            run_reference(a, cfg)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from velesdb_b200 import _native as nv

    nv.init(local)
    lib = nv.lib()
    kind = cfg["kind"]
    k, ef, dim = cfg["k"], cfg["ef"], cfg["dim"]
    # queries of this rank: weak scaling = nq per GPU with per-rank seeds; strong = one nq batch split contiguously
    if a.strong and use_dist:
        assert cfg["nq"] % world == 0
        nq = cfg["nq"] // world
        q_all = make_queries(torch, cfg, cfg["nq"], 1_000_003, dev)
        q_d = q_all[rank * nq:(rank + 1) * nq].contiguous()
    else:
        nq = cfg["nq"]
        q_d = make_queries(torch, cfg, nq, 1_000_003 + rank, dev)
    q_h = torch.empty((nq, dim), dtype=torch.float32).pin_memory()
    q_h.copy_(q_d)
    need_host = rank == 0 and not a.no_cpu_baseline and cfg["store"] != "f32"
    snap, x_host, t_gen, t_build = build_snapshot(torch, cfg, dev, keep_host=need_host)
    config = base_config(cfg, world, a.strong and use_dist)
    stream = torch.cuda.current_stream().cuda_stream
    sq8 = cfg.get("sq8", 0)
    if sq8:
        snap.attach_sq8(1000)

    ids_t = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dist_t = torch.empty((nq, k), dtype=torch.float32, device=dev)
    cnt_t = torch.empty(nq, dtype=torch.int32, device=dev)
    st_t = torch.zeros((nq, 4), dtype=torch.int32, device=dev)

    bm = corp = None
    if kind == "hybrid":
        from velesdb_b200 import Bm25Snapshot, hybrid_search_batch, rrf_hybrid_batch

        corp = bm25_corpus(cfg, nq, 11 + rank)
        bm = Bm25Snapshot(corp["term_ptr"], corp["post_doc"], corp["tf"], corp["df"], corp["lens"], corp["n_docs"], corp["total"])
    k_vec = 2 * k if kind == "hybrid" else k      # hybrid_search asks both sides for 2k (text.rs:137-140)
    if kind == "hybrid":
        ids_t = torch.empty((nq, k_vec), dtype=torch.int32, device=dev)
        dist_t = torch.empty((nq, k_vec), dtype=torch.float32, device=dev)

    def search_device(stats=None):
        if kind == "brute":
            snap.bruteforce_batch_device(q_d, k, ids_t, dist_t, stream)
        elif sq8:
            snap.search_batch_sq8_device(q_d, k, ef, sq8, ids_t, dist_t, cnt_t, stats, stream)
        else:
            snap.search_batch_device(q_d, k_vec, ef, ids_t, dist_t, cnt_t, stats, stream)

    # ---------------- untimed: results, stats, exact ground truth ----------------
    search_device(st_t if kind != "brute" else None)
    torch.cuda.synchronize()
    if kind != "brute":
        snap.search_status(stream)
    got, got_d = ids_t.cpu().numpy(), dist_t.cpu().numpy()
    recall = None
    if kind != "brute":
        ns = min(nq, 1024)
        gt_ids = torch.empty((ns, k_vec), dtype=torch.int32, device=dev)
        gt_sc = torch.empty((ns, k_vec), dtype=torch.float32, device=dev)
        snap.bruteforce_batch_device(q_d[:ns].contiguous(), k_vec, gt_ids, gt_sc, stream)
        torch.cuda.synchronize()
        gi, gs = gt_ids.cpu().numpy(), gt_sc.cpu().numpy()
        if cfg["store"] == "bin1":  # integer distances tie: a hit is any result at least as close as the exact k-th
            recall = float(np.mean(got_d[:ns, :k] <= gs[:, k - 1:k]))
        else:
            recall = float(np.mean([len(set(got[i, :k].tolist()) & set(gi[i, :k].tolist())) / k for i in range(ns)]))
    st = st_t.cpu().numpy().astype(np.int64)
    ndc, hops0, hops_up = st[:, 0] + st[:, 2], st[:, 1], st[:, 3]
    row_b = dim * ELT_BYTES[cfg["store"]]
    if kind == "brute":
        alg_bytes = int(cfg["n"] * row_b + nq * dim * 4 + nq * k * 8)             # DB charged once per batch
    elif sq8:
        alg_bytes = int((ndc * dim + hops0 * 2 * cfg["M"] * 4 + hops_up * cfg["M"] * 4 + dim * 4 + k * sq8 * dim * 4 + k * 8).sum())
    else:  # NDC*D*s + H*M0*4 + H_up*M*4 + D*4 + k*8 per query
        alg_bytes = int((ndc * row_b + hops0 * 2 * cfg["M"] * 4 + hops_up * cfg["M"] * 4 + dim * 4 + k_vec * 8).sum())

    # ---------------- multi-GPU gather inside the step ----------------
    comm = None
    hybrid_one = kind == "hybrid" and not a.hybrid_three_calls
    if use_dist and hybrid_one:
        gather_note = "none: every rank answers its own batch through veles_hybrid_search_batch (host buffers in and out)"
    elif use_dist:
        gather_note = None
        if a.gather == "p2p":
            from velesdb_b200.dist import PeerGather

            try:
                comm = PeerGather(snap, rank, world, nq, k_vec, dev)
            except RuntimeError as e:  # raised on every rank together (see PeerGather): fall back together
                gather_note = f"p2p gather unavailable ({e}); NCCL all-gather used"
        if comm is None and not hybrid_one:
            gath_ids = torch.empty((world * nq, k_vec), dtype=torch.int32, device=dev)
            gath_dist = torch.empty((world * nq, k_vec), dtype=torch.float32, device=dev)

    flush = None
    if cfg["n"] * row_b < 1e9:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        if flush is not None:
            flush.zero_()
        if comm is not None:
            comm.search_gather(q_d, ef, stream)
        else:
            search_device()
            if use_dist and not hybrid_one:
                dist.all_gather_into_tensor(gath_ids, ids_t)
                dist.all_gather_into_tensor(gath_dist, dist_t)
        if kind == "hybrid":
            torch.cuda.current_stream().synchronize()
            td, ts, tc = bm.search_batch(corp["q_ptr"], corp["q_terms"], k_vec)
            return rrf_hybrid_batch(ids_t.cpu().numpy().astype(np.uint32), cnt_t.cpu().numpy().astype(np.uint32), td, tc, k, 0.5)

    # hybrid: the product call is veles_hybrid_search_batch -- host buffers in, fused top-k out, both legs concurrent on
    # the device.  --hybrid-three-calls times round 2's first form (search, BM25 and RRF as three host-API calls).
    qn_host = q_h.numpy()

    def step_hybrid():
        return hybrid_search_batch(snap, bm, qn_host, corp["q_ptr"], corp["q_terms"], k, ef, 0.5, stream)

    three_ms = None
    if hybrid_one:
        three = step_device()
        one = step_hybrid()
        for _ in range(3):
            step_device()
        t3 = time.perf_counter()
        for _ in range(a.steps):
            step_device()
        three_ms = (time.perf_counter() - t3) * 1e3 / a.steps  # the three-call form on the same batch, for the record
        assert all(np.array_equal(x, y) for x, y in zip((three[0], three[1].view(np.uint32), three[2]),
                                                        (one[0], one[1].view(np.uint32), one[2]))), \
            "veles_hybrid_search_batch differs from the three separate calls"
        step_device = step_hybrid  # noqa: F811

    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()  # ncu --profile-from-start off: the launch list covers the timed regions only
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-launch events INSIDE the timed region (same stream as the launches): the roofline's duration
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = lib.veles_launch_count()
    t_wall = time.perf_counter()
    e0.record()
    for i in range(a.steps):
        if hybrid_one:
            kev[i][0].record()
            step_hybrid()
            kev[i][1].record()
            continue
        if flush is not None:
            flush.zero_()
        kev[i][0].record()
        if comm is not None:
            comm.search_gather(q_d, ef, stream, mid_event=kev[i][1])
        else:
            search_device()
            kev[i][1].record()
            if use_dist:
                dist.all_gather_into_tensor(gath_ids, ids_t)
                dist.all_gather_into_tensor(gath_dist, dist_t)
        if kind == "hybrid":
            torch.cuda.current_stream().synchronize()
            td, ts, tc = bm.search_batch(corp["q_ptr"], corp["q_terms"], k_vec)
            rrf_hybrid_batch(ids_t.cpu().numpy().astype(np.uint32), cnt_t.cpu().numpy().astype(np.uint32), td, tc, k, 0.5)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    if use_dist:
        dist.barrier()
    launches_timed = lib.veles_launch_count() - l0
    # hybrid steps leave the stream for BM25 + fusion through the host API: the step is timed by the wall clock
    ms_dev = wall_ms if kind == "hybrid" else e0.elapsed_time(e1)
    kernel_ms = float(np.mean([x.elapsed_time(y) for x, y in kev]))
    if comm is not None:
        assert comm.check(ids_t, dist_t, rank), "gathered window differs from the local results"

    # ---------------- timed: end to end through the host-pointer C ABI ----------------
    e2e = None
    if kind == "hnsw" and not sq8:
        bufs = []
        depth = max(1, min(a.pipe_depth, 8))
        for _ in range(depth):
            bufs.append((torch.empty((nq, k), dtype=torch.int32).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                         torch.empty(nq, dtype=torch.int32).pin_memory()))
        qn = q_h.numpy()

        def step_sync():
            o = bufs[0]
            nv.check(lib.veles_search_batch(snap.h, nv.ptr(qn), nq, k, ef, nv.ptr(o[0]), nv.ptr(o[1]), nv.ptr(o[2]), None, stream))

        def run_pipelined(steps):
            tickets = []
            for i in range(steps):
                o = bufs[i % depth]
                if len(tickets) == depth:
                    snap.search_wait(tickets.pop(0))
                tickets.append(snap.search_submit(q_h, k, ef, o[0], o[1], o[2]))
            for t in tickets:
                snap.search_wait(t)

        for _ in range(max(a.warmup, 5)):  # untimed: first touches of the pinned buffers, allocator and driver caches
            step_sync()
        run_pipelined(max(a.warmup, 5))
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        t = time.perf_counter()
        for _ in range(a.steps):
            step_sync()
        torch.cuda.synchronize()
        ms_sync = (time.perf_counter() - t) * 1e3
        if use_dist:
            dist.barrier()
        t = time.perf_counter()
        run_pipelined(a.steps)
        torch.cuda.synchronize()
        ms_pipe = (time.perf_counter() - t) * 1e3
        for o in bufs:
            assert np.array_equal(o[0].numpy(), got[:, :k]), "e2e results differ from the device-resident run"
        e2e = (ms_pipe, ms_sync)
    # brute-force config: the relaxed mode (tcgen05 fp16 candidate GEMM + exact f32 re-rank) on the
    # same batch, device-resident like `value`, with its agreement with the exact path -- reported beside the exact number
    relaxed = None
    if kind == "brute" and not use_dist:
        ri_t, rs_t = torch.empty_like(ids_t), torch.empty_like(dist_t)
        for _ in range(3):
            snap.bruteforce_batch_relaxed_device(q_d, k, 4, ri_t, rs_t, stream)
        torch.cuda.synchronize()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(a.steps):
            if flush is not None:
                flush.zero_()
            snap.bruteforce_batch_relaxed_device(q_d, k, 4, ri_t, rs_t, stream)
        r1.record()
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / a.steps
        ri = ri_t.cpu().numpy()
        relaxed = {"value": nq / (rms / 1e3), "unit": "queries/s", "ms_per_step": rms, "oversample": 4,
                   "api": "veles_bruteforce_batch_relaxed_d (L2 flush included in the step, as for `value`)",
                   "recall_vs_exact": float(np.mean([len(set(ri[i].tolist()) & set(got[i].tolist())) / k for i in range(nq)])),
                   "ids_identical_fraction": float((ri == got).all(axis=1).mean())}
    clocks = sampler.finish()
    torch.cuda.profiler.stop()

    if use_dist:
        tt = torch.tensor([ms_dev, e2e[0] if e2e else 0.0, e2e[1] if e2e else 0.0, kernel_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, kernel_ms = float(tt[0]), float(tt[3])
        if e2e:
            e2e = (float(tt[1]), float(tt[2]))
        if recall is not None:
            rr = torch.tensor([recall], device=dev)
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
            recall = float(rr[0]) / world

    total_q = nq * world * a.steps
    value = total_q / (ms_dev / 1e3)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    kname = {"brute": "bf_tile_smem_kernel + topk", "hybrid": "hnsw_search_kernel<f32>"}.get(kind, f"hnsw_search_kernel<{'sq8' if sq8 else cfg['store']}>")
    if hybrid_one:
        # both legs run concurrently inside one call, so the events bracket the call: charge it with both legs' bytes
        # (text leg: 12 bytes per posting of every query token + its 2k results)
        df_q = corp["df"][corp["q_terms"][corp["q_terms"] < len(corp["df"])]].astype(np.int64)
        alg_bytes += int(df_q.sum() * 12 + nq * k_vec * 8)
        achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
        kname = "veles_hybrid_search_batch: hnsw_search_kernel<f32> || bm25_sub_kernel, then rrf_hybrid_kernel (host copies included)"
    line = {"metric": metric_name(cfg), "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_dev / a.steps, "higher_is_better": True,
            "scaling": "strong" if (a.strong and use_dist) else "weak", "vs_baseline": None,
            "dtype": cfg["store"].replace("bin1", "u64"), "data": "synthetic", "config": config,
            "clocks": clocks, "gpu_launches": int(launches_timed), "build_s": round(t_build, 2), "datagen_s": round(t_gen, 2),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(cfg), "peak_source": peak_src, "kernel": kname, "kernel_ms": kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes}}
    if relaxed is not None:
        line["relaxed_mode"] = relaxed
    if recall is not None:
        line[f"recall_at_{k}"] = recall
        line["roofline"].update({"ndc_per_query": float(ndc.mean()), "ndc_max": int(ndc.max()),
                                 "ndc_p99": float(np.percentile(ndc, 99)), "expansions_per_query": float(hops0.mean())})
    if use_dist:
        line["gather"] = ("library: search epilogue stores each query's top-k into every peer's window over NVLink, flag barrier"
                          if comm is not None else (gather_note or "torch.distributed all_gather_into_tensor x2 (NCCL)"))
    if e2e:
        line["e2e"] = {"value": total_q / (e2e[0] / 1e3), "unit": "queries/s", "h2d_bytes_per_step": nq * dim * 4,
                       "d2h_bytes_per_step": nq * k * 8 + nq * 4 + 8, "ms_per_step": e2e[0] / a.steps,
                       "api": f"veles_search_submit / veles_search_wait, {depth} batches in flight, pinned host buffers",
                       "sync_value": total_q / (e2e[1] / 1e3), "sync_api": "veles_search_batch, one call at a time"}
    else:
        # these configs' public call is the host API already timed above (hybrid) or the device-pointer call
        line["e2e"] = {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "device-resident call only for this config"}
        if hybrid_one:
            line["e2e"].update({"h2d_bytes_per_step": int(nq * dim * 4 + corp["q_ptr"].nbytes + corp["q_terms"].nbytes),
                                "d2h_bytes_per_step": nq * k * 8 + nq * 4 + 8,
                                "api": "veles_hybrid_search_batch: pinned host queries + term ids in, fused top-k out, one call per step",
                                "note": "the timed step IS the host-API call; results equal the three separate calls bit for bit",
                                "three_calls_ms_per_step": three_ms})
        elif kind == "hybrid":
            line["e2e"].update({"h2d_bytes_per_step": int(nq * k_vec * 8 + nq * 8 + corp["q_ptr"].nbytes + corp["q_terms"].nbytes),
                                "d2h_bytes_per_step": nq * k_vec * 12 + nq * 8 + nq * k * 8 + nq * 4,
                                "api": "veles_search_batch_d + veles_bm25_search_batch + veles_rrf_hybrid (--hybrid-three-calls)",
                                "note": "hybrid steps already run BM25 + fusion through the host API"})

    # ---------------- CPU baseline: the oracle on the same graph and queries ----------------
    if rank == 0 and not a.no_cpu_baseline and not (a.strong and use_dist):
        from oracle import oracle as vo  # CPU baseline leg

        ncores = os.cpu_count() or 1
        mid = METRIC_ID[cfg["metric"]]
        qn = q_h.numpy()
        if kind == "brute":
            xs = snap.get_rows(0, cfg["n"], np.float32)
            ns = min(nq, 128)
            fn = lambda: vo.bruteforce_batch(mid, xs, qn[:ns], k, threads=ncores)
            ri, rs = fn()
            match = bool(np.array_equal(got[:ns], ri.astype(np.int32)) and np.array_equal(got_d[:ns].view(np.uint32), rs.view(np.uint32)))
        else:
            layers = snap.export_graph()
            line["graph_sha256"] = graph_digest(layers)
            if cfg["store"] == "f32":
                xs = snap.get_rows(0, cfg["n"], np.float32)
            else:
                xs = x_host.view(np.uint64) if cfg["store"] == "bin1" else x_host
            g = vo.frozen(mid, xs, layers, cfg["M"], 2 * cfg["M"], snap.entry_point, snap.max_layer, dim=dim)
            ns = min(nq, 1024 if kind == "hnsw" else 256)
            if sq8:
                dp = vo.DualPrecisionHnsw.from_graph(g, train_count=1000)
                fn = lambda: dp.search_int8_batch(qn[:ns], k, ef, sq8, order="canonical", threads=ncores)
            else:
                fn = lambda: g.search_batch(qn[:ns], k_vec, ef, order="canonical", threads=ncores)
            r = fn()
            oi, od, ost = r[0], r[1], r[3]
            keep = ost[:, 4] == 0 if ost is not None and ost.ndim == 2 and ost.shape[1] > 4 else np.ones(ns, bool)
            match = bool(np.array_equal(got[:ns][keep], oi[keep].astype(np.int32)) and
                         np.array_equal(got_d[:ns].view(np.uint32), od.view(np.uint32)))
            line["parity"] = {"queries": int(ns), "ids_and_distance_bits_equal_oracle": match, "ties_crossing_k": int((~keep).sum())}
        times = []
        t_all = time.time()
        while time.time() - t_all < a.cpu_seconds or not times:
            t = time.time()
            fn()
            times.append(time.time() - t)
        line["cpu_baseline"] = {"value": ns / float(np.median(times)), "unit": "queries/s", "cores": ncores, "kind": "port",
                                "sample": f"median of {len(times)} passes over the first {ns} queries of the batch, {ncores} threads, "
                                          "same graph and queries" + ("; vector-search leg only" if kind == "hybrid" else ""),
                                "ids_match_gpu": match}
    if rank == 0:
        print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
