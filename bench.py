#!/usr/bin/env python
"""bench.py -- HNSW search QPS (+ recall@10) on BASELINE.json configs[1]:
1M x 768-D fp32, cosine, k=10, ef_search=64, batch=1024 queries per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path (greedy descent + layer-0 beam + top-k) over one batch of
1024 synthetic queries.  Prints ONE JSON line (rank 0).  See DESIGN.md section 6 for every field.

  value         whole-job queries/s with queries and the index resident in HBM (CUDA events on the
                launching stream, max over ranks)
  e2e           same metric through the C ABI with HOST buffers (veles_search_batch: H2D of the
                queries, kernel, D2H of ids/distances/counts inside the timed region)
  roofline      algorithmic bytes of the search kernel / its measured duration vs MEASURED_PEAKS.json
  cpu_baseline  the oracle (C++ restatement of the reference algorithm; the Rust reference cannot be
                built here) on the box's host cores, same graph, same queries, bounded sample
  --impl reference   times that CPU restatement alone (all host threads) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "hnsw_search_qps_1Mx768_k10_ef64"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nq", type=int, default=1024)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--ef", type=int, default=64)
    ap.add_argument("--M", type=int, default=32)
    ap.add_argument("--latent", type=int, default=24)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def gen_data(torch, n, dim, latent, seed, device):
    """Synthetic 768-D embeddings-like vectors: a `latent`-dim Gaussian pushed through a fixed random
    linear map plus small isotropic noise, L2-normalised (why not the
    reference's hash generator).  Deterministic per (seed, torch version, device type)."""
    g = torch.Generator(device=device)
    g.manual_seed(1234)
    w = torch.randn(latent, dim, generator=g, device=device)
    g.manual_seed(seed)
    out = torch.empty(n, dim, device=device)
    step = 65536
    for i in range(0, n, step):
        m = min(step, n - i)
        z = torch.randn(m, latent, generator=g, device=device)
        e = torch.randn(m, dim, generator=g, device=device)
        x = z @ w + 0.5 * e
        out[i:i + m] = x / x.norm(dim=1, keepdim=True)
    return out


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


def ncu_traffic(a):
    """dram bytes per launch of the search kernel from the committed ncu capture, if it is this workload."""
    p = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
    key = f"n={a.n},dim={a.dim},k={a.k},ef={a.ef},nq={a.nq},metric=cosine,dtype=f32"
    try:
        j = json.load(open(p))
        return int(j["traffic_bytes"]) if j["workload_key"] == key else None
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def main():
    a = parse()
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and rank != 0:
        return  # the CPU arm runs on rank 0 only
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    use_dist = world > 1 and a.impl == "ours"
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from velesdb_b200 import DeviceSnapshot, DistanceMetric
    from velesdb_b200 import _native as nv

    nv.init(local)
    t0 = time.time()
    x_d = gen_data(torch, a.n, a.dim, a.latent, 7, dev)
    x_h = x_d.cpu().numpy()
    # queries: fresh draws from the same distribution (not database points); per-rank seeds (weak scaling)
    q_d = gen_data(torch, a.nq, a.dim, a.latent, 1_000_003 + rank, dev).contiguous()
    q_h = torch.empty((a.nq, a.dim), dtype=torch.float32).pin_memory()
    q_h.copy_(q_d)
    del x_d
    torch.cuda.empty_cache()
    snap = DeviceSnapshot.from_vectors(x_h, DistanceMetric.Cosine)
    t_gen = time.time() - t0
    t0 = time.time()
    snap.build_graph(a.M)
    torch.cuda.synchronize()
    t_build = time.time() - t0

    workload = (f"HNSW search {a.n}x{a.dim} fp32 cosine, k={a.k}, ef_search={a.ef}, batch={a.nq} queries per GPU, "
                f"M={a.M} M0={2 * a.M}, graph built by the GPU bulk builder")
    config = {"workload": workload, "n": a.n, "dim": a.dim, "k": a.k, "ef_search": a.ef, "batch": a.nq,
              "l2_policy": "index (3.1 GB) is far larger than the 126 MB L2; no flush between steps",
              "parallelism": f"queries sharded over {world} GPU(s), index replicated", "build_s": round(t_build, 2),
              "datagen_s": round(t_gen, 2), "data_generator": f"latent{a.latent}-gaussian+0.5*noise, normalised, seed 7"}

    # ---------------- exact ground truth + stats (untimed) ----------------
    ids_t = torch.empty((a.nq, a.k), dtype=torch.int32, device=dev)
    dist_t = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
    cnt_t = torch.empty(a.nq, dtype=torch.int32, device=dev)
    st_t = torch.empty((a.nq, 4), dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    snap.search_batch_device(q_d, a.k, a.ef, ids_t, dist_t, cnt_t, st_t, stream)
    torch.cuda.synchronize()
    gt_ids = torch.empty((a.nq, a.k), dtype=torch.int32, device=dev)
    gt_sc = torch.empty((a.nq, a.k), dtype=torch.float32, device=dev)
    snap.bruteforce_batch_device(q_d, a.k, gt_ids, gt_sc, stream)
    torch.cuda.synchronize()
    got, gt = ids_t.cpu().numpy(), gt_ids.cpu().numpy()
    recall = float(np.mean([len(set(got[i].tolist()) & set(gt[i].tolist())) / a.k for i in range(a.nq)]))
    st = st_t.cpu().numpy().astype(np.int64)
    ndc, hops0, hops_up = st[:, 0] + st[:, 2], st[:, 1], st[:, 3]
    # NDC*D*s + H*M0*4 + H_up*M*4 + D*4 + k*8 per query
    alg_bytes = int((ndc * a.dim * 4 + hops0 * 2 * a.M * 4 + hops_up * a.M * 4 + a.dim * 4 + a.k * 8).sum())

    graph_layers = None
    if a.impl == "reference" or not a.no_cpu_baseline:
        graph_layers = snap.export_graph() if rank == 0 else None

    def cpu_run(seconds, threads):
        from oracle import oracle as vo  # CPU baseline leg: the one place bench.py may execute oracle/

        g = vo.Hnsw.from_arrays(vo.COSINE, x_h, graph_layers, a.M, 2 * a.M, snap.entry_point, snap.max_layer)
        g.search_batch(q_h.numpy()[:64], a.k, a.ef, threads=threads)  # warm
        times, done = [], 0
        t_all = time.time()
        while time.time() - t_all < seconds or not times:
            t = time.time()
            oi, od, oc, _ = g.search_batch(q_h.numpy(), a.k, a.ef, threads=threads)
            times.append(time.time() - t)
            done += a.nq
        return a.nq / float(np.median(times)), oi, len(times)

    ncores = os.cpu_count() or 1

    if a.impl == "reference":
        # CPU arm: each step = one pass over the same 1024-query batch with all host threads
        from oracle import oracle as vo

        g = vo.Hnsw.from_arrays(vo.COSINE, x_h, graph_layers, a.M, 2 * a.M, snap.entry_point, snap.max_layer)
        qn = q_h.numpy()
        for _ in range(max(1, min(a.warmup, 2))):
            g.search_batch(qn, a.k, a.ef, threads=ncores)
        steps = min(a.steps, 10)
        t = time.time()
        for _ in range(steps):
            g.search_batch(qn, a.k, a.ef, threads=ncores)
        el = time.time() - t
        qps = a.nq * steps / el
        line = {"metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": a.gpus, "steps": steps, "warmup": a.warmup,
                "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config, "recall_at_10": recall,
                "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": ncores, "kind": "port",
                                 "sample": f"{steps} passes over the full {a.nq}-query batch"},
                "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------- timed: device resident ----------------
    launches0 = nv.lib().veles_launch_count()

    def step_device():
        snap.search_batch_device(q_d, a.k, a.ef, ids_t, dist_t, cnt_t, None, stream)
        if use_dist:
            dist.all_gather_into_tensor(gath_ids, ids_t)
            dist.all_gather_into_tensor(gath_dist, dist_t)

    if use_dist:
        gath_ids = torch.empty((world * a.nq, a.k), dtype=torch.int32, device=dev)
        gath_dist = torch.empty((world * a.nq, a.k), dtype=torch.float32, device=dev)
    for _ in range(max(a.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    if use_dist:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-launch events INSIDE the timed region (same stream as the launches): the roofline's duration
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    l0 = nv.lib().veles_launch_count()
    e0.record()
    for i in range(a.steps):
        kev[i][0].record()
        snap.search_batch_device(q_d, a.k, a.ef, ids_t, dist_t, cnt_t, None, stream)
        kev[i][1].record()
        if use_dist:
            dist.all_gather_into_tensor(gath_ids, ids_t)
            dist.all_gather_into_tensor(gath_dist, dist_t)
    e1.record()
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    launches_timed = nv.lib().veles_launch_count() - l0
    ms_dev = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([x.elapsed_time(y) for x, y in kev]))

    # ---------------- timed: end to end through the host-pointer C ABI ----------------
    # pinned host buffers, as a serving front-end would hold (pageable memory also works, slower D2H)
    out_ids_t = torch.empty((a.nq, a.k), dtype=torch.int32).pin_memory()
    out_dist_t = torch.empty((a.nq, a.k), dtype=torch.float32).pin_memory()
    out_cnt_t = torch.empty(a.nq, dtype=torch.int32).pin_memory()
    out_ids, out_dist, out_cnt = out_ids_t.numpy(), out_dist_t.numpy(), out_cnt_t.numpy()
    qn = q_h.numpy()

    def step_e2e():
        nv.check(nv.lib().veles_search_batch(snap.h, nv.ptr(qn), a.nq, a.k, a.ef, nv.ptr(out_ids), nv.ptr(out_dist),
                                             nv.ptr(out_cnt), None, stream))

    for _ in range(max(a.warmup, 5)):  # untimed: first touches of the pinned buffers, allocator and driver caches
        step_e2e()
    torch.cuda.synchronize()
    if use_dist:
        dist.barrier()
    t = time.perf_counter()
    for _ in range(a.steps):
        step_e2e()
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t) * 1e3
    clocks = sampler.finish()
    assert np.array_equal(out_ids.astype(np.int32), got), "e2e results differ from the device-resident run"

    if use_dist:
        tt = torch.tensor([ms_dev, ms_e2e], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e = float(tt[0]), float(tt[1])
        rr = torch.tensor([recall], device=dev)
        dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        recall = float(rr[0]) / world

    total_q = a.nq * world * a.steps
    value = total_q / (ms_dev / 1e3)
    e2e_v = total_q / (ms_e2e / 1e3)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "recall_at_10": recall,
            "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": "queries/s", "h2d_bytes_per_step": a.nq * a.dim * 4,
                    "d2h_bytes_per_step": a.nq * a.k * 8 + a.nq * 4 + 8, "ms_per_step": ms_e2e / a.steps},
            "gpu_launches": int(launches_timed),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(a), "peak_source": peak_src, "kernel": "hnsw_search_kernel<f32>",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "ndc_per_query": float(ndc.mean()), "ndc_max": int(ndc.max()),
                         "ndc_p99": float(np.percentile(ndc, 99)), "expansions_per_query": float(hops0.mean())}}
    if rank == 0 and not a.no_cpu_baseline:
        v, oi, passes = cpu_run(a.cpu_seconds, ncores)
        par = bool(np.array_equal(oi.astype(np.int32), got))
        line["cpu_baseline"] = {"value": v, "unit": "queries/s", "cores": ncores, "kind": "port",
                                "sample": f"median of {passes} passes over the full {a.nq}-query batch, "
                                          f"{ncores} threads, same graph and queries",
                                "ids_match_gpu": par}
    if rank == 0:
        print(json.dumps(line))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
