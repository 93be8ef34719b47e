"""Multi-GPU runtime for the search path: one process per GPU, the immutable snapshot replicated,
the query batch sharded by contiguous slices, and ONE collective -- the final gather of
[nq_local, k] ids + distances.  The reference has a single strategy too:
rayon ``par_iter`` over queries inside one process (index/hnsw/index/batch.rs:178-196).

``torch.distributed`` is plumbing only (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(nq: int, world: int, rank: int):
    """Contiguous, balanced slices: the first ``nq % world`` ranks get one extra query."""
    base, extra = divmod(nq, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_topk(ids: torch.Tensor, vals: torch.Tensor, counts: torch.Tensor, nq: int, group=None):
    """All-gathers per-rank [nq_local, k] results into [nq, k] tensors, in query order, on every rank.
    Slices may differ by one row, so ranks pad to the largest slice before the collective."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = ids.shape[1]
    max_local = (nq + world - 1) // world
    pad = max_local - ids.shape[0]

    def padded(t, fill):
        if pad == 0:
            return t.contiguous()
        extra = torch.full((pad,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
        return torch.cat([t, extra]).contiguous()

    g_ids = torch.empty((world * max_local, k), dtype=ids.dtype, device=ids.device)
    g_vals = torch.empty((world * max_local, k), dtype=vals.dtype, device=vals.device)
    g_cnt = torch.empty((world * max_local,), dtype=counts.dtype, device=counts.device)
    dist.all_gather_into_tensor(g_ids, padded(ids, -1), group=group)
    dist.all_gather_into_tensor(g_vals, padded(vals, float("nan")), group=group)
    dist.all_gather_into_tensor(g_cnt, padded(counts, 0), group=group)
    keep = []
    for r in range(world):
        lo, hi = shard_bounds(nq, world, r)
        keep.append(torch.arange(r * max_local, r * max_local + (hi - lo), device=ids.device))
    sel = torch.cat(keep)
    del rank
    return g_ids[sel], g_vals[sel], g_cnt[sel]


class ShardedSearcher:
    """Query-sharded search over replicated snapshots.

    ``local_search(queries[lo:hi], k, ef) -> (ids, dists, counts)`` is the per-rank device call
    (``DeviceSnapshot.search_batch_device`` wrapped by the caller); tests inject a stand-in."""

    def __init__(self, local_search, group=None):
        self.local_search = local_search
        self.group = group

    def search_batch(self, queries: torch.Tensor, k: int, ef: int):
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        nq = queries.shape[0]
        lo, hi = shard_bounds(nq, world, rank)
        ids, vals, cnt = self.local_search(queries[lo:hi], k, ef)
        return gather_topk(ids, vals, cnt, nq, self.group)


class PeerGather:
    """The library's own gather (csrc/comm.cu, include/veles_b200.h "multi-GPU"): the search kernel stores every
    query's top-k straight into each rank's gather window over NVLink, and a flag exchange closes the step.
    ``torch.distributed`` is used once, to hand the IPC handle blobs around."""

    def __init__(self, snapshot, rank, world, nq, k, device):
        import ctypes as C

        import numpy as np

        from . import _native as nv

        self.snap, self.rank, self.world, self.nq, self.k, self.device = snapshot, rank, world, nq, k, device
        lib = nv.lib()
        nb = lib.veles_comm_handle_bytes()
        blob = np.zeros(nb, np.uint8)
        h = C.c_void_p()
        nv.check(lib.veles_comm_create(rank, world, nq, k, C.byref(h), nv.ptr(blob)))
        self.h = h
        mine = torch.from_numpy(blob).to(device)
        every = torch.empty(world * nb, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(every, mine)
        self._blobs = every.cpu().numpy()
        # every rank must know whether EVERY rank mapped its peers, or a later collective would hang on the ones that
        # did not: agree on the outcome before anyone raises
        err = None
        try:
            nv.check(lib.veles_comm_connect(self.h, nv.ptr(self._blobs)))
        except Exception as e:  # noqa: BLE001
            err = e
        ok = torch.tensor([0 if err else 1], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"peer windows could not be mapped on every rank ({err or 'another rank failed'})")

    def search_gather(self, q_t, ef, stream=None, mid_event=None):
        from . import _native as nv

        ev = None if mid_event is None else mid_event.cuda_event
        if mid_event is not None and not ev:  # a torch event is created lazily by its first record()
            mid_event.record()
            ev = mid_event.cuda_event
        nv.check(nv.lib().veles_search_batch_gather_d(self.snap.h, self.h, nv.ptr(q_t), q_t.shape[0], self.k, ef, stream, ev))

    def window(self):
        """(ids [world*nq, k] int32, dist [world*nq, k] f32, counts [world*nq] int32) as torch views of the window."""
        import ctypes as C

        from . import _native as nv

        pi, pd, pc = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nv.check(nv.lib().veles_comm_window(self.h, C.byref(pi), C.byref(pd), C.byref(pc)))
        n = self.world * self.nq

        return (_device_view(pi.value, (n, self.k), torch.int32, self.device),
                _device_view(pd.value, (n, self.k), torch.float32, self.device),
                _device_view(pc.value, (n,), torch.int32, self.device))

    def status(self, stream=None):
        from . import _native as nv

        nv.check(nv.lib().veles_comm_status(self.h, stream))

    def check(self, ids_t, dist_t, rank):
        """After a step: the local window must hold every rank's results, in rank order.  `ids_t` / `dist_t` are this
        rank's results of the same queries from a plain veles_search_batch_d call; they are all-gathered with NCCL
        (the check, not the product path) and compared bit for bit with the window the library filled."""
        self.status(None)
        ids, dd, cnt = self.window()
        want_ids = torch.empty_like(ids)
        want_dd = torch.empty_like(dd)
        dist.all_gather_into_tensor(want_ids, ids_t.contiguous())
        dist.all_gather_into_tensor(want_dd, dist_t.contiguous())
        torch.cuda.synchronize()
        ok = torch.equal(ids, want_ids) and torch.equal(dd.view(torch.int32), want_dd.view(torch.int32))
        return bool(ok) and bool((cnt >= 0).all().item()) and bool((cnt <= self.k).all().item())

    def close(self):
        from . import _native as nv

        if self.h:
            nv.lib().veles_comm_destroy(self.h)
            self.h = None


class _CudaArray:
    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def _device_view(ptr, shape, dtype, device):
    typestr = {torch.int32: "<i4", torch.float32: "<f4"}[dtype]
    return torch.as_tensor(_CudaArray(ptr, shape, typestr), device=device)
