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
