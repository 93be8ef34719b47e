"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: query sharding + final top-k gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from velesdb_b200.dist import ShardedSearcher, gather_topk, shard_bounds


def test_shard_bounds_cover_everything():
    for nq in (0, 1, 2, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(nq, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nq
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nq, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = torch.arange(nq * 4, dtype=torch.float32).reshape(nq, 4)

    def fake_local(qs, k, ef):  # deterministic stand-in for the device search
        base = (qs[:, 0] / 4).to(torch.int32)
        ids = base[:, None] * 100 + torch.arange(k, dtype=torch.int32)[None, :]
        vals = ids.to(torch.float32) * 0.5
        cnt = torch.full((qs.shape[0],), k, dtype=torch.int32)
        return ids, vals, cnt

    ids, vals, cnt = ShardedSearcher(fake_local).search_batch(q, k, 64)
    if rank == 0:
        torch.save((ids, vals, cnt), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [8, 9, 1])
def test_sharded_search_gathers_in_query_order(tmp_path, nq):
    out = str(tmp_path / "r.pt")
    k = 3
    mp.spawn(_worker, args=(2, _free_port(), nq, k, out), nprocs=2, join=True)
    ids, vals, cnt = torch.load(out)
    exp = np.arange(nq)[:, None] * 100 + np.arange(k)[None, :]
    assert ids.shape == (nq, k) and np.array_equal(ids.numpy(), exp)
    assert np.array_equal(vals.numpy(), exp * 0.5) and (cnt.numpy() == k).all()
