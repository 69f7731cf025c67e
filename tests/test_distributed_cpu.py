"""world_size-2 gloo test of the row-shard plumbing (shard ranges, the single packed all-gather,
merge order).  The per-shard search and the merge are the ORACLE here (no GPU in this container);
the CUDA merge kernel itself is checked against the same oracle in tests/test_gpu_operator.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from panoptikon_b200 import sharding


def _np_merge(ids, dist_):
    """(distance asc, NaN last, shard, position) merge of sorted per-shard lists."""
    ids, dist_ = ids.numpy(), dist_.numpy()
    parts, nq, k = ids.shape
    o_ids = np.full((nq, k), -1, np.int64)
    o_dist = np.full((nq, k), np.nan, np.float32)
    cnt = np.zeros(nq, np.int32)
    for q in range(nq):
        ent = [(np.isnan(dist_[p, q, i]), dist_[p, q, i] if not np.isnan(dist_[p, q, i]) else 0.0, p, i)
               for p in range(parts) for i in range(k) if ids[p, q, i] != -1]
        ent.sort()
        ent = ent[:k]
        cnt[q] = len(ent)
        for j, (_, _, p, i) in enumerate(ent):
            o_ids[q, j], o_dist[q, j] = ids[p, q, i], dist_[p, q, i]
    return torch.from_numpy(o_ids), torch.from_numpy(o_dist), torch.from_numpy(cnt)


def _worker(rank, world, port, n, d, nq, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = orc.synthetic(n, d, 301)
    q = orc.synthetic(nq, d, 302)
    x[5] = x[n - 3]          # a tie across shards: lower global row must win
    b, e = sharding.shard_range(n, world, rank)
    rows, dd, _ = orc.topk(x[b:e], q, orc.COSINE, k)
    ids = np.where(rows >= 0, rows + b, -1)
    m_ids, m_dist, m_cnt = sharding.gather_and_merge(torch.from_numpy(ids), torch.from_numpy(dd), _np_merge)
    if rank == 0:
        want = orc.topk(x, q, orc.COSINE, k)
        out["ok"] = bool(np.array_equal(m_ids.numpy(), want[0]) and
                         np.array_equal(m_dist.numpy().view(np.uint32), want[1].view(np.uint32)) and
                         np.array_equal(m_cnt.numpy(), want[2]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,k", [(1000, 20), (7, 5)])
def test_two_rank_shard_gather_merge(n, k):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, 64, 3, k, out), nprocs=2, join=True)
    assert out.get("ok") is True


def test_shard_ranges_cover_rows_contiguously():
    for n in (0, 1, 7, 10_000_000):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                b, e = sharding.shard_range(n, world, r)
                assert b == prev and e >= b
                prev = e
            assert prev == n


def test_pack_unpack_roundtrip():
    ids = torch.tensor([[1, -1, 2**40 + 3]], dtype=torch.int64)
    d = torch.tensor([[0.5, float("nan"), -0.0]], dtype=torch.float32)
    p = sharding.pack_results(ids, d)
    i2, d2 = sharding.unpack_results(p[None])
    assert torch.equal(i2[0], ids) and torch.equal(d2[0].view(torch.int32), d.view(torch.int32))
