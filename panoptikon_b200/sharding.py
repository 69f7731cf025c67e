"""Row-sharded search across the GPUs of one box: contiguous row ranges per rank,
ONE all-gather of the per-shard top-k lists, then a k-way merge under the same total order.

The collective plumbing is backend-agnostic torch.distributed (NCCL over NVLink on the GPUs; the
CPU tests drive it over gloo with a NumPy merge standing in for the CUDA merge kernel)."""
from __future__ import annotations

from typing import Callable, Tuple


def shard_range(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [begin, end) owned by `rank`: ceil(N/G) contiguous rows per rank (insertion order)."""
    per = (n_rows + world - 1) // world
    begin = min(rank * per, n_rows)
    return begin, min(begin + per, n_rows)


def pack_results(ids, dist):
    """[nq,k] int64 ids + [nq,k] f32 distances -> one [nq,k,3] int32 buffer (one collective)."""
    import torch

    nq, k = ids.shape
    packed = torch.empty((nq, k, 3), dtype=torch.int32, device=ids.device)
    packed[..., :2] = ids.contiguous().view(torch.int32).view(nq, k, 2)
    packed[..., 2] = dist.contiguous().view(torch.int32)
    return packed


def unpack_results(packed):
    """[parts,nq,k,3] int32 -> ([parts,nq,k] int64, [parts,nq,k] f32)."""
    import torch

    parts, nq, k, _ = packed.shape
    ids = packed[..., :2].contiguous().view(torch.int64).view(parts, nq, k)
    dist = packed[..., 2].contiguous().view(torch.float32)
    return ids, dist


def gather_and_merge(ids, dist, merge: Callable = None, group=None):
    """All ranks contribute their shard's (ids, dist); every rank gets the merged global top-k.
    `merge(ids[parts,nq,k], dist[parts,nq,k]) -> (ids, dist, counts)`; on GPUs pass
    panoptikon_b200.merge_topk."""
    import torch
    import torch.distributed as dist_mod

    world = dist_mod.get_world_size(group)
    if ids.is_cuda and merge is None:
        # GPU fast path: one pack kernel, ONE all-gather, one merge kernel reading the gathered buffer directly
        import panoptikon_b200 as pk

        dev = ids.device.index or 0
        packed = pk.pack_topk(ids, dist, device=dev)
        nq, k, _ = packed.shape
        gathered = torch.empty((world, nq, k, 3), dtype=torch.int32, device=packed.device)
        dist_mod.all_gather_into_tensor(gathered, packed, group=group)
        return pk.merge_packed(gathered, device=dev)
    packed = pack_results(ids, dist)
    nq, k, _ = packed.shape
    gathered = torch.empty((world * nq, k, 3), dtype=torch.int32, device=packed.device)  # rank-major concat
    dist_mod.all_gather_into_tensor(gathered, packed, group=group)
    g_ids, g_dist = unpack_results(gathered.view(world, nq, k, 3))
    return merge(g_ids, g_dist)
