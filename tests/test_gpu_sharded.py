"""Row sharding inside the library (pkv_sharded_* / pkv_comm_*, include/pkv.h "sharding"): the C ABI alone produces the
global top-k from several shards, with the order and the values of a single index holding all rows."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk, assert_exact, int8_space

pytestmark = pytest.mark.gpu


def _same(a, b):
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
            and np.array_equal(a[2], b[2]))


@pytest.mark.parametrize("shards", [2, 3, 5, 8])
def test_single_process_shards_equal_one_index_f32(shards):
    n, d = 70_001, 256
    x, q = orc.synthetic(n, d, 401), orc.synthetic(70, d, 402)
    with pk.ShardedIndex(d, pk.F32, devices=[0] * shards, total_rows=n) as sx:
        for b in range(0, n, 9_999):      # appends straddle the shard boundaries
            sx.append(x[b:b + 9_999])
        sx.seal()
        assert sx.rows == n and sum(sx.shard_rows()) == n and len(sx.shard_rows()) == shards
        got = sx.search(q, 100, pk.COSINE)
        got_l2 = sx.search(q, 7, pk.L2)
    with pk.VectorIndex(d, pk.F32) as ix:
        ix.append(x)
        ix.seal()
        one = ix.search(q, 100, pk.COSINE)
        one_l2 = ix.search(q, 7, pk.L2)
    assert _same(got, one) and _same(got_l2, one_l2)
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 100, threads=16), x, q, orc.COSINE)


def test_single_process_shards_int8_bitmap_and_row_ids():
    n, d = 50_000, 128
    x, q, scale, xc, qc = int8_space(n, d, seed=411, nq=33)
    ids = (np.arange(n, dtype=np.int64) * 7 + 1000)          # item_data.id values
    rng = np.random.default_rng(41)
    bm = np.packbits(rng.random(((n + 63) // 64) * 64) < 0.3, bitorder="little").view(np.uint64)
    with pk.ShardedIndex(d, pk.I8, devices=[0, 0, 0], total_rows=n) as sx:
        sx.set_scale_artifact(pk.scale_artifact(scale))
        sx.append(xc, ids)
        sx.seal()
        got = sx.search(qc, 50, pk.COSINE)
        got_bm = sx.search(qc, 50, pk.L2, bitmap=bm)
        with pytest.raises(pk.PkvError):
            sx.search(qc[:, :64], 5, pk.COSINE)
    want = orc.topk(xc, qc, orc.COSINE, 50, threads=8)
    assert_exact((got[0], got[1], got[2]), (ids[want[0]], want[1], want[2]))
    want_bm = orc.topk(xc, qc, orc.L2, 50, bitmap=bm, threads=8)
    rows = np.where(want_bm[0] >= 0, ids[np.maximum(want_bm[0], 0)], -1)
    assert_exact(got_bm, (rows, want_bm[1], want_bm[2]))


def test_fewer_rows_than_shards_and_empty_tail_shards():
    x, q = orc.synthetic(100, 64, 421), orc.synthetic(3, 64, 422)
    with pk.ShardedIndex(64, pk.F32, devices=[0, 0, 0, 0], total_rows=100) as sx:   # 64 rows per shard: two stay empty
        sx.append(x)
        sx.seal()
        assert sx.shard_rows() == [64, 36, 0, 0]
        got = sx.search(q, 10, pk.COSINE)
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 10, threads=3), x, q, orc.COSINE)


def test_nccl_communicator_single_rank_round_trip():
    # one rank on the one visible GPU: the full pack -> ncclAllGather -> merge path of pkv_search_sharded_device
    import torch

    x, q = orc.synthetic(30_000, 128, 431), orc.synthetic(40, 128, 432)
    comm = pk.Comm(0, 0, 1, pk.Comm.unique_id())
    try:
        with pk.VectorIndex(128, pk.F32) as ix:
            ix.append(x)
            ix.seal()
            qd = torch.from_numpy(q).cuda()
            ids, dist, cnt = comm.search(ix, qd, 20, pk.COSINE)
            got = (ids.cpu().numpy(), dist.cpu().numpy(), cnt.cpu().numpy())
            one = ix.search(q, 20, pk.COSINE)
    finally:
        comm.close()
    assert _same(got, one)


def test_single_process_shards_over_every_visible_gpu():
    """The 8-shard layout of BASELINE configs 4/5 with one shard per device when the box has several (the driver's
    single-GPU box runs it as 8 shards on device 0: same library path, same merge order)."""
    import torch

    ndev = torch.cuda.device_count()
    devices = [i % ndev for i in range(8)]
    n, d = 400_000, 128
    x, q = orc.synthetic(n, d, 441), orc.synthetic(64, d, 442)
    rng = np.random.default_rng(44)
    bm = np.packbits(rng.random(((n + 63) // 64) * 64) < 0.1, bitorder="little").view(np.uint64)
    with pk.ShardedIndex(d, pk.F32, devices=devices, total_rows=n) as sx:
        sx.append(x)
        sx.seal()
        assert len(sx.shard_rows()) == 8 and sum(sx.shard_rows()) == n
        got = sx.search(q, 100, pk.COSINE)
        got_bm = sx.search(q, 100, pk.COSINE, bitmap=bm)      # config 5: one global tag bitmap, sliced per shard
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 100, threads=16), x, q, orc.COSINE)
    assert_close_topk(got_bm, orc.topk(x, q, orc.COSINE, 100, bitmap=bm, threads=16), x, q, orc.COSINE)
