"""Multi-GPU parity: row-sharded search + ONE all-gather + merge kernel == oracle over the whole corpus."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import panoptikon_b200 as pk  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from panoptikon_b200 import sharding  # noqa: E402
from tests.helpers import assert_close_topk, assert_exact, int8_space  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# the library's own communicator: rank 0 makes the id, torch.distributed carries the 128 bytes
uid = torch.tensor(list(pk.Comm.unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
comm = pk.Comm(local, rank, world, bytes(uid.cpu().numpy().tolist()))
n, d, nq, k = 200_003, 256, 300, 100
x, q, scale, xc, qc = int8_space(n, d, 401, nq)
b, e = sharding.shard_range(n, world, rank)
for dtype, data, queries in ((pk.I8, xc, qc), (pk.F32, x, q)):
    ix = pk.VectorIndex(d, dtype, device=local)
    ix.set_row_base(b)
    ix.append(data[b:e])
    ix.seal()
    for metric in (pk.COSINE, pk.L2):
        ids, dd, _ = ix.search(torch.from_numpy(queries).cuda(), k, metric)
        m_ids, m_dist, m_cnt = sharding.gather_and_merge(ids, dd, lambda i, s: pk.merge_topk(i, s, device=local))
        f_ids, f_dist, f_cnt = sharding.gather_and_merge(ids, dd)   # pack kernel + one all-gather + merge of the packed buffer
        assert torch.equal(m_ids, f_ids) and torch.equal(m_cnt, f_cnt)
        assert torch.equal(m_dist.view(torch.int32), f_dist.view(torch.int32))
        # the same exchange entirely inside libpkv.so (pkv_search_sharded_device: scan + pack + ncclAllGather + merge)
        l_ids, l_dist, l_cnt = comm.search(ix, torch.from_numpy(queries).cuda(), k, metric)
        assert torch.equal(m_ids, l_ids) and torch.equal(m_cnt, l_cnt)
        assert torch.equal(m_dist.view(torch.int32), l_dist.view(torch.int32))
        if rank == 0:
            got = (m_ids.cpu().numpy(), m_dist.cpu().numpy(), m_cnt.cpu().numpy())
            want = orc.topk(data, queries, metric, k, threads=16)
            if dtype == pk.I8:
                assert_exact(got, want)
            else:
                assert_close_topk(got, want, data, queries, metric)
            print(f"world {world} dtype {dtype} metric {metric}: sharded search == oracle", flush=True)
    ix.close()
dist.barrier()
comm.close()
dist.destroy_process_group()
