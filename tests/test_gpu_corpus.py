"""Corpus lifecycle: the reference's pair state machine around the HBM replica - chunk-idempotent
backfill keyed by (profile_id, artifact_rev) (db/vector_quants.rs:1085-1163), the inline hook and its wrong-dimension
downgrade (:1347-1438), readiness at (artifact_rev, epoch) (:1829-1850, db/epochs.rs:38-44)."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk, assert_exact, int8_space

pytestmark = pytest.mark.gpu


def test_int8_replica_backfill_is_chunk_idempotent_and_quantises_on_the_gpu():
    n, d = 12_000, 64
    x, q, scale, xc, qc = int8_space(n, d, seed=901, nq=6)
    ids = np.arange(n, dtype=np.int64) * 2 + 100                     # ascending item_data.id
    art = pk.scale_artifact(scale)
    with pk.Corpus(d, pk.I8, "index.db", "clip/test", profile_id=4) as c:
        assert c.ready(1, 10) is None and c.info().state == pk.Corpus.PENDING
        with pytest.raises(pk.PkvError, match="Invalid vector quant scale artifact"):
            c.begin(1, b"\x00\x00\x00\x00", 10)
        c.begin(1, art, 10)
        cursor = None
        for b in range(0, n, 5000):                                 # the reference's 5000-row chunks, f32 source blobs
            written, cursor = c.upload_chunk(1, ids[b:b + 5000], x[b:b + 5000])
            assert written == len(ids[b:b + 5000]) and cursor == ids[min(b + 5000, n) - 1]
        assert c.upload_chunk(1, ids[5000:10000], x[5000:10000]) == (0, cursor)      # a replayed chunk writes nothing
        assert c.upload_chunk(2, ids[:10], x[:10])[0] == 0                           # another revision: refused, zero rows
        assert c.ready(1, 10) is None                                                # still building
        c.finish(1, 10)
        view, pair = c.ready(1, 10)
        assert pair == (4, scale, d) and c.info().rows == n
        got = view.search(qc, 50, pk.COSINE)
        want = orc.topk(xc, qc, orc.COSINE, 50, threads=4)                           # GPU codec == quantize_int8
        assert_exact(got, (ids[want[0]], want[1], want[2]))
        assert c.ready(2, 10) is None and c.ready(1, 11) is None                     # other revision / the DB moved on
        # the inline hook: a new embedding is searchable at once and moves the synced epoch
        new = orc.synthetic(1, d, 902)
        c.append_inline(int(ids[-1]) + 7, new[0], 11)
        c.append_inline(int(ids[-1]) + 7, new[0], 11)                                # ON CONFLICT: no duplicate
        view, _ = c.ready(1, 11)
        assert c.info().rows == n + 1
        hit = view.search(orc.quantize_rows(new, scale), 1, pk.COSINE)
        assert hit[0][0][0] == ids[-1] + 7
        # a vector of the wrong dimensionality downgrades the pair: searches fall back to exact until rebuilt
        with pytest.raises(pk.PkvError, match="downgraded to pending"):
            c.append_inline(int(ids[-1]) + 9, np.zeros(d + 1, np.float32), 12)
        assert c.ready(1, 12) is None and c.info().state == pk.Corpus.PENDING
        # rebuild under a new scale at the next revision: the old rows are gone
        c.begin(2, pk.scale_artifact(scale * 2), 13)
        assert c.info().rows == 0 and c.info().state == pk.Corpus.BUILDING


def test_exact_replica_resumes_a_build_and_is_invalidated_by_an_epoch_bump():
    n, d = 9_000, 32
    x, q = orc.synthetic(n, d, 911), orc.synthetic(4, d, 912)
    ids = np.arange(n, dtype=np.int64) + 1
    with pk.Corpus(d, pk.F32, "index.db", "all-mpnet-base-v2") as c:
        c.begin(0, None, 3)
        c.upload_chunk(0, ids[:4000], x[:4000])
        c.begin(0, None, 3)                                          # the worker restarted: same revision resumes
        assert c.info().rows == 4000 and c.info().cursor == 4000
        c.upload_chunk(0, ids[3000:], x[3000:])                      # overlaps what is there: only the new rows land
        assert c.info().rows == n
        with pytest.raises(pk.PkvError):
            c.upload_chunk(0, ids[:3][::-1].copy(), x[:3])           # ids must ascend (ORDER BY d.id)
        c.finish(0, 3)
        view, pair = c.ready(0, 3)
        got = view.search(q, 20, pk.L2)
        want = orc.topk(x, q, orc.L2, 20, threads=4)
        assert np.array_equal(got[0], ids[want[0]]) and np.allclose(got[1], want[1], rtol=1e-5)
        c.invalidate()                                               # an unmirrored write bumped the epoch
        assert c.ready(0, 3) is None
        with pytest.raises(pk.PkvError, match="not building"):
            c.finish(0, 4)
