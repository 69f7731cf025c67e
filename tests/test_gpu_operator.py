"""GPU tests of the pieces around the scan: PQL policy (Space), shard merge, aggregation, device API."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk, assert_exact, int8_space

pytestmark = pytest.mark.gpu


def _space(n=5000, d=64, ready=True, default=True):
    x, q, scale, xc, qc = int8_space(n, d, 81, 3)
    fx = pk.VectorIndex(d, pk.F32); fx.append(x); fx.seal()
    qx = pk.VectorIndex(d, pk.I8); qx.set_scale_artifact(pk.scale_artifact(scale)); qx.append(xc); qx.seal()
    sp = pk.Space("clip/model", fx)
    if ready:
        sp.set_quant("plain", pk.ReadyPair(7, scale, d), qx, is_default=default)
    return sp, (x, q, scale, xc, qc), (fx, qx)


def test_index_mode_policy_end_to_end():
    sp, (x, q, scale, xc, qc), _keep = _space()
    want_exact = orc.topk(x, q, orc.COSINE, 20)
    want_quant = orc.topk(xc, qc, orc.COSINE, 20)
    ids, dist, cnt, used = sp.search(q, pk.COSINE, index=pk.INDEX_EXACT, depth=20)
    assert used == -1
    assert_close_topk((ids, dist, cnt), want_exact, x, q, orc.COSINE)
    for mode, variant in ((pk.INDEX_AUTO, None), (pk.INDEX_AUTO, "  "), (pk.INDEX_QUANT, None), (pk.INDEX_AUTO, "plain")):
        ids, dist, cnt, used = sp.search(q, pk.COSINE, index=mode, variant=variant, depth=20)
        assert used == 7
        assert_exact((ids, dist, cnt), want_quant)
    # k is validated but otherwise inert (docs/vector-int8-quant.md:86-88)
    a = sp.search(q, pk.COSINE, k=1, depth=20)
    b = sp.search(q, pk.COSINE, k=10000, depth=20)
    assert np.array_equal(a[0], b[0])
    with pytest.raises(pk.PqlError, match="k must be a positive integer"):
        sp.search(q, pk.COSINE, k=0)
    with pytest.raises(pk.PqlError, match="reserved"):
        sp.search(q, pk.COSINE, index=pk.INDEX_ANN)
    # dimension mismatch: auto falls back to exact (which then rejects), strict errors name the pair's dim
    with pytest.raises(pk.PqlError, match=r"expected 64, got 32"):
        sp.search(np.zeros((1, 32), np.float32), pk.COSINE, index=pk.INDEX_QUANT)


def test_auto_falls_back_and_quant_is_strict_when_not_ready():
    sp, (x, q, scale, xc, qc), _keep = _space(ready=False)
    ids, dist, cnt, used = sp.search(q, pk.L2, index=pk.INDEX_AUTO, depth=10)
    assert used == -1
    assert_close_topk((ids, dist, cnt), orc.topk(x, q, orc.L2, 10), x, q, orc.L2)
    with pytest.raises(pk.PqlError, match="no default vector quant profile is configured"):
        sp.search(q, pk.L2, index=pk.INDEX_QUANT)
    with pytest.raises(pk.PqlError, match="vector quant profile 'plain' does not exist or is not ready for model 'clip/model'"):
        sp.search(q, pk.L2, index=pk.INDEX_AUTO, variant="plain")


def test_shard_merge_equals_single_index():
    import torch

    x, q, scale, xc, qc = int8_space(40000, 128, 91, 6)
    parts, k = 4, 50
    bounds = np.linspace(0, len(xc), parts + 1).astype(int)
    ids_l, dist_l = [], []
    for p in range(parts):
        with pk.VectorIndex(128, pk.I8) as ix:
            ix.set_row_base(int(bounds[p]))
            ix.append(xc[bounds[p]:bounds[p + 1]]); ix.seal()
            i, d, _ = ix.search(torch.from_numpy(qc).cuda(), k, pk.COSINE)
            ids_l.append(i); dist_l.append(d)
    ids, dist, cnt = pk.merge_topk(torch.stack(ids_l).contiguous(), torch.stack(dist_l).contiguous())
    assert_exact((ids.cpu().numpy(), dist.cpu().numpy(), cnt.cpu().numpy()), orc.topk(xc, qc, orc.COSINE, k, threads=4))
    # ragged: one shard smaller than k, one empty
    ids_l, dist_l = [], []
    cuts = [0, 10, 10, len(xc)]
    for p in range(3):
        with pk.VectorIndex(128, pk.I8) as ix:
            ix.set_row_base(cuts[p])
            ix.append(xc[cuts[p]:cuts[p + 1]]); ix.seal()
            i, d, _ = ix.search(torch.from_numpy(qc).cuda(), k, pk.L2)
            ids_l.append(i); dist_l.append(d)
    ids, dist, cnt = pk.merge_topk(torch.stack(ids_l).contiguous(), torch.stack(dist_l).contiguous())
    assert_exact((ids.cpu().numpy(), dist.cpu().numpy(), cnt.cpu().numpy()), orc.topk(xc, qc, orc.L2, k, threads=4))


def test_device_api_matches_host_api():
    import torch

    x, q = orc.synthetic(10000, 256, 95), orc.synthetic(4, 256, 96)
    with pk.VectorIndex(256, pk.F32) as ix:
        ix.append(torch.from_numpy(x).cuda()); ix.seal()
        h = ix.search(q, 30, pk.COSINE)
        d = ix.search(torch.from_numpy(q).cuda(), 30, pk.COSINE)
    assert np.array_equal(h[0], d[0].cpu().numpy()) and np.array_equal(h[1], d[1].cpu().numpy())


def test_aggregate_matches_oracle():
    import torch

    rng = np.random.default_rng(3)
    n, items = 20000, 700
    d = rng.random(n).astype(np.float32)
    d[::97] = np.nan
    item = rng.integers(0, items, n).astype(np.int64)
    w = (rng.random(n) + 0.1).astype(np.float32)
    td, ti, tw = torch.from_numpy(d).cuda(), torch.from_numpy(item).cuda(), torch.from_numpy(w).cuda()
    for agg in (pk.AGG_MIN, pk.AGG_MAX, pk.AGG_AVG):
        got = pk.aggregate(td, ti, items + 5, agg).cpu().numpy()
        want = orc.aggregate(d, item, items + 5, agg)
        assert np.allclose(got, want, rtol=1e-12, atol=0, equal_nan=True)
    got = pk.aggregate(td, ti, items + 5, pk.AGG_AVG, weights=tw).cpu().numpy()
    assert np.allclose(got, orc.aggregate(d, item, items + 5, orc.AGG_AVG, weights=w), rtol=1e-12, equal_nan=True)


@pytest.mark.parametrize("dtype", ["f32", "i8"])
def test_untruncated_scoring_and_item_aggregation(dtype):
    """The reference scores EVERY candidate row and aggregates per file (MIN/MAX/AVG,
    builder/filters/exact.rs:67-80) before any LIMIT: pkv_distances_device + pkv_aggregate_device."""
    import torch

    n, d, items = 30011, 128, 4000
    x, q, scale, xc, qc = int8_space(n, d, 97, 3)
    data, queries, code = (x, q, pk.F32) if dtype == "f32" else (xc, qc, pk.I8)
    rng = np.random.default_rng(5)
    item = rng.integers(0, items, n).astype(np.int64)      # several embeddings per item (video frames, text chunks)
    with pk.VectorIndex(d, code) as ix:
        ix.append(data); ix.seal()
        for metric in (pk.COSINE, pk.L2):
            dist = ix.distances(torch.from_numpy(queries).cuda(), metric)
            assert dist.shape == (3, n)
            for qi in range(3):
                want = orc.distances(data, queries[qi], metric)
                got = dist[qi].cpu().numpy()
                if dtype == "i8":
                    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
                else:
                    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
                for agg in (pk.AGG_MIN, pk.AGG_AVG, pk.AGG_MAX):
                    g = pk.aggregate(dist[qi].contiguous(), torch.from_numpy(item).cuda(), items, agg).cpu().numpy()
                    w = orc.aggregate(want, item, items, agg)
                    assert np.allclose(g, w, rtol=1e-5 if dtype == "f32" else 1e-12, atol=1e-7, equal_nan=True)
