"""Floating-point tensor-core paths (reduced-precision tcgen05 filter + exact FFMA re-scoring): scores within
1e-5 relative of the oracle, and the same answer as the CUDA-core kernel (same summation order in the
re-scoring).  Filter operands (`use_shadow`): 2 = int8 image (per-row scales, kind::i8, queries in TMEM),
1 = fp16 image (kind::f16), 0 = the f32 rows themselves (kind::tf32)."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk

pytestmark = pytest.mark.gpu
METRICS = [pk.L2, pk.COSINE, pk.DOT]


KIND = {2: 8, 1: 7, 0: 4}   # counters().last_scan_kind per filter operand


def _index(x):
    ix = pk.VectorIndex(x.shape[1], pk.F32)
    ix.set_option("image_mask", 3)   # build both images so that every filter can be selected
    ix.append(x)
    ix.seal()
    return ix


@pytest.mark.parametrize("shadow", [2, 1, 0])
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq", [17, 128, 300])
def test_tc_f32_matches_oracle_and_simt(metric, nq, shadow):
    x, q = orc.synthetic(60001, 768, 201), orc.synthetic(nq, 768, 202)
    with _index(x) as ix:
        ix.set_option("use_shadow", shadow)
        got = ix.search(q, 100, metric)
        assert ix.counters().last_scan_kind == KIND[shadow], "tensor-core kernel did not run"
        ix.set_option("force_simt", 1)
        simt = ix.search(q, 100, metric)
        assert ix.counters().last_scan_kind == 1
    want = orc.topk(x, q, metric, 100, threads=16)
    assert_close_topk(got, want, x, q, metric)
    assert_close_topk(simt, want, x, q, metric)
    assert np.array_equal(got[0], simt[0]), "tensor-core and CUDA-core paths returned different ids"
    assert np.array_equal(got[1].view(np.uint32), simt[1].view(np.uint32)), "paths returned different scores"


@pytest.mark.parametrize("dim", [8, 100, 512, 1024, 1536])
def test_tc_f32_dims_unnormalised(dim):
    # unnormalised rows with a wide norm spread stress the per-row error bound of the filter
    rng = np.random.default_rng(11)
    x = orc.synthetic(5003, dim, 211, normalise=False) * rng.uniform(0.01, 30.0, size=(5003, 1)).astype(np.float32)
    x[17] = 0.0
    q = orc.synthetic(40, dim, 212, normalise=False)
    q[3] *= 100.0
    with _index(x) as ix:
        for shadow in (2, 1):
            ix.set_option("use_shadow", shadow)
            for metric in METRICS:
                got = ix.search(q, 33, metric)
                assert ix.counters().last_scan_kind in (4, 7, 8)
                assert_close_topk(got, orc.topk(x, q, metric, 33, threads=8), x, q, metric)


def test_tc_f32_near_duplicates_and_bitmap():
    # clusters of almost identical rows: many pairs sit inside the tf32 error band of the threshold
    rng = np.random.default_rng(12)
    centers = orc.synthetic(50, 256, 221)
    x = np.repeat(centers, 400, axis=0) + rng.standard_normal((20000, 256)).astype(np.float32) * 1e-4
    x = np.ascontiguousarray(x.astype(np.float32))
    q = np.ascontiguousarray(centers[:24] + rng.standard_normal((24, 256)).astype(np.float32) * 1e-3)
    words = (len(x) + 63) // 64
    bm = np.packbits(rng.random(words * 64) < 0.5, bitorder="little").view(np.uint64)
    with _index(x) as ix:
        for shadow in (1, 2):
            ix.set_option("use_shadow", shadow)
            for metric in (pk.COSINE, pk.L2):
                assert_close_topk(ix.search(q, 200, metric), orc.topk(x, q, metric, 200, threads=8), x, q, metric)
        got = ix.search(q, 50, pk.COSINE, bitmap=bm)
        want = orc.topk(x, q, orc.COSINE, 50, bitmap=bm, threads=8)
        assert np.array_equal(got[2], want[2])
        assert np.allclose(got[1], want[1], rtol=1e-5, atol=1e-5)
        ix.set_option("candidate_capacity", 512)
        assert_close_topk(ix.search(q, 100, pk.L2), orc.topk(x, q, orc.L2, 100, threads=8), x, q, orc.L2)


def test_tc_shadow_tiny_and_huge_rows():
    # rows spanning 12 orders of magnitude in norm: tiny rows underflow in the fp16 image and must
    # still be found (cosine is scale invariant), huge rows must not overflow it
    rng = np.random.default_rng(13)
    x = orc.synthetic(6000, 128, 231)
    scale = np.power(10.0, rng.uniform(-9, 3, size=(6000, 1))).astype(np.float32)
    x = np.ascontiguousarray(x * scale)
    q = orc.synthetic(32, 128, 232)
    q[:8] = x[:8] / np.linalg.norm(x[:8], axis=1, keepdims=True)   # queries parallel to rows of every scale
    with _index(x) as ix:
        for shadow in (1, 2):
            ix.set_option("use_shadow", shadow)
            for metric in (pk.COSINE, pk.DOT, pk.L2):
                got = ix.search(q, 25, metric)
                assert ix.counters().last_scan_kind == KIND[shadow]
                assert_close_topk(got, orc.topk(x, q, metric, 25, threads=8), x, q, metric)
        # a query parallel to row i finds that row first (cosine distance ~0), whatever the row's scale; another
        # row may only take its place inside the f32 rounding of a zero distance
        for i in range(8):
            ids, dist, _ = ix.search(q[i:i + 1], 1, pk.COSINE)
            assert ids[0][0] == i or abs(float(dist[0][0])) <= 2e-6, (i, ids[0][0], dist[0][0])


def test_tc_f16_index():
    x = orc.synthetic(30000, 512, 241).astype(np.float16)
    q = orc.synthetic(70, 512, 242).astype(np.float16)
    ix = pk.VectorIndex(512, pk.F16)
    ix.append(x)
    ix.seal()
    with ix:
        for metric in METRICS:
            img = ix.search(q, 100, metric)           # int8 image of the f16 rows
            assert ix.counters().last_scan_kind == 8
            assert_close_topk(img, orc.topk(x, q, metric, 100, threads=8), x, q, metric)
            ix.set_option("use_shadow", 0)            # kind::f16 on the rows themselves
            got = ix.search(q, 100, metric)
            ix.set_option("use_shadow", -1)
            assert ix.counters().last_scan_kind == 6
            assert np.array_equal(got[0], img[0]) and np.array_equal(got[1].view(np.uint32), img[1].view(np.uint32))
            assert_close_topk(got, orc.topk(x, q, metric, 100, threads=8), x, q, metric)
            ix.set_option("force_simt", 1)
            simt = ix.search(q, 100, metric)
            ix.set_option("force_simt", 0)
            assert np.array_equal(got[0], simt[0]) and np.array_equal(got[1].view(np.uint32), simt[1].view(np.uint32))


def test_shadow_survives_incremental_appends_with_growing_range():
    # a later batch with a 1000x larger component forces the image to be rebuilt at a new scale
    x1 = orc.synthetic(5000, 64, 251)
    x2 = orc.synthetic(3000, 64, 252) * np.float32(1000.0)
    q = orc.synthetic(20, 64, 253)
    ix = pk.VectorIndex(64, pk.F32)
    with ix:
        ix.append(x1); ix.seal()
        a = ix.search(q, 10, pk.COSINE)
        assert_close_topk(a, orc.topk(x1, q, orc.COSINE, 10), x1, q, orc.COSINE)
        ix.append(x2); ix.seal()
        x = np.concatenate([x1, x2])
        for metric in METRICS:
            assert_close_topk(ix.search(q, 10, metric), orc.topk(x, q, metric, 10, threads=4), x, q, metric)


@pytest.mark.parametrize("shadow", [2, 1, 0])
def test_tc_float_cta_pair_matches_single_cta(shadow):
    x, q = orc.synthetic(70007, 512, 261), orc.synthetic(256, 512, 262)
    with _index(x) as ix:
        ix.set_option("use_shadow", shadow)
        for metric in METRICS:
            pair = ix.search(q, 50, metric)
            ix.set_option("tc_cta2", 0)
            single = ix.search(q, 50, metric)
            ix.set_option("tc_cta2", 1)
            assert np.array_equal(pair[0], single[0]) and np.array_equal(pair[1].view(np.uint32), single[1].view(np.uint32))
            assert_close_topk(pair, orc.topk(x, q, metric, 50, threads=16), x, q, metric)


def test_img8_non_finite_rows_and_queries():
    # NaN / inf components cannot be quantised: such rows (and queries) bypass the filter and are re-scored
    x = orc.synthetic(9000, 256, 271)
    x[5, 3] = np.inf
    x[6, 7] = np.nan
    x[7] = 0.0
    x[8] *= np.float32(1e-30)
    x[9] *= np.float32(1e30)
    q = orc.synthetic(130, 256, 272)
    q[2] *= np.float32(1e-20)
    with _index(x) as ix:
        for shadow in (1, 2):
            ix.set_option("use_shadow", shadow)
            for metric in METRICS:
                got = ix.search(q, 20, metric)
                assert ix.counters().last_scan_kind == KIND[shadow]
                # the 1e30-scaled row has a dot product 1000x smaller than its terms: summation order alone moves
                # it by > 1e-5 relative, and its size puts it in every DOT top-k
                rtol = 1e-3 if metric == pk.DOT else 1e-5
                assert_close_topk(got, orc.topk(x, q, metric, 20, threads=8), x, q, metric, rtol=rtol)


@pytest.mark.parametrize("nq", [1, 64, 129, 512, 1024])
def test_img8_batches_and_k(nq):
    x, q = orc.synthetic(50021, 768, 281), orc.synthetic(nq, 768, 282)
    with _index(x) as ix:
        got = ix.search(q, 100, pk.COSINE)
        assert ix.counters().last_scan_kind == 8
        assert_close_topk(got, orc.topk(x, q, orc.COSINE, 100, threads=16), x, q, orc.COSINE)
        if nq <= 64:
            big = ix.search(q, 1000, pk.L2)   # deep k: the filter passes several rows per true candidate
            assert_close_topk(big, orc.topk(x, q, orc.L2, 1000, threads=16), x, q, orc.L2)


def test_img8_incremental_appends():
    x1 = orc.synthetic(5000, 64, 291)
    x2 = orc.synthetic(3000, 64, 292) * np.float32(1000.0)
    q = orc.synthetic(20, 64, 293)
    ix = pk.VectorIndex(64, pk.F32)
    with ix:
        ix.append(x1); ix.seal()
        assert_close_topk(ix.search(q, 10, pk.COSINE), orc.topk(x1, q, orc.COSINE, 10), x1, q, orc.COSINE)
        assert ix.counters().last_scan_kind == 8
        ix.append(x2); ix.seal()
        x = np.concatenate([x1, x2])
        for metric in METRICS:
            assert_close_topk(ix.search(q, 10, metric), orc.topk(x, q, metric, 10, threads=4), x, q, metric)


def test_img8_peaky_and_sparse_rows():
    # one-hot-like rows saturate the direction quantiser (their peakiness is far above the index's reference):
    # their error term grows instead of their codes overflowing, and they are still ranked exactly
    rng = np.random.default_rng(31)
    x = orc.synthetic(12000, 384, 301)
    hot = rng.integers(0, 384, size=3000)
    x[:3000] *= np.float32(0.02)
    x[np.arange(3000), hot] += rng.choice(np.float32([-1.0, 1.0]), size=3000)
    x[3000:3100] = 0.0
    x[3000:3100, 7] = np.float32(5.0)
    q = orc.synthetic(140, 384, 302)
    q[:20] = x[:20] * np.float32(3.0)          # queries parallel to peaky rows
    q[20:30] = 0.0
    q[20:30, 7] = 1.0                          # sparse queries
    ix = pk.VectorIndex(384, pk.F32)
    ix.append(x)
    ix.seal()
    with ix:
        for metric in METRICS:
            got = ix.search(q, 60, metric)
            assert ix.counters().last_scan_kind == 8
            assert_close_topk(got, orc.topk(x, q, metric, 60, threads=8), x, q, metric)


@pytest.mark.parametrize("buffers", [2, 3])
@pytest.mark.parametrize("nq", [20, 300])
def test_img8_accumulator_buffer_shapes(nq, buffers):
    x, q = orc.synthetic(30011, 768, 311), orc.synthetic(nq, 768, 312)
    with _index(x) as ix:
        ix.set_option("ts_acc_buffers", buffers)
        got = ix.search(q, 40, pk.COSINE)
        assert ix.counters().last_scan_kind == 8
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 40, threads=16), x, q, orc.COSINE)
