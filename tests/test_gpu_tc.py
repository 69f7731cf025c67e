"""Tensor-core (tcgen05 kind::i8) scan against the oracle: bit-exact ids and distances, and
identical to the CUDA-core kernel on the same index."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_exact, int8_space

pytestmark = pytest.mark.gpu
METRICS = [pk.L2, pk.COSINE, pk.DOT]


def _index(xc, scale=None):
    ix = pk.VectorIndex(xc.shape[1], pk.I8)
    if scale is not None:
        ix.set_scale_artifact(pk.scale_artifact(scale))
    ix.append(xc)
    ix.seal()
    return ix


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq", [1, 9, 128, 130, 300, 513])
def test_tc_int8_bit_exact(metric, nq):
    x, q, scale, xc, qc = int8_space(70001, 768, 101, nq)
    with _index(xc, scale) as ix:
        got = ix.search(qc, 100, metric)
        assert ix.counters().last_scan_kind == 3, "tensor-core kernel did not run"
        ix.set_option("force_simt", 1)
        simt = ix.search(qc, 100, metric)
        assert ix.counters().last_scan_kind == 2, "CUDA-core kernel did not run"
    want = orc.topk(xc, qc, metric, 100, threads=16)
    assert_exact(got, want)
    assert_exact(simt, want)


@pytest.mark.parametrize("dim", [8, 128, 520, 1024])
def test_tc_dims_saturation_zero_rows(dim):
    rng = np.random.default_rng(7)
    xc = rng.integers(-128, 128, size=(3000, dim), dtype=np.int8)
    xc[7] = -128
    xc[8] = 127
    xc[9] = 0
    xc[2999] = 0
    qc = rng.integers(-128, 128, size=(20, dim), dtype=np.int8)
    qc[0] = -128
    qc[1] = 0   # zero query: every cosine is NaN
    with _index(xc) as ix:
        for metric in METRICS:
            k = 3000 if dim == 8 else 77
            got = ix.search(qc, k, metric)
            assert ix.counters().last_scan_kind == 3
            assert_exact(got, orc.topk(xc, qc, metric, k, threads=8))


def test_tc_bitmap_duplicates_and_overflow():
    base = np.random.default_rng(8).integers(-100, 100, size=(7, 256), dtype=np.int8)
    xc = np.ascontiguousarray(np.tile(base, (6000, 1)))
    qc = np.ascontiguousarray(np.tile(base, (3, 1))[:16])
    rng = np.random.default_rng(9)
    words = (len(xc) + 63) // 64
    shared = np.packbits(rng.random(words * 64) < 0.3, bitorder="little").view(np.uint64)
    with _index(xc) as ix:
        assert_exact(ix.search(qc, 500, pk.COSINE), orc.topk(xc, qc, orc.COSINE, 500, threads=8))
        assert_exact(ix.search(qc, 200, pk.L2, bitmap=shared), orc.topk(xc, qc, orc.L2, 200, bitmap=shared, threads=8))
        ix.set_option("candidate_capacity", 256)   # force range splitting
        assert_exact(ix.search(qc, 100, pk.COSINE), orc.topk(xc, qc, orc.COSINE, 100, threads=8))
        assert ix.counters().last_scan_kind == 3


def test_tc_adversarial_order():
    x, q, scale, xc, qc = int8_space(50000, 128, 111, 12)
    order = np.argsort(xc.astype(np.int32) @ qc[0].astype(np.int32))   # ascending similarity to query 0
    xc = np.ascontiguousarray(xc[order])
    with _index(xc, scale) as ix:
        ix.set_option("candidate_capacity", 512)
        got = ix.search(qc, 50, pk.COSINE)
        assert ix.counters().fallback_queries > 0
    assert_exact(got, orc.topk(xc, qc, orc.COSINE, 50, threads=8))


def test_tc_cta_pair_matches_single_cta():
    # 256 queries per pass on a CTA pair (cta_group::2) vs two 128-query passes on single CTAs
    x, q, scale, xc, qc = int8_space(90011, 512, 121, 256)
    with _index(xc, scale) as ix:
        for metric in METRICS:
            pair = ix.search(qc, 64, metric)
            ix.set_option("tc_cta2", 0)
            single = ix.search(qc, 64, metric)
            ix.set_option("tc_cta2", 1)
            assert_exact(pair, single)
            assert_exact(pair, orc.topk(xc, qc, metric, 64, threads=16))


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq,groups", [(129, 1), (256, 2), (300, 2), (513, 1), (700, 4), (1024, 2), (1030, 4)])
def test_ts_queries_in_tmem_bit_exact(metric, nq, groups):
    # queries resident in TMEM (tcgen05.mma A operand from tensor memory), `groups` x 256 queries per launch
    x, q, scale, xc, qc = int8_space(60013, 768, 131, nq)
    with _index(xc, scale) as ix:
        ix.set_option("ts_groups", groups)
        ts = ix.search(qc, 100, metric)
        assert ix.counters().last_scan_kind == 3
        ix.set_option("tc_ts", 0)
        ss = ix.search(qc, 100, metric)
    assert_exact(ts, ss)
    assert_exact(ts, orc.topk(xc, qc, metric, 100, threads=16))


@pytest.mark.parametrize("dim", [8, 100, 512, 1024])
def test_ts_dims_saturation_zero_rows(dim):
    rng = np.random.default_rng(17)
    xc = rng.integers(-128, 128, size=(5000, dim), dtype=np.int8)
    xc[7] = -128
    xc[8] = 127
    xc[9] = 0
    xc[4999] = 0
    qc = rng.integers(-128, 128, size=(260, dim), dtype=np.int8)
    qc[0] = -128
    qc[1] = 0   # zero query: every cosine is NaN
    qc[259] = 127
    with _index(xc) as ix:
        for metric in METRICS:
            got = ix.search(qc, 33, metric)
            assert ix.counters().last_scan_kind == 3
            assert_exact(got, orc.topk(xc, qc, metric, 33, threads=8))


def test_ts_bitmap_and_small_buffers():
    x, q, scale, xc, qc = int8_space(40000, 256, 141, 200)
    rng = np.random.default_rng(19)
    words = (len(xc) + 63) // 64
    shared = np.packbits(rng.random(words * 64) < 0.2, bitorder="little").view(np.uint64)
    with _index(xc, scale) as ix:
        assert_exact(ix.search(qc, 50, pk.L2, bitmap=shared), orc.topk(xc, qc, orc.L2, 50, bitmap=shared, threads=8))
        ix.set_option("candidate_capacity", 256)   # force range splitting at odd row offsets
        assert_exact(ix.search(qc, 100, pk.COSINE), orc.topk(xc, qc, orc.COSINE, 100, threads=8))


@pytest.mark.parametrize("dim", [128, 640, 768, 1000])
@pytest.mark.parametrize("chunks", [1, 2, 3, 4])
def test_ts_chunks_per_stage(dim, chunks):
    # a stage of the row pipeline holds `chunks` 8 KB K-chunks; rows whose chunk count is not a multiple end on a short stage
    x, q, scale, xc, qc = int8_space(30011, dim, 151, 257)
    with _index(xc, scale) as ix:
        ix.set_option("ts_chunks", chunks)
        got = ix.search(qc, 40, pk.COSINE)
        assert ix.counters().last_scan_kind == 3
    assert_exact(got, orc.topk(xc, qc, orc.COSINE, 40, threads=16))


@pytest.mark.parametrize("buffers", [2, 3])
@pytest.mark.parametrize("dim", [512, 768, 896, 1024])
def test_ts_accumulator_buffer_shapes(dim, buffers):
    # 3 buffers: 128-row tiles up to D=512, 96-row tiles up to D=896, else 2 x 128 rows; all must agree with the oracle
    x, q, scale, xc, qc = int8_space(20017, dim, 161, 300)
    with _index(xc, scale) as ix:
        ix.set_option("ts_acc_buffers", buffers)
        got = ix.search(qc, 50, pk.L2)
        assert ix.counters().last_scan_kind == 3
    assert_exact(got, orc.topk(xc, qc, orc.L2, 50, threads=16))
