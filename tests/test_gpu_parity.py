"""GPU parity: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.
int8: ids and distances bit-exact.  f32/f16: distances within 1e-5 relative (north_star)."""
import json
import os

import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk, assert_exact, int8_space

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))
METRICS = [pk.L2, pk.COSINE, pk.DOT]


def build(x, dtype, ids=None, scale=None, chunks=1):
    ix = pk.VectorIndex(x.shape[1], dtype)
    if scale is not None:
        ix.set_scale_artifact(pk.scale_artifact(scale))
    n = x.shape[0]
    step = max(1, (n + chunks - 1) // chunks)
    for b in range(0, n, step):
        ix.append(x[b:b + step], None if ids is None else ids[b:b + step])
    ix.seal()
    return ix


# ---------------------------------------------------------------- BASELINE config 1
def test_config1_1k_x512_f32_cosine_top10_single_query():
    x, q = orc.synthetic(1000, 512), orc.synthetic(1, 512, orc.CORPUS_SEED + 1)
    with build(x, pk.F32) as ix:
        got = ix.search(q, 10, pk.COSINE)
    want = orc.topk(x, q, orc.COSINE, 10)
    assert_close_topk(got, want, x, q, orc.COSINE)
    assert np.array_equal(got[0], want[0])  # gaps at N=1000 are far wider than 1e-5


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq", [1, 3, 8, 33])
def test_f32_matches_oracle(metric, nq):
    x, q = orc.synthetic(20000, 768, 11), orc.synthetic(nq, 768, 12)
    with build(x, pk.F32, chunks=3) as ix:
        got = ix.search(q, 100, metric)
    assert_close_topk(got, orc.topk(x, q, metric, 100, threads=8), x, q, metric)


@pytest.mark.parametrize("dim", [8, 100, 512, 1024, 1536])
def test_f32_dims_and_ragged_rows(dim):
    x, q = orc.synthetic(3001, dim, 21, normalise=False), orc.synthetic(5, dim, 22, normalise=False)
    with build(x, pk.F32) as ix:
        for metric in (pk.L2, pk.COSINE):
            assert_close_topk(ix.search(q, 17, metric), orc.topk(x, q, metric, 17, threads=4), x, q, metric)


# ---------------------------------------------------------------- int8: bit exact
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq", [1, 4, 17])
def test_int8_bit_exact(metric, nq):
    x, q, scale, xc, qc = int8_space(50000, 768, 31, nq)
    with build(xc, pk.I8, scale=scale, chunks=2) as ix:
        got_codes = ix.search(qc, 100, metric)
        got_f32q = ix.search(q, 100, metric)  # f32 queries quantised on the GPU with the index scale
    want = orc.topk(xc, qc, metric, 100, threads=8)
    assert_exact(got_codes, want)
    assert_exact(got_f32q, want)


@pytest.mark.parametrize("dim", [8, 128, 520, 1024])
def test_int8_dims_and_saturated_codes(dim):
    rng = np.random.default_rng(5)
    xc = rng.integers(-128, 128, size=(2500, dim), dtype=np.int8)
    xc[7] = -128   # extreme norms
    xc[8] = 127
    xc[9] = 0      # zero row: NaN cosine, ordered last
    qc = rng.integers(-128, 128, size=(6, dim), dtype=np.int8)
    qc[0] = -128
    with build(xc, pk.I8) as ix:
        for metric in METRICS:
            assert_exact(ix.search(qc, 2500 if dim == 8 else 64, metric), orc.topk(xc, qc, metric, 2500 if dim == 8 else 64, threads=4))


# ---------------------------------------------------------------- reference KAT fixtures on the GPU
def test_reference_order_parity_fixture_on_gpu():
    g = GOLD["order_parity"]
    seeded = np.array(g["vectors"], np.float32)
    space = np.concatenate([seeded, np.tile(np.array(g["filler"], np.float32), (g["total_vectors"] - len(seeded), 1))])
    q = np.array([g["query"]], np.float32)
    n = len(seeded)
    absmax = pk.blob_absmax(space)
    assert absmax == 11.0
    scale = pk.scale_from_absmax(absmax)
    codes = pk.quantize_int8(space, scale)
    assert np.array_equal(codes, orc.quantize_rows(space, scale))
    with build(seeded, pk.F32) as fx, build(codes[:n], pk.I8, scale=scale) as qx:
        exact = fx.search(q, n, pk.COSINE)
        quant = qx.search(q, n, pk.COSINE)
        assert list(exact[0][0]) == list(quant[0][0])  # int8 ordering == exact ordering
        assert list(quant[0][0]) == list(orc.topk(codes[:n], orc.quantize_rows(q, scale), orc.COSINE, n)[0][0])
        again = qx.search(q, n, pk.COSINE)
        assert np.array_equal(quant[0], again[0])  # deterministic
        for k in (1, 3):  # k only truncates
            assert list(qx.search(q, k, pk.COSINE)[0][0]) == list(quant[0][0][:k])


def test_reference_int8_distance_kats_on_gpu():
    g = GOLD["int8_distances"]
    for case in g["cases"]:
        a = np.array([case["left"]], np.int8)
        b = np.array([case["right"]], np.int8)
        with build(a, pk.I8) as ix:
            l2 = ix.search(b, 1, pk.L2)[1][0][0]
            cos = ix.search(b, 1, pk.COSINE)[1][0][0]
        assert abs(l2 - case["l2"]) <= g["l2_rel_tol"] * max(case["l2"], 1.0)
        assert abs(cos - case["cosine"]) <= g["cosine_abs_tol"]
        assert l2 == np.float32(case["l2"]) and cos == np.float32(case["cosine"])


def test_reference_similar_to_fixture_on_gpu():
    g = GOLD["similar_to"]
    seeded = np.array(g["vectors"], np.float32)
    space = np.concatenate([seeded, np.tile(np.array(g["filler"], np.float32), (g["total_vectors"] - len(seeded), 1))])
    scale = pk.scale_from_absmax(pk.blob_absmax(space))
    codes = pk.quantize_int8(space, scale)
    n, t = len(seeded), g["target_index"]
    member = np.zeros(1, np.uint64)
    for i in range(n):
        if i != t:
            member[0] |= np.uint64(1) << np.uint64(i)   # "other_embeddings.sha256 != target"
    with build(seeded, pk.F32) as fx, build(codes[:n], pk.I8, scale=scale) as qx:
        exact = fx.search(seeded[t:t + 1], n, pk.L2, bitmap=member)
        quant = qx.search(codes[t:t + 1], n, pk.L2, bitmap=member)
    assert exact[2][0] == 7 and quant[2][0] == 7
    assert list(exact[0][0][:7]) == list(quant[0][0][:7])


# ---------------------------------------------------------------- codec on the GPU
def test_gpu_codec_bit_exact():
    for case in GOLD["codec"]["quantize"]:
        scale = case.get("scale") or pk.scale_from_absmax(case["scale_from_absmax"])
        assert list(pk.quantize_int8(np.array(case["values"], np.float32), scale)) == case["codes"]
    rng = np.random.default_rng(9)
    x = (rng.standard_normal(300001) * 3).astype(np.float32)
    x[:8] = [np.nan, np.inf, -np.inf, 0.5, 1.5, 2.5, -0.5, 1e30]
    for scale in (1.0, 0.0371, float(np.float32(np.abs(x[8:]).max()) / np.float32(127))):
        assert np.array_equal(pk.quantize_int8(x, scale), orc.np_quantize_int8(x, scale))
        assert pk.quantize_int8(x, scale).tobytes() == orc.quantize_int8(x.tobytes(), scale)
    assert pk.blob_absmax(x[3:]) == orc.blob_absmax(x[3:].tobytes())
    assert pk.blob_absmax(np.array([1.0, np.nan, -2.0], np.float32)) == 2.0
    assert pk.blob_absmax(np.zeros(0, np.float32)) == 0.0


# ---------------------------------------------------------------- edge cases
def test_empty_ragged_and_oversized_k():
    q = orc.synthetic(2, 64, 2)
    with pk.VectorIndex(64, pk.F32) as ix:
        ix.seal()
        ids, dist, cnt = ix.search(q, 5, pk.COSINE)
        assert np.all(ids == -1) and np.all(np.isnan(dist)) and list(cnt) == [0, 0]
        x = orc.synthetic(7, 64, 3)
        ix.append(x)
        with pytest.raises(pk.PkvError) as e:
            ix.search(q, 5, pk.COSINE)
        assert e.value.status == 3  # not sealed
        ix.seal()
        got = ix.search(q, 10, pk.L2)   # k > N
        assert list(got[2]) == [7, 7]
        assert_close_topk(got, orc.topk(x, q, orc.L2, 10), x, q, orc.L2)
        ids0, _, cnt0 = ix.search(np.zeros((0, 64), np.float32), 3, pk.L2)
        assert ids0.shape == (0, 3)


def test_ties_nan_rows_row_ids_and_appends():
    x = np.zeros((6, 4), np.float32)
    x[0] = [1, 0, 0, 0]; x[1] = [0, 1, 0, 0]; x[2] = [1, 0, 0, 0]; x[4] = [2, 0, 0, 0]; x[5] = [0, 1, 0, 0]
    q = np.array([[1, 0, 0, 0]], np.float32)
    ids = np.array([100, 7, 42, 9, 1000, 5], np.int64)
    with build(x, pk.F32, ids=ids, chunks=3) as ix:
        got = ix.search(q, 8, pk.COSINE)
    want_rows = [0, 2, 4, 1, 5, 3]  # ties by insertion position, the zero row (NaN) last
    assert list(got[0][0][:6]) == [int(ids[r]) for r in want_rows] and list(got[0][0][6:]) == [-1, -1]
    assert got[2][0] == 6 and np.isnan(got[1][0][5])
    # ids appended only on a later batch: earlier rows keep position ids
    with pk.VectorIndex(4, pk.F32) as ix:
        ix.append(x[:3]); ix.append(x[3:], ids[3:]); ix.seal()
        got = ix.search(q, 6, pk.COSINE)
    assert list(got[0][0]) == [0, 2, 1000, 1, 5, 9]


@pytest.mark.parametrize("dtype", ["f32", "i8"])
def test_bitmap_filter_shared_and_per_query(dtype):
    n, d, nq = 30000, 128, 5
    x, q, scale, xc, qc = int8_space(n, d, 41, nq)
    data, queries, code = (x, q, pk.F32) if dtype == "f32" else (xc, qc, pk.I8)
    rng = np.random.default_rng(43)
    words = (n + 63) // 64
    for p in (0.5, 0.01):
        shared = np.packbits(rng.random(words * 64) < p, bitorder="little").view(np.uint64)
        per_q = np.stack([np.packbits(rng.random(words * 64) < p, bitorder="little").view(np.uint64) for _ in range(nq)])
        with build(data, code, scale=scale if code == pk.I8 else None) as ix:
            for bm, stride in ((shared, 0), (per_q.reshape(-1), words)):
                got = ix.search(queries, 50, pk.COSINE, bitmap=bm, bitmap_stride_words=stride)
                want = orc.topk(data, queries, orc.COSINE, 50, bitmap=bm, bitmap_stride=stride, threads=4)
                if code == pk.I8:
                    assert_exact(got, want)
                else:
                    assert np.array_equal(got[2], want[2])
                    assert np.allclose(got[1], want[1], rtol=1e-5, atol=1e-7, equal_nan=True)


def test_adversarial_order_forces_overflow_rescan():
    # rows sorted from worst to best: every new row beats the running threshold, so candidate
    # buffers overflow and the driver has to split ranges; results must not change
    n, d = 60000, 32
    x = orc.synthetic(n, d, 51)
    q = orc.synthetic(3, d, 52)
    order = np.argsort(-(x @ q[0]))[::-1]  # ascending similarity to query 0
    x = np.ascontiguousarray(x[order])
    with build(x, pk.F32) as ix:
        ix.set_option("candidate_capacity", 256)
        got = ix.search(q, 20, pk.COSINE)
        assert ix.counters().fallback_queries > 0
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 20, threads=3), x, q, orc.COSINE)


def test_duplicates_everywhere():
    # many identical rows: ties resolved by row position, bit-exact on int8
    base = np.random.default_rng(6).integers(-100, 100, size=(5, 64), dtype=np.int8)
    xc = np.ascontiguousarray(np.tile(base, (4000, 1)))
    qc = base[:2].copy()
    with build(xc, pk.I8) as ix:
        assert_exact(ix.search(qc, 300, pk.COSINE), orc.topk(xc, qc, orc.COSINE, 300, threads=2))
        assert_exact(ix.search(qc, 300, pk.L2), orc.topk(xc, qc, orc.L2, 300, threads=2))


def test_large_k():
    x, q, scale, xc, qc = int8_space(40000, 256, 61, 3)
    with build(xc, pk.I8) as ix:
        assert_exact(ix.search(qc, 4096, pk.COSINE), orc.topk(xc, qc, orc.COSINE, 4096, threads=3))
        with pytest.raises(pk.PkvError):
            ix.search(qc, 4097, pk.COSINE)
        with pytest.raises(pk.PkvError, match="positive integer"):
            ix.search(qc, 0, pk.COSINE)


def test_f16_extension_matches_oracle():
    x = orc.synthetic(20000, 512, 71).astype(np.float16)
    q = orc.synthetic(9, 512, 72).astype(np.float16)
    with build(x, pk.F16) as ix:
        for metric in (pk.COSINE, pk.L2):
            got = ix.search(q, 100, metric)
            assert_close_topk(got, orc.topk(x, q, metric, 100, threads=8), x, q, metric)


def test_dim_mismatch_and_bad_args():
    with pk.VectorIndex(64, pk.F32) as ix:
        ix.append(orc.synthetic(10, 64)); ix.seal()
        with pytest.raises(pk.PkvError) as e:
            ix.search(np.zeros((1, 32), np.float32), 3)
        assert e.value.status == 2
        with pytest.raises(pk.PkvError):
            ix.search(np.zeros((1, 64), np.int8), 3)   # int8 codes against an f32 index
    with pk.VectorIndex(64, pk.I8) as ix:
        ix.append(np.zeros((4, 64), np.int8)); ix.seal()
        with pytest.raises(pk.PkvError) as e:   # f32 queries need the scale artifact
            ix.search(np.zeros((1, 64), np.float32), 3)
        assert e.value.status == 3
        with pytest.raises(pk.PkvError):
            ix.set_scale_artifact(b"\0\0\0\0")


def test_concurrent_searches_from_many_threads():
    # the reference's read pool runs 16 connection threads (db/connection.rs:235): pkv_search is re-entrant
    import threading

    x, q, scale, xc, qc = int8_space(60000, 256, 171, 64)
    want = orc.topk(xc, qc, orc.COSINE, 50, threads=16)
    with build(xc, pk.I8, scale=scale) as ix:
        results, errors = [None] * 16, []

        def worker(t):
            try:
                sl = slice(t * 4, t * 4 + 4)
                for _ in range(5):
                    results[t] = ix.search(qc[sl], 50, pk.COSINE)
            except Exception as e:  # pragma: no cover
                errors.append(e)

        threads = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
        [t.start() for t in threads]
        [t.join() for t in threads]
        assert not errors, errors
        for t in range(16):
            sl = slice(t * 4, t * 4 + 4)
            assert_exact(results[t], (want[0][sl], want[1][sl], want[2][sl]))


def test_more_queries_than_one_internal_pass():
    x, q, scale, xc, qc = int8_space(20000, 128, 181, 1500)
    with build(xc, pk.I8, scale=scale) as ix:
        assert_exact(ix.search(qc, 10, pk.L2), orc.topk(xc, qc, orc.L2, 10, threads=16))
    with build(x, pk.F32) as ix:
        assert_close_topk(ix.search(q, 10, pk.COSINE), orc.topk(x, q, orc.COSINE, 10, threads=16), x, q, orc.COSINE)


def test_int8_wide_rows_fall_back_to_cuda_cores():
    rng = np.random.default_rng(19)
    xc = rng.integers(-128, 128, size=(3000, 1536), dtype=np.int8)
    xc[5] = -128
    xc[6] = 127
    qc = rng.integers(-128, 128, size=(12, 1536), dtype=np.int8)
    qc[0] = 127
    with build(xc, pk.I8) as ix:
        for metric in METRICS:
            assert_exact(ix.search(qc, 40, metric), orc.topk(xc, qc, metric, 40, threads=8))
        assert ix.counters().last_scan_kind == 2


def test_packed_shard_exchange_matches_plain_merge():
    # two shards of one corpus on one GPU: pack -> (concatenation = what the all-gather delivers) -> merge of the
    # packed buffer must equal the unpacked merge and the search over the whole corpus
    import torch

    x = orc.synthetic(30000, 64, 71)
    q = orc.synthetic(37, 64, 72)
    parts = []
    for b, e in ((0, 17000), (17000, 30000)):
        ix = pk.VectorIndex(64, pk.F32)
        ix.set_row_base(b)
        ix.append(x[b:e])
        ix.seal()
        with ix:
            ids, dd, _ = ix.search(torch.from_numpy(q).cuda(), 25, pk.COSINE)
            parts.append((ids.clone(), dd.clone()))
    g_ids = torch.stack([p[0] for p in parts])
    g_dist = torch.stack([p[1] for p in parts])
    plain = pk.merge_topk(g_ids, g_dist)
    packed = torch.stack([pk.pack_topk(i, d) for i, d in parts])
    fast = pk.merge_packed(packed)
    torch.cuda.synchronize()
    assert torch.equal(plain[0], fast[0]) and torch.equal(plain[2], fast[2])
    assert torch.equal(plain[1].view(torch.int32), fast[1].view(torch.int32))
    want = orc.topk(x, q, orc.COSINE, 25, threads=8)
    assert_close_topk(tuple(t.cpu().numpy() for t in fast), want, x, q, orc.COSINE)
