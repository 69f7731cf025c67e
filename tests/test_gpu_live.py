"""Live mode: ONE scan launch over every row behind the bootstrap chunk, the per-query
thresholds re-read tile by tile and re-selected in-kernel, the int8-image survivors re-scored by warps of the scan
kernel itself.  Results must equal the chunked schedule bit for bit (same exact keys, same summation order) and
the oracle (int8: bit-exact, f32: 1e-5 relative)."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk, assert_exact, int8_space

pytestmark = pytest.mark.gpu
METRICS = [pk.L2, pk.COSINE, pk.DOT]


def _same(a, b):
    return (np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
            and np.array_equal(a[2], b[2]))


def _f32_index(x):
    ix = pk.VectorIndex(x.shape[1], pk.F32)
    ix.set_option("live", 2)   # live whenever the kernels can (the default waits for corpora of millions of rows)
    ix.append(x)
    ix.seal()
    return ix


def _i8_index(xc, scale):
    ix = pk.VectorIndex(xc.shape[1], pk.I8)
    ix.set_option("live", 2)
    ix.set_scale_artifact(pk.scale_artifact(scale))
    ix.append(xc)
    ix.seal()
    return ix


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq", [1, 100, 256, 300])
def test_live_f32_equals_chunked_and_oracle(metric, nq):
    x, q = orc.synthetic(150_001, 256, 301), orc.synthetic(nq, 256, 302)
    with _f32_index(x) as ix:
        c0 = ix.counters()
        live = ix.search(q, 100, metric)
        c1 = ix.counters()
        assert c1.last_scan_kind == 8
        assert c1.fallback_queries == c0.fallback_queries, "the live launch overflowed on benign data"
        assert c1.rescored_pairs > c0.rescored_pairs, "the scan kernel's own warps re-scored nothing"
        ix.set_option("live_start_rows", 4096)   # live right behind the bootstrap chunk: the hardest start
        early = ix.search(q, 100, metric)
        ix.set_option("live", 0)
        chunked = ix.search(q, 100, metric)
        ix.set_option("img8_fused", 2)   # every chunk re-scored in-kernel
        fused = ix.search(q, 100, metric)
        unfused = chunked
    assert _same(live, chunked), "live and chunked schedules disagree"
    assert _same(early, chunked), "an early live start changed the result"
    assert _same(fused, unfused), "in-kernel and separate re-scoring disagree"
    assert_close_topk(live, orc.topk(x, q, metric, 100, threads=16), x, q, metric)


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("nq", [1, 128, 300, 1030])
def test_live_int8_bit_exact(metric, nq):
    x, q, scale, xc, qc = int8_space(150_001, 256, seed=311, nq=nq)
    with _i8_index(xc, scale) as ix:
        c0 = ix.counters()
        live = ix.search(qc, 100, metric)
        c1 = ix.counters()
        assert c1.last_scan_kind == 3
        assert c1.fallback_queries == c0.fallback_queries
        ix.set_option("live_start_rows", 4096)
        early = ix.search(qc, 100, metric)
        ix.set_option("live", 0)
        chunked = ix.search(qc, 100, metric)
    assert _same(live, chunked)
    assert _same(early, chunked)
    assert_exact(live, orc.topk(xc, qc, metric, 100, threads=16))


@pytest.mark.parametrize("k", [1, 10, 128])
def test_live_depths(k):
    x, q = orc.synthetic(90_000, 768, 321), orc.synthetic(130, 768, 322)
    with _f32_index(x) as ix:
        got = ix.search(q, k, pk.COSINE)
        assert ix.counters().fallback_queries == 0
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, k, threads=16), x, q, orc.COSINE)
    x, q, scale, xc, qc = int8_space(90_000, 768, seed=323, nq=130)
    with _i8_index(xc, scale) as ix:
        assert_exact(ix.search(qc, k, pk.COSINE), orc.topk(xc, qc, orc.COSINE, k, threads=16))


def test_live_ties_and_duplicates():
    # every row occurs 3000 times: the k-th best sits inside a huge tie group, the in-kernel selection has to resolve
    # the ties by row exactly like the final select does
    base = np.random.default_rng(6).integers(-100, 100, size=(5, 64), dtype=np.int8)
    xc = np.ascontiguousarray(np.tile(base, (3000, 1)))
    qc = base[:3].copy()
    with pk.VectorIndex(64, pk.I8) as ix:
        ix.set_option("live", 2)
        ix.append(xc)
        ix.seal()
        for metric in (pk.COSINE, pk.L2):
            assert_exact(ix.search(qc, 100, metric), orc.topk(xc, qc, metric, 100, threads=3))
    xf = np.ascontiguousarray(np.tile(orc.synthetic(7, 128, 331), (2500, 1)))
    qf = orc.synthetic(9, 128, 332)
    with _f32_index(xf) as ix:
        got = ix.search(qf, 100, pk.COSINE)
    assert_close_topk(got, orc.topk(xf, qf, orc.COSINE, 100, threads=8), xf, qf, orc.COSINE)


def test_live_adversarial_order_falls_back():
    # rows sorted from worst to best for query 0: every row beats the running threshold, the live launch overflows
    # its candidate buffer and the search is redone on the careful chunked schedule
    n, d = 80_000, 64
    x = orc.synthetic(n, d, 341)
    q = orc.synthetic(5, d, 342)
    order = np.argsort(x @ q[0])   # ascending similarity = descending cosine distance
    x = np.ascontiguousarray(x[order])
    with _f32_index(x) as ix:
        ix.set_option("guess", 0)          # learnt thresholds: the prefix sees only the worst rows
        ix.set_option("live_start_rows", 4096)
        got = ix.search(q, 50, pk.COSINE)
        assert ix.counters().fallback_queries > 0
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 50, threads=5), x, q, orc.COSINE)


@pytest.mark.parametrize("density", [0.5, 0.02, 0.0005])
def test_live_with_membership_bitmap(density):
    # dense, sparse, and "fewer members than k" contexts (the AND of a PQL filter, image_embeddings.rs:140-199)
    rng = np.random.default_rng(35)
    x, q = orc.synthetic(120_000, 128, 351), orc.synthetic(40, 128, 352)
    bm = np.packbits(rng.random(((len(x) + 63) // 64) * 64) < density, bitorder="little").view(np.uint64)
    with _f32_index(x) as ix:
        c0 = ix.counters()
        got = ix.search(q, 100, pk.COSINE, bitmap=bm)
        launches = ix.counters().kernel_launches - c0.kernel_launches
    want = orc.topk(x, q, orc.COSINE, 100, bitmap=bm, threads=8)
    assert np.array_equal(got[2], want[2])
    assert_close_topk(got, want, x, q, orc.COSINE)
    assert launches < 200, f"{launches} launches: the threshold-less bitmap chunks crawl"
    x, q, scale, xc, qc = int8_space(120_000, 128, seed=353, nq=40)
    with _i8_index(xc, scale) as ix:
        assert_exact(ix.search(qc, 100, pk.L2, bitmap=bm), orc.topk(xc, qc, orc.L2, 100, bitmap=bm, threads=8))


def test_f16_index_live_rescoring():
    x = orc.synthetic(100_000, 512, 361).astype(np.float16)
    q = orc.synthetic(300, 512, 362).astype(np.float16)
    with pk.VectorIndex(512, pk.F16) as ix:
        ix.set_option("live", 2)
        ix.append(x)
        ix.seal()
        live = ix.search(q, 100, pk.COSINE)
        assert ix.counters().last_scan_kind == 8
        ix.set_option("live", 0)
        old = ix.search(q, 100, pk.COSINE)
    assert _same(live, old)
    xf, qf = x.astype(np.float32), q.astype(np.float32)
    assert_close_topk(live, orc.topk(xf, qf, orc.COSINE, 100, threads=16), xf, qf, orc.COSINE)


def test_device_api_orders_behind_the_callers_stream():
    # queries produced by a kernel on torch's current stream immediately before the search (no host
    # synchronisation in between) must be seen by the search
    import torch

    x = orc.synthetic(50_000, 256, 371)
    with _f32_index(x) as ix:
        dev = torch.device("cuda", 0)
        base = torch.from_numpy(orc.synthetic(64, 256, 372)).to(dev)
        torch.cuda.synchronize()
        big = torch.randn((4096, 4096), device=dev)
        for _ in range(3):   # keep the stream busy so the producer kernel below is still pending
            big = big @ big
            big = big / big.norm()
        q = (base * 2.0 + 0.0).contiguous()   # produced asynchronously on the current stream
        ids, dist, cnt = ix.search(q, 10, pk.COSINE)
        got = (ids.cpu().numpy(), dist.cpu().numpy(), cnt.cpu().numpy())
    qh = (base.cpu().numpy() * 2.0).astype(np.float32)
    assert_close_topk(got, orc.topk(x, qh, orc.COSINE, 10, threads=8), x, qh, orc.COSINE)


def test_guessed_start_equals_learnt_thresholds():
    """Mid-size corpora start the live launch from thresholds GUESSED on a strided sample and verify them at the end."""
    x, q = orc.synthetic(200_000, 256, 381), orc.synthetic(256, 256, 382)
    with _f32_index(x) as ix:
        c0 = ix.counters()
        guessed = ix.search(q, 100, pk.COSINE)
        c1 = ix.counters()
        assert c1.fallback_queries == c0.fallback_queries, "a guess failed on benign data"
        assert c1.kernel_launches - c0.kernel_launches <= 8   # prep, reset, sample scan, guess select, codes, live, deferred, select
        ix.set_option("guess", 0)
        learnt = ix.search(q, 100, pk.COSINE)
        assert ix.counters().kernel_launches - c1.kernel_launches > 8
    assert _same(guessed, learnt)
    assert_close_topk(guessed, orc.topk(x, q, orc.COSINE, 100, threads=16), x, q, orc.COSINE)
    x, q, scale, xc, qc = int8_space(200_000, 256, seed=383, nq=300)
    with _i8_index(xc, scale) as ix:
        guessed = ix.search(qc, 100, pk.L2)
        assert ix.counters().fallback_queries == 0
    assert_exact(guessed, orc.topk(xc, qc, orc.L2, 100, threads=16))


def test_guess_too_tight_is_detected_and_redone():
    # the only rows close to query 0 sit exactly on the sample's stride: the sample is far better than the corpus, the
    # guessed threshold admits fewer than k rows, the end-of-search check notices and the search is redone
    n, d, k = 150_001, 64, 100
    x, q = orc.synthetic(n, d, 391), orc.synthetic(4, d, 392)
    stride = n // 3968
    rng = np.random.default_rng(39)
    for j in range(60):
        v = q[0] + rng.standard_normal(d).astype(np.float32) * 0.01
        x[j * stride * 3] = v / np.linalg.norm(v)
    with _f32_index(x) as ix:
        got = ix.search(q, k, pk.COSINE)
        assert ix.counters().fallback_queries > 0, "the too-tight guess went unnoticed"
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, k, threads=4), x, q, orc.COSINE)


def test_guess_survives_sorted_insertion_order():
    # rows inserted from worst to best for query 0: a prefix sample would be hopeless, the strided one is not
    n, d = 120_000, 64
    x, q = orc.synthetic(n, d, 395), orc.synthetic(5, d, 396)
    x = np.ascontiguousarray(x[np.argsort(x @ q[0])])
    with _f32_index(x) as ix:
        got = ix.search(q, 50, pk.COSINE)
        assert ix.counters().fallback_queries == 0
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, 50, threads=5), x, q, orc.COSINE)


@pytest.mark.parametrize("metric", METRICS)
def test_exact_epilogue_culling_equals_flush_time_culling(metric):
    """img8_epi = 1 (default): the epilogue bounds each surviving pair exactly in registers (row figures shuffled from
    the lane that prefetched them) before it is held; 0: the same bound at flush time from the figures in global
    memory.  Same filter, same exact keys: identical results, live and chunked."""
    x, q = orc.synthetic(220_000, 320, 501), orc.synthetic(257, 320, 502)
    x[1000:1010] *= 30.0          # rows whose figures differ from their neighbours' (the warp-wide bound is loose there)
    x[5000] = np.nan
    with _f32_index(x) as ix:
        a = ix.search(q, 100, metric)
        ix.set_option("img8_epi", 0)
        b = ix.search(q, 100, metric)
        ix.set_option("live", 0)
        c = ix.search(q, 100, metric)
        ix.set_option("img8_epi", 1)
        d = ix.search(q, 100, metric)
    assert _same(a, b) and _same(c, d) and _same(a, c)
    assert_close_topk(a, orc.topk(x, q, metric, 100, threads=16), x, q, metric)


def test_many_tight_guesses_fall_back_to_learnt_thresholds():
    # rows close to each of 40 queries sit on the sample's stride: more failed guesses than the repair takes one by one
    n, d, k = 150_001, 64, 100
    x, q = orc.synthetic(n, d, 393), orc.synthetic(48, d, 394)
    stride = n // 3968
    rng = np.random.default_rng(40)
    for i in range(40):
        for j in range(50):
            v = q[i] + rng.standard_normal(d).astype(np.float32) * 0.01
            x[(i * 50 + j) * stride] = v / np.linalg.norm(v)
    with _f32_index(x) as ix:
        c0 = ix.counters()
        got = ix.search(q, k, pk.COSINE)
        assert ix.counters().fallback_queries > c0.fallback_queries
    assert_close_topk(got, orc.topk(x, q, orc.COSINE, k, threads=8), x, q, orc.COSINE)


def test_guess_rank_follows_the_corpus_size():
    """The guessed start serves corpora from ~64k rows up; the repaired queries (if any) are counted, never wrong."""
    for n, d, nq in ((70_000, 128, 64), (400_000, 128, 200)):
        x, q = orc.synthetic(n, d, 511 + n % 7), orc.synthetic(nq, d, 512)
        with _f32_index(x) as ix:
            c0 = ix.counters()
            got = ix.search(q, 100, pk.COSINE)
            c1 = ix.counters()
            assert c1.kernel_launches - c0.kernel_launches <= 8 or c1.fallback_queries > c0.fallback_queries
        assert_close_topk(got, orc.topk(x, q, orc.COSINE, 100, threads=16), x, q, orc.COSINE)
