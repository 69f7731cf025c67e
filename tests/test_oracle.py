"""The oracle against the reference's own known-answer tests (tests/golden/reference_kats.json,
transcribed from panoptikon/src/db/vector_quants.rs and pql/*.rs) and against its NumPy mirror."""
import json
import math
import os
import struct

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def le_bytes(values):
    return struct.pack("<%df" % len(values), *values)


def as_i8(b: bytes):
    return list(np.frombuffer(b, dtype=np.int8))


# ---- db/vector_quants.rs:3588-3626
def test_codec_rounds_ties_to_even_and_clamps():
    for case in GOLD["codec"]["quantize"]:
        scale = case.get("scale")
        if scale is None:
            scale = orc.scale_from_absmax(case["scale_from_absmax"])
        assert as_i8(orc.quantize_int8(le_bytes(case["values"]), scale)) == case["codes"]
        assert list(orc.np_quantize_int8(np.array(case["values"], np.float32), scale)) == case["codes"]
    assert orc.quantize_int8(le_bytes([1.0, 2.0, 3.0]), 1.0) == bytes([1, 2, 3])
    assert orc.scale_from_absmax(0.0) == 1.0
    assert orc.scale_from_absmax(float("nan")) == 1.0
    assert orc.scale_from_absmax(float("inf")) == 1.0
    assert orc.scale_from_absmax(-3.0) == 1.0


def test_artifact_roundtrip_and_rejections():
    scale = orc.scale_from_absmax(GOLD["codec"]["artifact_roundtrip_absmax"])
    assert orc.artifact_scale(orc.scale_artifact(scale)) == scale
    assert orc.scale_artifact(scale) == struct.pack("<f", scale)
    for hexed in GOLD["codec"]["artifact_rejects_hex"]:
        assert orc.artifact_scale(bytes.fromhex(hexed)) is None
    assert orc.artifact_scale(struct.pack("<f", float("inf"))) is None


# ---- db/vector_quants.rs:2194-2244
def test_build_scale_is_absmax_over_127():
    g = GOLD["build_scale"]
    blob = le_bytes([0.5, -g["absmax"], 3.0, 0.02])
    s = orc.scale_from_absmax(orc.blob_absmax(blob))
    assert abs(s - g["expected_scale"]) < g["tol"]
    assert float(orc.np_scale_from_absmax(np.float32(g["absmax"]))) == s
    # NaN never replaces the running max (`>` comparison)
    assert orc.blob_absmax(le_bytes([1.0, float("nan"), -2.0])) == 2.0


def test_quantize_nan_inf_semantics():
    # Rust: NaN.clamp() stays NaN, `as i8` -> 0; +-inf saturate
    codes = as_i8(orc.quantize_int8(le_bytes([float("nan"), float("inf"), float("-inf"), 127.5, -128.5, 126.5]), 1.0))
    assert codes == [0, 127, -128, 127, -128, 126]
    assert list(orc.np_quantize_int8(np.array([np.nan, np.inf, -np.inf, 127.5, -128.5, 126.5], np.float32), 1.0)) == codes


# ---- db/vector_quants.rs:3632-3687
def test_int8_distances_match_reference_kat():
    g = GOLD["int8_distances"]
    for case in g["cases"]:
        a = np.array(case["left"], np.int8)
        b = np.array(case["right"], np.int8)
        l2 = orc.distance(a, b, orc.L2)
        cos = orc.distance(a, b, orc.COSINE)
        assert abs(l2 - case["l2"]) <= g["l2_rel_tol"] * max(case["l2"], 1.0)
        assert abs(cos - case["cosine"]) <= g["cosine_abs_tol"]
        # oracle is one f32 rounding away from the f64 formula on integer-exact sums
        assert l2 == np.float32(case["l2"])
        assert cos == np.float32(case["cosine"])


def _space(fixture):
    seeded = np.array(fixture["vectors"], np.float32)
    filler = np.tile(np.array(fixture["filler"], np.float32), (fixture["total_vectors"] - len(seeded), 1))
    return seeded, np.concatenate([seeded, filler], axis=0)


# ---- db/vector_quants.rs:3254-3278,3324-3382
def test_order_parity_int8_equals_f32_on_separated_vectors():
    g = GOLD["order_parity"]
    seeded, space = _space(g)
    q = np.array([g["query"]], np.float32)
    n = len(seeded)
    scale = orc.scale_from_absmax(orc.blob_absmax(space.tobytes()))
    assert abs(scale - 11.0 / 127.0) < 1e-6
    codes = orc.quantize_rows(space, scale)
    qcodes = orc.quantize_rows(q, scale)
    rows_f, dist_f, _ = orc.topk(seeded, q, orc.COSINE, n)
    rows_q, dist_q, _ = orc.topk(codes[:n], qcodes, orc.COSINE, n)
    assert list(rows_f[0]) == list(rows_q[0])
    # well separated: no ties anywhere in either ordering
    assert len(set(dist_f[0])) == n and len(set(dist_q[0])) == n
    # deterministic and independent of k (k only truncates)
    for k in (1, 3, n):
        r2, _, c2 = orc.topk(codes[:n], qcodes, orc.COSINE, k)
        assert list(r2[0]) == list(rows_q[0][:k]) and c2[0] == k
    # page walk == single shot (:3386-3415): pages are prefixes of one total order
    walked = []
    for page in range(3):
        walked.extend(rows_q[0][page * 4:(page + 1) * 4])
    assert walked == list(rows_q[0])


# ---- db/vector_quants.rs:3532-3582
def test_similar_to_int8_order_equals_f32():
    g = GOLD["similar_to"]
    seeded, space = _space(g)
    scale = orc.scale_from_absmax(orc.blob_absmax(space.tobytes()))
    codes = orc.quantize_rows(space, scale)
    n = len(seeded)
    t = g["target_index"]
    others = [i for i in range(n) if i != t]
    rf, _, _ = orc.topk(seeded[others], seeded[t:t + 1], orc.L2, n - 1)
    rq, _, _ = orc.topk(codes[:n][others], codes[t:t + 1], orc.L2, n - 1)
    assert len(rf[0]) == 7 and list(rf[0]) == list(rq[0])


def test_query_blob_layout():
    g = GOLD["query_blob"]
    assert np.array(g["values"], "<f4").tobytes().hex() == g["le_hex"]


# ---- C restatement vs NumPy mirror (independent second restatement)
@pytest.mark.parametrize("metric", [orc.L2, orc.COSINE, orc.DOT])
@pytest.mark.parametrize("dtype", ["f32", "i8", "f16"])
def test_c_oracle_equals_numpy_mirror(metric, dtype):
    x = orc.synthetic(300, 96, seed=7)
    q = orc.synthetic(3, 96, seed=8)
    if dtype == "i8":
        s = orc.scale_from_absmax(float(np.abs(x).max()))
        x, q = orc.quantize_rows(x, s), orc.quantize_rows(q, s)
    elif dtype == "f16":
        x, q = x.astype(np.float16), q.astype(np.float16)
    for qi in range(q.shape[0]):
        dc = orc.distances(x, q[qi], metric)
        dn = orc.np_distances(x, q[qi], metric)
        assert np.array_equal(dc, dn, equal_nan=True)
    rows_c, dist_c, cnt_c = orc.topk(x, q, metric, 17, threads=2)
    rows_n, dist_n, cnt_n = orc.np_topk(x, q, metric, 17)
    assert np.array_equal(rows_c, rows_n) and np.array_equal(dist_c, dist_n, equal_nan=True)
    assert np.array_equal(cnt_c, cnt_n)


def test_topk_ties_nan_and_bitmap():
    # duplicates tie -> ascending row; zero row -> NaN cosine -> last; k > members pads with -1/NaN
    x = np.zeros((6, 4), np.float32)
    x[0] = [1, 0, 0, 0]
    x[1] = [0, 1, 0, 0]
    x[2] = [1, 0, 0, 0]
    x[4] = [2, 0, 0, 0]  # same direction as rows 0 and 2
    x[5] = [0, 1, 0, 0]
    q = np.array([[1, 0, 0, 0]], np.float32)
    rows, dist, cnt = orc.topk(x, q, orc.COSINE, 8)
    assert list(rows[0]) == [0, 2, 4, 1, 5, 3, -1, -1]
    assert cnt[0] == 6 and math.isnan(dist[0][5]) and math.isnan(dist[0][6])
    bitmap = np.array([0b110110], np.uint64)  # rows 1,2,4,5
    rows, dist, cnt = orc.topk(x, q, orc.COSINE, 3, bitmap=bitmap)
    assert list(rows[0]) == [2, 4, 1] and cnt[0] == 3
    rn, dn, cn = orc.np_topk(x, q, orc.COSINE, 3, bitmap=bitmap)
    assert np.array_equal(rows, rn)
    # empty corpus
    rows, dist, cnt = orc.topk(np.zeros((0, 4), np.float32), q, orc.L2, 2)
    assert list(rows[0]) == [-1, -1] and cnt[0] == 0


def test_int8_sums_are_integer_exact_up_to_dim_1024():
    # f32 accumulators hold the int sums exactly (< 2^24), so integer
    # arithmetic on the GPU reproduces the oracle bit for bit
    rng = np.random.default_rng(3)
    a = rng.integers(-128, 128, size=(64, 1024), dtype=np.int8)
    a[0, :] = -128
    b = a[::-1].copy()
    for i in range(8):
        ai, bi = a[i].astype(np.int64), b[i].astype(np.int64)
        dot, am, bm = int((ai * bi).sum()), int((ai * ai).sum()), int((bi * bi).sum())
        want = np.float32(1.0 - dot / (math.sqrt(am) * math.sqrt(bm)))
        assert orc.distance(a[i], b[i], orc.COSINE) == want
        assert orc.distance(a[i], b[i], orc.DOT) == np.float32(-dot)


def test_aggregate_min_max_avg_weighted():
    d = np.array([1.0, 3.0, 2.0, np.nan, 5.0], np.float32)
    item = np.array([0, 0, 1, 1, 2], np.int64)
    assert list(orc.aggregate(d, item, 4, orc.AGG_MIN))[:3] == [1.0, 2.0, 5.0]
    assert list(orc.aggregate(d, item, 4, orc.AGG_MAX))[:3] == [3.0, 2.0, 5.0]
    avg = orc.aggregate(d, item, 4, orc.AGG_AVG)
    assert avg[0] == 2.0 and avg[1] == 2.0 and math.isnan(avg[3])
    w = np.array([1.0, 3.0, 1.0, 1.0, 2.0], np.float32)
    wa = orc.aggregate(d, item, 4, orc.AGG_AVG, weights=w)
    assert wa[0] == (1.0 * 1 + 3.0 * 3) / 4.0
