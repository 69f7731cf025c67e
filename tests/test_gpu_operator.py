"""GPU tests of the pieces around the scan: PQL policy (Space), shard merge, aggregation, device API."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from panoptikon_b200 import _native as N
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


def _rank_reference(dist_rows, group_of_row, n_groups, agg, weights=None):
    """NumPy restatement of GROUP BY + ORDER BY AGG(d) ASC NULLS LAST over oracle distances."""
    nq, n = dist_rows.shape
    flat_d = dist_rows.reshape(-1)
    flat_g = np.tile(group_of_row, nq)
    flat_w = None if weights is None else np.tile(weights, nq)
    ok = flat_g >= 0
    a = orc.aggregate(flat_d[ok], flat_g[ok], n_groups, agg, None if flat_w is None else flat_w[ok])
    present = np.zeros(n_groups, bool)
    present[flat_g[ok]] = True
    ids = np.nonzero(present)[0]
    nan = np.isnan(a[ids])
    order = ids[np.lexsort((ids, np.where(nan, np.inf, a[ids]), nan))]
    return order, a


@pytest.mark.parametrize("dtype", ["f32", "i8"])
def test_grouped_operator_rank_matches_sql_semantics(dtype):
    import torch

    n, d, n_groups = 40009, 96, 6000
    x, q, scale, xc, qc = int8_space(n, d, 131, 2)
    data, queries, code = (x, q, pk.F32) if dtype == "f32" else (xc, qc, pk.I8)
    rng = np.random.default_rng(7)
    group = rng.integers(0, n_groups - 50, n).astype(np.int64)   # the last 50 groups own no row: absent
    group[rng.random(n) < 0.2] = -1                               # rows outside the context CTE
    data = data.copy()
    data[group == 17] = 0                                         # a group with only zero vectors: NULL cosine
    w = (rng.random(n) + 0.05).astype(np.float32)
    with pk.VectorIndex(d, code) as ix:
        ix.append(data); ix.seal()
        tg, tw = torch.from_numpy(group).cuda(), torch.from_numpy(w).cuda()
        for metric in (pk.COSINE, pk.L2):
            dist = np.stack([orc.distances(data, queries[0], metric)])
            for agg, weights in ((pk.AGG_MIN, None), (pk.AGG_MAX, None), (pk.AGG_AVG, None), (pk.AGG_AVG, w)):
                order, a = _rank_reference(dist, group, n_groups, agg, weights)
                for offset, limit in ((0, 320), (320, 100)):
                    g, v, cnt = ix.rank_groups(torch.from_numpy(queries[:1]).cuda(), tg, n_groups, agg, metric,
                                               None if weights is None else tw, offset, limit)
                    g, v = g.cpu().numpy(), v.cpu().numpy()
                    want = order[offset:offset + limit]
                    assert cnt == len(want)
                    if dtype == "i8" and weights is None and agg != pk.AGG_AVG:
                        assert list(g[:cnt]) == list(want)          # exact distances: exact order
                    else:
                        # float sums differ in the last bits: same aggregates, order equal up to near-ties
                        assert np.allclose(v[:cnt], a[want], rtol=2e-5, atol=1e-6, equal_nan=True)
                        assert len(set(g[:cnt]) ^ set(want)) <= 4
                    assert np.allclose(v[:cnt], a[g[:cnt]], rtol=2e-5, atol=1e-6, equal_nan=True)
        # the all-NULL group is ranked last (NULLS LAST), absent groups never appear
        g, v, cnt = ix.rank_groups(torch.from_numpy(queries[:1]).cuda(), tg, n_groups, pk.AGG_MIN, pk.COSINE, None, 0, 2048)
        assert cnt == 2048 or 17 in g.cpu().numpy()[:cnt]


def test_similar_to_matches_reference_fixture_and_oracle():
    """db/vector_quants.rs:3532-3582 fixture through the operator (AVG over target x candidate pairs,
    target excluded), f32 and int8 agree; plus a multi-vector target against the NumPy restatement."""
    import json, os
    import torch

    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))["similar_to"]
    seeded = np.array(g["vectors"], np.float32)
    space = np.concatenate([seeded, np.tile(np.array(g["filler"], np.float32), (g["total_vectors"] - len(seeded), 1))])
    scale = pk.scale_from_absmax(pk.blob_absmax(space))
    codes = pk.quantize_int8(space, scale)
    n = len(seeded)
    item = torch.arange(n, dtype=torch.int64).cuda()
    orders = []
    for code, data in ((pk.F32, seeded), (pk.I8, codes[:n])):
        with pk.VectorIndex(8, code) as ix:
            ix.append(data); ix.seal()
            tgt = torch.tensor([g["target_index"]], dtype=torch.int64).cuda()
            grp = torch.where(item == g["target_index"], torch.full_like(item, -1), item).contiguous()
            items, agg, cnt = pk.similar_to(ix, tgt, grp, n, pk.L2)
            assert cnt == 7
            orders.append(list(items.cpu().numpy()[:cnt]))
    assert orders[0] == orders[1] and g["target_index"] not in orders[0]

    # multi-vector items: 3 vectors per item, AVG over all 3 x 3 pairs
    x = orc.synthetic(3000, 64, 141)
    it = np.repeat(np.arange(1000), 3).astype(np.int64)
    with pk.VectorIndex(64, pk.F32) as ix:
        ix.append(x); ix.seal()
        for agg in (pk.AGG_AVG, pk.AGG_MIN):
            itd = torch.from_numpy(it).cuda()
            tgt = torch.nonzero(itd == 42).flatten().contiguous()
            grp_d = torch.where(itd == 42, torch.full_like(itd, -1), itd).contiguous()
            items, vals, cnt = pk.similar_to(ix, tgt, grp_d, 1000, pk.COSINE, agg, limit=50)
            tq = x[it == 42]
            dist = np.stack([orc.distances(x, t, orc.COSINE) for t in tq])
            grp = np.where(it == 42, -1, it)
            order, a = _rank_reference(dist, grp, 1000, agg)
            got = items.cpu().numpy()[:cnt]
            assert cnt == 50 and 42 not in got
            assert np.allclose(vals.cpu().numpy()[:cnt], a[order[:50]], rtol=2e-5, atol=1e-6)
            assert len(set(got) ^ set(order[:50])) <= 2


def test_sqlite_table_valued_function_joins_like_the_distance_cte():
    """a15/a16/f4: pkv_topk(...) loaded into SQLite returns the (item_data.id, d) rows of one query, joined the way
    builder/filters/exact.rs:106-165 joins `embeddings`: per-file MIN(d) + row_number() rank, compared with the same
    aggregate computed from the oracle's distances."""
    import sqlite3

    rng = np.random.default_rng(77)
    n, d = 20_000, 64
    x, q = orc.synthetic(n, d, 701), orc.synthetic(1, d, 702)
    data_ids = np.arange(n, dtype=np.int64) * 3 + 11           # item_data.id of every embedding row
    file_of = rng.integers(0, 4000, size=n)                    # several embeddings per file (video frames, text chunks)
    con = sqlite3.connect(":memory:")
    con.enable_load_extension(True)
    con.load_extension(N.LIB_PATH)
    con.execute("CREATE TABLE item_data(id INTEGER PRIMARY KEY, file_id INTEGER)")
    con.executemany("INSERT INTO item_data VALUES (?, ?)", zip(data_ids.tolist(), file_of.tolist()))
    with pk.VectorIndex(d, pk.F32) as ix:
        ix.append(x, data_ids)
        ix.seal()
        assert N.lib().pkv_sqlite_register_index(b"clip/test", ix._h) == 0
        try:
            k = 500
            rows = con.execute(
                "WITH dist AS MATERIALIZED (SELECT item_data.file_id AS file_id, t.d AS d "
                "  FROM pkv_topk('clip/test', ?, ?, 'COSINE') AS t JOIN item_data ON item_data.id = t.id) "
                "SELECT file_id, MIN(d) AS agg, row_number() OVER (ORDER BY MIN(d) ASC) AS order_rank "
                "FROM dist GROUP BY file_id ORDER BY order_rank LIMIT 50", (q[0].tobytes(), k)).fetchall()
            raw = con.execute("SELECT id, d, rank FROM pkv_topk('clip/test', ?, 7)", (q[0].tobytes(),)).fetchall()
            assert con.execute("SELECT pkv_last_execute_ms()").fetchone()[0] > 0.0
            with pytest.raises(sqlite3.OperationalError, match="stores 64-dimensional"):
                con.execute("SELECT id FROM pkv_topk('clip/test', ?, 7)", (q[0, :32].tobytes(),)).fetchall()
            with pytest.raises(sqlite3.OperationalError, match="k must be a positive integer"):
                con.execute("SELECT id FROM pkv_topk('clip/test', ?, 0)", (q[0].tobytes(),)).fetchall()
        finally:
            N.lib().pkv_sqlite_register_index(b"clip/test", None)
    want = orc.topk(x, q, orc.COSINE, k, threads=4)
    best = {}
    for r, dist in zip(want[0][0], want[1][0]):
        f = int(file_of[r])
        best[f] = min(best.get(f, np.inf), float(dist))
    ranked = sorted(best.items(), key=lambda kv: (kv[1], kv[0]))[:50]
    assert [r[2] for r in rows] == list(range(1, 51))
    for (f_got, agg_got, _), (f_want, agg_want) in zip(rows, ranked):
        assert abs(agg_got - agg_want) <= 1e-5 * max(abs(agg_want), abs(1 - agg_want)) + 1e-7
    assert [r[0] for r in raw] == data_ids[want[0][0][:7]].tolist() and [r[2] for r in raw] == list(range(1, 8))


def _similar_reference(x, target_rows, grp, n_groups, metric, agg, modality=None, clip_xmodal=False, i2i=True, t2t=True,
                       w=None):
    """NumPy restatement of the self-join of item_similarity.rs:432-581 (float64 aggregates like SQLite)."""
    acc = {}
    for m in target_rows:
        d = orc.distances(x, x[m], metric).astype(np.float64)
        for r in range(len(x)):
            g = grp[r]
            if g < 0:
                continue
            if modality is not None:
                rm, qm = modality[r], modality[m]
                if not clip_xmodal and (rm != 0 or qm != 0):
                    continue
                if clip_xmodal and not i2i and rm == 0 and qm == 0:
                    continue
                if clip_xmodal and not t2t and rm == 1 and qm == 1:
                    continue
            ww = 1.0 if w is None else float(w[r]) * float(w[m])
            acc.setdefault(g, []).append((d[r], ww))
    out = {}
    for g, pairs in acc.items():
        ds = np.array([p[0] for p in pairs]); ws = np.array([p[1] for p in pairs])
        if w is not None:
            out[g] = float((ds * ws).sum() / ws.sum())
        else:
            out[g] = float({pk.AGG_MIN: ds.min(), pk.AGG_MAX: ds.max(), pk.AGG_AVG: ds.mean()}[agg])
    return sorted(out.items(), key=lambda kv: (kv[1], kv[0]))


@pytest.mark.parametrize("flags", [(False, True, True), (True, True, True), (True, False, True), (True, True, False),
                                   (True, False, False)])
def test_similar_to_cross_modal_pair_rules_and_weights(flags):
    """a11/f2: clip_xmodal / xmodal_i2i / xmodal_t2t (item_similarity.rs:468-488) and the confidence weights
    (:523-560) inside pkv_similar_to_device, one index holding the image setter's and the text sibling's rows."""
    import torch

    clip_xmodal, i2i, t2t = flags
    rng = np.random.default_rng(91)
    n, d, n_items = 900, 32, 300
    x = orc.synthetic(n, d, 801)
    item = rng.integers(0, n_items, size=n).astype(np.int64)
    modality = (rng.random(n) < 0.4).astype(np.uint8)            # 40 % text-sibling rows
    w = rng.uniform(0.2, 1.0, size=n).astype(np.float32)
    target_item = int(item[5])
    target_rows = np.nonzero(item == target_item)[0].astype(np.int64)
    grp = np.where(item == target_item, -1, item)
    with pk.VectorIndex(d, pk.F32) as ix:
        ix.append(x)
        ix.seal()
        dev = lambda a: torch.from_numpy(a).cuda()
        for agg, weights in ((pk.AGG_AVG, None), (pk.AGG_MIN, None), (pk.AGG_AVG, w)):
            got_g, got_a, cnt = pk.similar_to(ix, dev(target_rows), dev(grp), n_items, pk.COSINE, agg,
                                              weights=None if weights is None else dev(weights), modality=dev(modality),
                                              clip_xmodal=clip_xmodal, xmodal_i2i=i2i, xmodal_t2t=t2t, limit=40)
            want = _similar_reference(x, target_rows, grp, n_items, orc.COSINE, agg, modality, clip_xmodal, i2i, t2t, weights)
            assert cnt == min(40, len(want))
            got_vals = got_a.cpu().numpy()[:cnt]
            assert np.allclose(got_vals, [v for _, v in want[:cnt]], rtol=2e-5, atol=1e-6)
            assert len(set(got_g.cpu().numpy()[:cnt].tolist()) ^ {g for g, _ in want[:cnt]}) <= 2   # near-tie swaps
        with pytest.raises(pk.PqlError, match="clip_xmodal needs the per-row modality"):
            pk.similar_to(ix, dev(target_rows), dev(grp), n_items, pk.COSINE, clip_xmodal=True)


def test_cross_modal_space_membership_and_auto_fallback():
    """a6/a9: one space = image setter + "t"-sibling rows (db/vector_quants.rs:480-510); without clip_xmodal only
    the image rows are members (image_embeddings.rs:140-199); a quant index whose scale disagrees with the ready pair
    makes non-strict `auto` fall back to exact and strict `quant` fail (preprocess.rs:356-362)."""
    rng = np.random.default_rng(92)
    n, d = 30_000, 64
    x, q, scale, xc, qc = int8_space(n, d, seed=811, nq=5)
    modality = (rng.random(n) < 0.5).astype(np.uint8)
    bm = np.packbits(np.concatenate([modality == 0, np.zeros((-n) % 64, bool)]), bitorder="little").view(np.uint64)
    with pk.VectorIndex(d, pk.F32) as exact, pk.VectorIndex(d, pk.I8) as quant:
        exact.append(x); exact.seal()
        quant.set_scale_artifact(pk.scale_artifact(scale)); quant.append(xc); quant.seal()
        sp = pk.Space("clip/ViT", exact)
        sp.set_modality(modality)
        sp.set_quant("int8-gsym", pk.ReadyPair(7, scale, d), quant)
        ids, dist, cnt, used = sp.search(q, pk.COSINE, pk.INDEX_EXACT, depth=20)                  # image rows only
        assert used == -1 and np.all(modality[ids] == 0)
        assert_close_topk((ids, dist, cnt), orc.topk(x, q, orc.COSINE, 20, bitmap=bm, threads=4), x, q, orc.COSINE)
        ids, dist, cnt, used = sp.search(q, pk.COSINE, pk.INDEX_EXACT, depth=20, clip_xmodal=True)  # the whole space
        assert_close_topk((ids, dist, cnt), orc.topk(x, q, orc.COSINE, 20, threads=4), x, q, orc.COSINE)
        assert np.any(modality[ids] == 1)
        ids, dist, cnt, used = sp.search(q, pk.COSINE, pk.INDEX_AUTO, depth=20)                   # quant, image rows only
        assert used == 7
        assert_exact((ids, dist, cnt), orc.topk(xc, qc, orc.COSINE, 20, bitmap=bm, threads=4))
        # the pair's frozen scale moved on (a rebuild is pending): auto falls back, quant refuses
        sp.set_quant("int8-gsym", pk.ReadyPair(7, scale * 2, d), quant)
        ids, dist, cnt, used = sp.search(q, pk.COSINE, pk.INDEX_AUTO, depth=20, clip_xmodal=True)
        assert used == -1
        with pytest.raises(pk.PqlError, match="frozen scale"):
            sp.search(q, pk.COSINE, pk.INDEX_QUANT, depth=20)
        sp.close()
