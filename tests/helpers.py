"""Shared helpers for the parity tests."""
import numpy as np

from oracle import oracle as orc

F32_RTOL = 1e-5  # north_star: cosine scores within 1e-5 relative on fp32


def assert_exact(got, want):
    """Bit-exact parity (int8 path): ids, distances (NaN == NaN) and counts."""
    ids_g, dist_g, cnt_g = got
    rows_o, dist_o, cnt_o = want
    assert np.array_equal(np.asarray(cnt_g), cnt_o), (cnt_g, cnt_o)
    assert np.array_equal(np.asarray(ids_g), rows_o), _first_diff(ids_g, rows_o, dist_g, dist_o)
    # bit-exact distances; every NaN (SQL NULL) counts as the same value whatever its sign/payload bits
    bg, bo = _canon_bits(np.asarray(dist_g)), _canon_bits(dist_o)
    assert np.array_equal(bg, bo), _first_diff(bg, bo, ids_g, rows_o)


def _canon_bits(d):
    d = np.ascontiguousarray(d, dtype=np.float32)
    return np.where(np.isnan(d), np.uint32(0x7FC00000), d.view(np.uint32))


def _first_diff(a, b, c, d):
    a, b = np.asarray(a), np.asarray(b)
    idx = np.argwhere(a != b)
    if len(idx) == 0:
        return "no diff"
    i = tuple(idx[0])
    return f"first difference at {i}: got {a[i]} want {b[i]} (aux got {np.asarray(c)[i]} want {np.asarray(d)[i]}); {len(idx)} differing entries"


def _tol(d, metric, rtol, atol):
    """Allowed |difference| for oracle distance(s) d.  Cosine: the *score* is the cosine similarity
    1 - d, so 1e-5 relative on the score is rtol * max(|d|, |1 - d|) on the distance (for
    near-duplicate vectors d ~ 1e-6 is a cancellation of O(1) sums and no summation order can hold
    1e-5 relative on d itself).  L2 / DOT: relative on the distance."""
    d = np.abs(d)
    if metric == orc.COSINE:
        return rtol * np.maximum(d, np.abs(1.0 - d)) + atol
    return rtol * d + atol


def assert_close_topk(got, want, corpus, queries, metric, rtol=F32_RTOL, atol=1e-7):
    """f32 parity: the GPU sums in a different order than the scalar loop, so distances agree
    to `rtol` and ids may differ only where the oracle's own distances are that close."""
    ids_g, dist_g, cnt_g = (np.asarray(x) for x in got)
    rows_o, dist_o, cnt_o = want
    assert np.array_equal(cnt_g, cnt_o), (cnt_g, cnt_o)
    nq, k = rows_o.shape
    for q in range(nq):
        m = cnt_o[q]
        dg, do = dist_g[q, :m], dist_o[q, :m]
        nan_g, nan_o = np.isnan(dg), np.isnan(do)
        assert np.array_equal(nan_g, nan_o), f"query {q}: NaN placement differs"
        ok = ~nan_o
        tol = _tol(do[ok], metric, rtol, atol)
        with np.errstate(invalid="ignore"):
            close = (np.abs(dg[ok] - do[ok]) <= tol) | (dg[ok] == do[ok])   # equal infinities are equal
        if not np.all(close):
            i = int(np.argmin(close))
            raise AssertionError(f"query {q}: position {i} of the valid entries: got d={dg[ok][i]!r} (id {ids_g[q, :m][ok][i]}) "
                                 f"want d={do[ok][i]!r} (row {rows_o[q, :m][ok][i]}); {int((~close).sum())} entries differ")
        assert np.all(ids_g[q, m:] == -1)
        if np.array_equal(ids_g[q, :m], rows_o[q, :m]):
            continue
        # every row the GPU returned must carry (to rtol) the distance the oracle gives that row,
        # and rows only one side returned must sit within rtol of the k-th distance
        d_rows = orc.distances(corpus[ids_g[q, :m]], queries[q], metric)
        fin = ~np.isnan(d_rows)
        with np.errstate(invalid="ignore"):
            same = (np.abs(d_rows[fin] - dg[fin]) <= _tol(d_rows[fin], metric, rtol, atol)) | (d_rows[fin] == dg[fin])
        assert np.all(same), f"query {q}: row distance mismatch"
        kth = do[ok][-1] if ok.any() else 0.0
        only_g = np.setdiff1d(ids_g[q, :m], rows_o[q, :m])
        only_o = np.setdiff1d(rows_o[q, :m], ids_g[q, :m])
        assert len(only_g) == len(only_o)
        for r in np.concatenate([only_g, only_o]):
            d = orc.distances(corpus[r:r + 1], queries[q], metric)[0]
            assert abs(d - kth) <= 2 * _tol(kth, metric, rtol, atol), f"query {q}: row {r} swapped across a gap {abs(d-kth)}"
        # rows both sides returned may be permuted only within near-ties
        pos_o = {r: i for i, r in enumerate(rows_o[q, :m])}
        for i, r in enumerate(ids_g[q, :m]):
            j = pos_o.get(r)
            if j is not None and j != i:
                assert abs(do[j] - do[i]) <= 2 * _tol(do[i], metric, rtol, atol) or (np.isnan(do[j]) and np.isnan(do[i]))


def int8_space(n, d, seed=orc.CORPUS_SEED, nq=4):
    """Synthetic corpus + queries quantised with the corpus' absmax scale."""
    x = orc.synthetic(n, d, seed)
    q = orc.synthetic(nq, d, seed + 1)
    scale = orc.scale_from_absmax(float(np.abs(x).max())) if n else 1.0
    return x, q, scale, orc.quantize_rows(x, scale), orc.quantize_rows(q, scale)
