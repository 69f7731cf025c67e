"""CPU model of the int8-image filter bound (panoptikon_b200/csrc/pkv_scan_img8.cu): a NumPy float32 restatement
of img8_build_kernel / img8_prep_queries_kernel / pair_bound, checked against exact float64 arithmetic.

The property the CUDA filter relies on: a (row, query) pair whose exact distance is within the k-th best distance is
NEVER below the bound, for every metric, including saturating (peaky) rows, unnormalised rows and tiny thresholds.
The GPU tests check the kernels end to end; this one pins the algebra (in particular the product form of the L2
bound) without a GPU."""
import numpy as np
import pytest

from oracle import oracle as orc

F = np.float32
L2, COSINE, DOT = orc.L2, orc.COSINE, orc.DOT


def build_image(x, peak_sigma=3.0):
    """img8_stats_kernel + img8_build_kernel: direction quantisation with code length U for every row."""
    x = x.astype(F)
    nrm2 = (x * x).sum(axis=1, dtype=F)
    amax = np.abs(x).max(axis=1)
    ok = (nrm2 > 0) & np.isfinite(nrm2) & np.isfinite(amax)
    p = amax[ok].astype(np.float64) / np.sqrt(nrm2[ok].astype(np.float64))
    p_ref = min(max(p.mean() + peak_sigma * p.std(), 1e-3), 1.0) if ok.any() else 1.0
    U = F(127.0 / p_ref)
    s = np.where(ok, np.sqrt(nrm2) / U, F(1.0)).astype(F)
    s[~(s > 0)] = F(1.0)
    r = (F(1.0) / s).astype(F)
    codes = np.clip(np.rint(x / s[:, None]), -127, 127).astype(F)
    err = (x - s[:, None] * codes).astype(F)
    meta = np.stack([np.sqrt(nrm2) * r,                                   # u = |a| / s_a
                     np.sqrt((codes * codes).sum(axis=1, dtype=F)),       # v = |c|
                     np.sqrt((err * err).sum(axis=1, dtype=F)) * F(1.0001) * r,   # w = |a - s_a c| / s_a
                     r], axis=1).astype(F)
    return codes.astype(np.int32), meta


def prep_queries(q):
    q = q.astype(F)
    amax = np.abs(q).max(axis=1)
    s = (amax / F(127.0)).astype(F)
    s[~(s > 0)] = F(1.0)
    r = (F(1.0) / s).astype(F)
    codes = np.clip(np.rint(q / s[:, None]), -127, 127).astype(F)
    err = (q - s[:, None] * codes).astype(F)
    nrm2 = (q * q).sum(axis=1, dtype=F)
    meta = np.stack([r, np.sqrt((err * err).sum(axis=1, dtype=F)) * F(1.0001) * r,
                     np.sqrt(nrm2) * F(1.00001) * r, nrm2], axis=1).astype(F)
    return codes.astype(np.int32), meta


def pair_bound(metric, thr_f, qm, rm):
    """pair_bound<METRIC> for one query (qm: its 4 figures) against every row (rm: [n, 4]), lo == hi per row."""
    iq, ty, tz, q2 = (F(v) for v in qm)
    u, v, w, r = (rm[:, i] for i in range(4))
    if metric == L2:
        c1, c2 = iq, F(0.5) * (q2 - F(thr_f))
        qs = F(1e-4) * F(0.5) * iq * (q2 + abs(F(thr_f)))
        na = u / r
        x1 = F(0.5) * na * na
        summ = x1 + c2
        lead = c1 * r * summ
        mag = c1 * r * (x1 + abs(c2))
    else:
        c1 = -F(thr_f) * iq
        qs = F(0.0)
        x1 = u if metric == COSINE else r
        lead = c1 * x1
        mag = abs(c1) * x1
    neg = ty * v + tz * w
    return (lead - neg - (F(3e-5) * (mag + neg) + qs * r + F(4e-6) * tz * u + F(1.0))).astype(F)


def filter_threshold(metric, d_k, q2):
    """select_kernel's filter_threshold for the FilterSpec of the int8-image path (abs = 0)."""
    if metric == DOT:
        return F(d_k)
    if metric == COSINE:
        T = (1.0 - float(d_k) - 2.4e-7) * np.sqrt(float(q2))
        return F(np.nextafter(F(-T + abs(T) * 1e-5 + 1e-30), F(np.inf)))
    return F(np.nextafter(F(float(d_k) ** 2 * (1.0 + 1e-6) + 1e-30), F(np.inf)))


def corpora():
    rng = np.random.default_rng(5)
    unit = orc.synthetic(4000, 256, 901)
    spread = orc.synthetic(4000, 96, 902, normalise=False) * rng.uniform(0.01, 30.0, size=(4000, 1)).astype(F)
    peaky = orc.synthetic(4000, 128, 903) * F(0.02)
    peaky[np.arange(4000), rng.integers(0, 128, 4000)] += F(1.0)
    peaky[:50] = orc.synthetic(50, 128, 904)
    return {"unit": unit, "spread": spread, "peaky": peaky}


@pytest.mark.parametrize("name", ["unit", "spread", "peaky"])
@pytest.mark.parametrize("metric", [L2, COSINE, DOT])
def test_filter_never_drops_a_top_k_pair(name, metric):
    x = corpora()[name]
    q = orc.synthetic(24, x.shape[1], 905, normalise=(name == "unit"))
    if name == "peaky":
        q[:8] = x[:8] * F(2.0)
    codes, rmeta = build_image(x)
    qcodes, qmeta = prep_queries(q)
    acc = codes @ qcodes.T                                     # exact s32 dot products of the codes
    x64, q64 = x.astype(np.float64), q.astype(np.float64)
    dots = x64 @ q64.T
    if metric == DOT:
        dist = -dots
    elif metric == COSINE:
        dist = 1.0 - dots / (np.linalg.norm(x64, axis=1)[:, None] * np.linalg.norm(q64, axis=1)[None, :])
    else:
        dist = np.sqrt(((x64[:, None, :] - q64[None, :, :]) ** 2).sum(axis=2))
    kept_total = 0
    for k in (1, 10, 200):
        for j in range(q.shape[0]):
            d_sorted = np.sort(dist[:, j])
            d_k = F(d_sorted[k - 1])                            # exact k-th best distance, rounded like a key
            thr = filter_threshold(metric, d_k, qmeta[j, 3])
            bound = pair_bound(metric, thr, qmeta[j], rmeta)
            keep = ~(acc[:, j].astype(F) < bound)
            members = dist[:, j] <= float(d_k)
            assert np.all(keep[members]), (name, metric, k, j, int((~keep[members]).sum()))
            kept_total += int(keep.sum())
    # the filter must also be a filter: on the well-conditioned corpus it passes a small multiple of k
    if name == "unit":
        assert kept_total < 30 * (1 + 10 + 200) * q.shape[0]


def test_no_threshold_keeps_everything_and_padding_is_harmless():
    x = corpora()["unit"][:512]
    codes, rmeta = build_image(x)
    qcodes, qmeta = prep_queries(orc.synthetic(3, 256, 906))
    acc = codes @ qcodes.T
    for metric in (L2, COSINE, DOT):
        with np.errstate(invalid="ignore", over="ignore"):
            bound = pair_bound(metric, F(np.inf), qmeta[0], rmeta)   # thr_f = +inf: the query has no k-th best yet
        assert np.all(~(acc[:, 0].astype(F) < bound))
