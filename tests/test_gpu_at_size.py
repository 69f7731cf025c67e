"""Parity at benchmark size: the kernels bench.py times - the CTA-pair int8-image filter at
batch 256 (f32) and the TMEM-resident-query int8 scan at batch 1024 - checked against the oracle on a 2M x 768
corpus, 64 queries spread over the whole batch (so every query group / CTA of the launch is covered), both on the
default schedule and with the live launch forced."""
import numpy as np
import pytest

import panoptikon_b200 as pk
from oracle import oracle as orc
from tests.helpers import assert_close_topk, assert_exact

pytestmark = pytest.mark.gpu
N, D, K = 2_000_000, 768, 100


@pytest.fixture(scope="module")
def corpus():
    return orc.synthetic(N, D, 0x5EED)


def _spread(nq, n=64):
    return np.unique(np.linspace(0, nq - 1, n).astype(np.int64))


@pytest.mark.parametrize("live", [1, 2])
def test_f32_cosine_batch256_at_2M_rows(corpus, live):
    q = orc.synthetic(256, D, 0x5EED + 1)
    with pk.VectorIndex(D, pk.F32) as ix:
        ix.set_option("live", live)
        ix.append(corpus)
        ix.seal()
        got = ix.search(q, K, pk.COSINE)
        c = ix.counters()
        assert c.last_scan_kind == 8 and c.fallback_queries == 0
        if live == 2:
            assert c.live_refreshes > 0, "the live launch never re-selected a threshold"
    sel = _spread(256)
    want = orc.topk(corpus, q[sel], orc.COSINE, K, threads=64)
    assert_close_topk(tuple(g[sel] for g in got), want, corpus, q[sel], orc.COSINE)


@pytest.mark.parametrize("live", [1, 2])
def test_int8_batch1024_at_2M_rows_bit_exact(corpus, live):
    scale = orc.scale_from_absmax(float(np.abs(corpus).max()))
    xc = orc.quantize_rows(corpus, scale)
    qc = orc.quantize_rows(orc.synthetic(1024, D, 0x5EED + 1), scale)
    with pk.VectorIndex(D, pk.I8) as ix:
        ix.set_option("live", live)
        ix.set_scale_artifact(pk.scale_artifact(scale))
        ix.append(xc)
        ix.seal()
        got = ix.search(qc, K, pk.COSINE)
        dot = ix.search(qc, K, pk.DOT)
        assert ix.counters().last_scan_kind == 3 and ix.counters().fallback_queries == 0
    sel = _spread(1024)
    assert_exact(tuple(g[sel] for g in got), orc.topk(xc, qc[sel], orc.COSINE, K, threads=64))
    assert_exact(tuple(g[sel] for g in dot), orc.topk(xc, qc[sel], orc.DOT, K, threads=64))
