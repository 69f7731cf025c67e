#!/usr/bin/env python3
"""CPU model of how many rows the int8-image filter passes to the exact re-scorer (NumPy restatement of
pkv_scan_img8.cu, shared with tests/test_img8_bound_model.py): for a unit-norm synthetic corpus it replays the search
driver's chunk schedule (thresholds = exact k-th best of the rows seen so far) and counts, per query, the pairs whose
int8 estimate is not below the bound, next to the number of pairs that really beat the threshold.  No GPU needed."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from tests.test_img8_bound_model import build_image, filter_threshold, pair_bound, prep_queries  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=400_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=16)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--growth", type=float, default=1.5)
    ap.add_argument("--sigma", type=float, default=3.0)
    a = ap.parse_args()
    x = orc.synthetic(a.rows, a.dim, 0x5EED)
    q = orc.synthetic(a.queries, a.dim, 0x5EED + 1)
    codes, rmeta = build_image(x, peak_sigma=a.sigma)
    qcodes, qmeta = prep_queries(q)
    acc = (codes.astype(np.float32) @ qcodes.astype(np.float32).T)          # exact: |acc| < 2^24
    cos = 1.0 - (x.astype(np.float64) @ q.astype(np.float64).T)             # unit norm: cosine distance
    first = 3968
    pos, chunks = first, []
    while pos < a.rows:
        c = min(int(pos * a.growth) // 128 * 128 or 128, a.rows - pos)
        chunks.append((pos, pos + c))
        pos += c
    passed = np.zeros(a.queries)
    true_new = np.zeros(a.queries)
    print(f"{a.rows} x {a.dim}, k={a.k}, growth x{a.growth}, saturation quantile mean+{a.sigma} sigma; {len(chunks)} tensor chunks")
    for b, e in chunks:
        p_chunk = np.zeros(a.queries)
        for j in range(a.queries):
            d_k = np.float32(np.partition(cos[:b, j], a.k - 1)[a.k - 1])
            thr = filter_threshold(orc.COSINE, d_k, qmeta[j, 3])
            bound = pair_bound(orc.COSINE, thr, qmeta[j], rmeta[b:e])
            p_chunk[j] = np.count_nonzero(~(acc[b:e, j] < bound))
            true_new[j] += np.count_nonzero(cos[b:e, j] <= d_k)
        passed += p_chunk
        print(f"  rows [{b:>8}, {e:>8}): passed to the re-scorer per query: mean {p_chunk.mean():7.1f}")
    print(f"total per query: passed {passed.mean():.0f}, really below the stale threshold {true_new.mean():.0f} "
          f"-> {passed.mean() / max(true_new.mean(), 1):.1f} rows re-scored per true candidate")


if __name__ == "__main__":
    main()
