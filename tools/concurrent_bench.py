#!/usr/bin/env python3
"""Request-level concurrency: T host threads each issue single-query pkv_search calls (how the Rust
server's read pool would call it), with and without combining.  Prints one JSON line per setting."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import panoptikon_b200 as pk  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--requests", type=int, default=40)
    a = ap.parse_args()
    import torch

    dev = torch.device("cuda", 0)
    code = {"f32": pk.F32, "i8": pk.I8}[a.dtype]
    ix = pk.VectorIndex(a.dim, code)
    ix.reserve(a.rows)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    scale = 0.2 / 127
    if code == pk.I8:
        ix.set_scale_artifact(pk.scale_artifact(scale))
    for b in range(0, a.rows, 250_000):
        n = min(250_000, a.rows - b)
        x = torch.randn((n, a.dim), generator=g, device=dev)
        x /= x.norm(dim=1, keepdim=True)
        ix.append(pk.quantize_int8(x, scale) if code == pk.I8 else x)
    ix.seal()
    q = torch.randn((a.threads, a.dim), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True)
    qh = (pk.quantize_int8(q, scale) if code == pk.I8 else q).cpu().numpy()
    for combine in (0, 1):
        ix.set_option("combine", combine)
        c0 = ix.counters().combined_searches

        def worker(t):
            for _ in range(a.requests):
                ix.search(qh[t:t + 1], 100, pk.COSINE)

        for t in range(2):
            worker(t)  # warm up
        threads = [threading.Thread(target=worker, args=(t,)) for t in range(a.threads)]
        t0 = time.perf_counter()
        [t.start() for t in threads]
        [t.join() for t in threads]
        dt = time.perf_counter() - t0
        total = a.threads * a.requests
        print(json.dumps({"workload": f"{a.rows}x{a.dim} {a.dtype} cosine top-100, {a.threads} threads x 1 query/request",
                          "combine": combine, "requests": total, "queries_per_s": total / dt,
                          "mean_latency_ms": dt / a.requests * 1e3,
                          "combined_searches": int(ix.counters().combined_searches - c0)}), flush=True)


if __name__ == "__main__":
    main()
