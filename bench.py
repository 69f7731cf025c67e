#!/usr/bin/env python3
"""bench.py — queries/sec of the vector-similarity hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic queries: B queries against
the whole N x D corpus (resident in HBM), exact top-k per query.  With --gpus G the corpus is
row-sharded over G ranks (one process per GPU, torchrun), every rank scans its shard for the
same batch, one NCCL all-gather collects the per-shard top-k and a merge kernel produces the
global top-k (strong scaling: the corpus is fixed, per-GPU work shrinks).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--rows N] [--dim D] [--batch B] [--k K] [--dtype f32|i8|f16] [--metric cosine|l2|dot]

Prints ONE JSON line (see DESIGN.md "Measurement" for every key).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CORPUS_SEED = 0x5EED
BLOCK_ROWS = 250_000  # synthetic corpus is generated in blocks seeded by block index


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--dtype", default="f32", choices=["f32", "i8", "f16"])
    ap.add_argument("--metric", default="cosine", choices=["cosine", "l2", "dot"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--bitmap-density", type=float, default=0.0,
                    help="BASELINE config 5: AND the scan with a tag-filter row bitmap of this density (shared by the batch)")
    ap.add_argument("--force-simt", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="index option name=value (tuning experiments)")
    return ap.parse_args()


def workload_name(a) -> str:
    n = f"{a.rows // 1_000_000}M" if a.rows % 1_000_000 == 0 else str(a.rows)
    tag = f" AND tag bitmap (density {a.bitmap_density:g})" if a.bitmap_density > 0 else ""
    return f"{n}x{a.dim} {a.dtype} {a.metric} top-{a.k}, batch={a.batch} queries{tag}"


# ------------------------------------------------------------------ synthetic data
def gen_block(torch, block: int, rows: int, dim: int, device):
    """Rows [block*BLOCK_ROWS, +rows): standard normal, L2-normalised (SURVEY §8d recipe), a pure
    function of (seed, block) so every rank/shard layout sees the same global corpus."""
    g = torch.Generator(device=device)
    g.manual_seed(CORPUS_SEED * 1_000_003 + block)
    x = torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)
    x /= x.norm(dim=1, keepdim=True)
    return x


def gen_queries(torch, nq: int, dim: int, device, step: int = 0):
    g = torch.Generator(device=device)
    g.manual_seed((CORPUS_SEED + 1) * 1_000_003 + step)
    q = torch.randn((nq, dim), generator=g, device=device, dtype=torch.float32)
    q /= q.norm(dim=1, keepdim=True)
    return q


def corpus_blocks(row_begin: int, row_end: int):
    """(block index, offset in block, rows) pieces covering [row_begin, row_end)."""
    r = row_begin
    while r < row_end:
        b = r // BLOCK_ROWS
        off = r - b * BLOCK_ROWS
        n = min(BLOCK_ROWS - off, row_end - r)
        yield b, off, n
        r += n


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip().splitlines()
                if out:
                    self.samples.append([c.strip() for c in out[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline (oracle)
def cpu_baseline(a, corpus_sample: np.ndarray, queries: np.ndarray, metric_code: int, seconds: float):
    """Times the oracle's restatement of the reference scan (one query per thread, like one SQLite
    connection per thread) on a bounded sample: T queries x R rows; QPS at the full N rows is
    extrapolated linearly (the scan is linear in rows)."""
    from oracle import oracle as orc

    cores = os.cpu_count() or 1
    threads = min(cores, queries.shape[0])
    q = np.ascontiguousarray(queries[:threads])
    probe = min(20_000, corpus_sample.shape[0])
    t0 = time.perf_counter()
    orc.topk(corpus_sample[:probe], q, metric_code, a.k, threads=threads)
    rate = probe / max(time.perf_counter() - t0, 1e-6)  # rows/s per thread (all threads in parallel)
    rows = int(min(corpus_sample.shape[0], max(probe, rate * seconds)))
    t0 = time.perf_counter()
    out = orc.topk(corpus_sample[:rows], q, metric_code, a.k, threads=threads)
    dt = time.perf_counter() - t0
    qps_full = (threads * rows / dt) / a.rows
    return {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{threads} queries (1 per thread) x first {rows} rows in {dt:.2f} s, "
                      f"linearly extrapolated to {a.rows} rows",
            "rows_per_s_per_core": rows / dt}, rows, out


def reference_arm(a):
    """--impl reference: the reference's CPU algorithm (oracle port: the Rust/sqlite-vec path cannot be
    built here) on the box's host cores, same config/metric/unit, each step a bounded sample."""
    from oracle import oracle as orc

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc.build()
    metric_code = {"l2": orc.L2, "cosine": orc.COSINE, "dot": orc.DOT}[a.metric]
    cores = os.cpu_count() or 1
    # the sample: `cores` queries x R rows of the same synthetic recipe, generated on the host
    rate_guess = 0.6e6 if a.dtype != "i8" else 0.8e6
    per_step_s = max(0.25, min(2.0, 120.0 / max(a.steps + a.warmup, 1)))
    rows = int(min(a.rows, 1_000_000, max(20_000, rate_guess * per_step_s)))
    x = orc.synthetic(rows, a.dim, CORPUS_SEED)
    q = orc.synthetic(cores, a.dim, CORPUS_SEED + 1)
    if a.dtype == "i8":
        s = orc.scale_from_absmax(float(np.abs(x).max()))
        x, q = orc.quantize_rows(x, s), orc.quantize_rows(q, s)
    elif a.dtype == "f16":
        x, q = x.astype(np.float16), q.astype(np.float16)
    times = []
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        orc.topk(x, q, metric_code, a.k, threads=cores)
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    qps = (cores * rows / dt) / a.rows
    sample = (f"each step: {cores} queries (1 per thread) x {rows} rows of the same synthetic recipe; "
              f"queries/s extrapolated linearly to {a.rows} rows")
    line = {
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k,
                   "metric": a.metric},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ our arm
def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
        return

    import torch
    import torch.distributed as dist

    import panoptikon_b200 as pk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: panoptikon_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, f"--gpus {a.gpus} but WORLD_SIZE={world}"

    dtype_code = {"f32": pk.F32, "i8": pk.I8, "f16": pk.F16}[a.dtype]
    metric_code = {"l2": pk.L2, "cosine": pk.COSINE, "dot": pk.DOT}[a.metric]
    elem = {"f32": 4, "i8": 1, "f16": 2}[a.dtype]

    # ---- corpus shard of this rank, generated on the device and appended through the C ABI
    per = (a.rows + world - 1) // world
    r0, r1 = min(rank * per, a.rows), min((rank + 1) * per, a.rows)
    ix = pk.VectorIndex(a.dim, dtype_code, device=local_rank)
    if a.force_simt:
        ix.set_option("force_simt", 1)
    for kv in a.opt:   # before the first allocation: image_mask decides which filter images exist
        name, value = kv.split("=")
        ix.set_option(name, int(value))
    ix.reserve(max(r1 - r0, 1))
    ix.set_row_base(r0)
    scale = None
    if a.dtype == "i8":
        # global symmetric absmax scale over the whole corpus (docs/vector-int8-quant.md:11-49)
        absmax = 0.0
        for b, off, n in corpus_blocks(r0, r1):
            absmax = max(absmax, pk.blob_absmax(gen_block(torch, b, BLOCK_ROWS, a.dim, dev)[off:off + n].contiguous(),
                                                device=local_rank))
        t = torch.tensor([absmax], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        scale = pk.scale_from_absmax(float(t.item()))
        ix.set_scale_artifact(pk.scale_artifact(scale))
    sample_rows = min(r1 - r0, 2_000_000) if rank == 0 else 0
    sample_host = []
    for b, off, n in corpus_blocks(r0, r1):
        x = gen_block(torch, b, BLOCK_ROWS, a.dim, dev)[off:off + n].contiguous()
        if a.dtype == "i8":
            x = pk.quantize_int8(x, scale, device=local_rank)
        elif a.dtype == "f16":
            x = x.half()
        ix.append(x)
        have = sum(s.shape[0] for s in sample_host)
        if have < sample_rows:
            sample_host.append(x[: sample_rows - have].cpu().numpy())
        del x
    ix.seal()
    torch.cuda.synchronize()

    def make_queries(step):
        q = gen_queries(torch, a.batch, a.dim, dev, step)
        if a.dtype == "i8":
            return pk.quantize_int8(q, scale, device=local_rank)
        if a.dtype == "f16":
            return q.half()
        return q

    n_q_sets = 4
    q_dev = [make_queries(s) for s in range(n_q_sets)]
    q_pinned = [q.cpu().pin_memory() for q in q_dev]
    out_dev = (torch.empty((a.batch, a.k), dtype=torch.int64, device=dev),
               torch.empty((a.batch, a.k), dtype=torch.float32, device=dev),
               torch.empty(a.batch, dtype=torch.int32, device=dev))
    comm = None
    if world > 1:
        merged_out = (torch.empty((a.batch, a.k), dtype=torch.int64, device=dev),
                      torch.empty((a.batch, a.k), dtype=torch.float32, device=dev),
                      torch.empty(a.batch, dtype=torch.int32, device=dev))
        # the exchange lives in libpkv.so (pkv_comm_* / pkv_search_sharded_device: scan + pack + ONE ncclAllGather +
        # merge on one stream); torch.distributed only carries the 128-byte NCCL id and the timing barriers
        uid = torch.tensor(list(pk.Comm.unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device=dev)
        dist.broadcast(uid, 0)
        comm = pk.Comm(local_rank, rank, world, bytes(uid.cpu().numpy().tolist()))

    merge_launches = [0]
    # config 5: membership bits over this shard's rows (Bernoulli(p), seed 0x5EED+2 over GLOBAL rows: SURVEY 8d),
    # packed LSB-first into u64 words; the same bitmap for every query of the batch (context of an AND filter)
    bm_dev = bm_host = None
    if a.bitmap_density > 0:
        member = np.random.default_rng(CORPUS_SEED + 2).random(a.rows) < a.bitmap_density
        local = np.zeros(((r1 - r0 + 63) // 64) * 64, dtype=bool)
        local[: r1 - r0] = member[r0:r1]
        bm_host = np.packbits(local, bitorder="little").view(np.uint64).copy()
        bm_dev = torch.from_numpy(bm_host.view(np.int64)).to(dev)

    def step_device(q):
        if world == 1:
            return ix.search(q, a.k, metric_code, out=out_dev, bitmap=bm_dev)
        # ONE all-gather of the per-shard candidates: a pack kernel (12-byte entries), the collective, and a merge
        # kernel that reads the gathered buffer directly - all inside the library call
        merge_launches[0] += 2
        return comm.search(ix, q, a.k, metric_code, bitmap=bm_dev, out=merged_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (inputs already in HBM): `value`
    for i in range(max(a.warmup, 3)):
        step_device(q_dev[i % n_q_sets])
    barrier()
    c0 = ix.counters()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    scan_ms = 0.0
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(a.steps):
        step_device(q_dev[i % n_q_sets])
        scan_ms += ix.counters().last_scan_ms
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    c1 = ix.counters()
    t = torch.tensor([ms, scan_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, scan_ms_max = float(t[0].item()), float(t[1].item())
    ms_per_step = ms_total / a.steps
    value = a.batch / (ms_per_step * 1e-3)

    # ---- end to end through the public API with HOST buffers: `e2e`
    h2d = a.batch * a.dim * elem + (bm_host.nbytes if bm_host is not None else 0)
    d2h = a.batch * a.k * 12 + a.batch * 4
    out_host = (np.empty((a.batch, a.k), np.int64), np.empty((a.batch, a.k), np.float32), np.empty(a.batch, np.int32))
    out_pinned = [torch.from_numpy(o).pin_memory() for o in out_host]
    out_pinned_np = tuple(o.numpy() for o in out_pinned)

    def step_e2e(qp):
        if world == 1:
            # the C-ABI host call: H2D of the queries, scan, D2H of ids/dist/counts, all inside
            return ix.search(qp.numpy(), a.k, metric_code, out=out_pinned_np, bitmap=bm_host)
        q = qp.to(dev, non_blocking=True)
        ids, dst, cnt = step_device(q)
        return ids.cpu(), dst.cpu(), cnt.cpu()

    for i in range(3):
        step_e2e(q_pinned[i % n_q_sets])
    barrier()
    t0 = time.perf_counter()
    for i in range(a.steps):
        step_e2e(q_pinned[i % n_q_sets])
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_qps = a.batch / (float(t.item()) / a.steps)

    if rank != 0:
        if world > 1:
            comm.close()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (scan) kernel, from the library's own CUDA-event timing of its
    # scan launches on the stream they run on; algorithmic bytes/ops as DESIGN.md states them
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    scan_launches = (c1.scan_launches - c0.scan_launches) / a.steps
    scan_ms_step = scan_ms_max / a.steps
    shard_rows = r1 - r0
    kind = c1.last_scan_kind
    # bytes one pass of the dominant kernel must read per row: the image it scans (+4 B row norm on
    # the paths that read it).  kind 7 scans the index's fp16 image of the f32 rows.
    pad = lambda n, m: (n + m - 1) // m * m
    row_bytes = {1: a.dim * 4, 2: a.dim + 4, 3: a.dim + 4, 4: a.dim * 4 + 4, 5: a.dim * 2, 6: a.dim * 2 + 4,
                 7: pad(a.dim, 64) * 2 + 4, 8: pad(a.dim, 128) + 16}.get(kind, a.dim * elem)
    alg_bytes = shard_rows * row_bytes
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    achieved_gbs = alg_bytes / (scan_ms_step * 1e-3) / 1e9 if scan_ms_step > 0 else 0.0
    ops = 2.0 * a.batch * shard_rows * a.dim
    tput = ops / (scan_ms_step * 1e-3) / 1e12 if scan_ms_step > 0 else 0.0
    # Which roof binds: passes of <= 128 queries stream the scanned image at HBM speed (ncu: ~85 % DRAM
    # throughput, tensor pipe ~40 % active); 256-query passes keep the tensor pipe ~77 % active with DRAM
    # at ~53 % (profiles/r01_ncu_*), i.e. they are tensor/smem-bound.
    tensor_bound = kind in (3, 4, 6, 7, 8) and a.batch > 128
    int8_pipe = a.dtype == "i8" or kind == 8
    bf16_peak = peaks.get("bf16_tflops", 1590.0)   # burst figure: the timed region is a fraction of a second
    tensor_peak = 2.0 * bf16_peak if int8_pipe else (0.5 * bf16_peak if kind == 4 else bf16_peak)
    roofline = {
        "bound": "tensor" if tensor_bound else "hbm",
        "achieved": tput if tensor_bound else achieved_gbs,
        "peak": tensor_peak if tensor_bound else hbm_peak,
        "unit": ("TOP/s" if int8_pipe else "TFLOP/s") if tensor_bound else "GB/s",
        "frac": None, "traffic": None,
        "kernel": {1: "scan_f32_simt", 2: "scan_i8_simt", 3: "scan_i8_ts (tcgen05 kind::i8, queries resident in TMEM)" if (a.batch > 128 and not any(o.startswith("tc_ts=0") for o in a.opt)) else "scan_i8_tc (tcgen05 kind::i8)",
                   4: "scan_float_tc (tcgen05 kind::tf32 filter on f32 rows + exact rescore)", 5: "scan_f16_simt",
                   6: "scan_float_tc (tcgen05 kind::f16 on f16 rows + exact rescore)",
                   7: "scan_float_tc (tcgen05 kind::f16 filter on the fp16 image of the f32 rows + exact rescore)",
                   8: "scan_img8 (tcgen05 kind::i8 filter on the per-row-scaled int8 image of the rows, queries in TMEM, + exact rescore)"
                   }.get(kind, str(kind)),
        "algorithmic_bytes_per_row": row_bytes,
        "f32_equivalent_gbs": (shard_rows * a.dim * 4 / (scan_ms_step * 1e-3) / 1e9) if (kind in (7, 8) and a.dtype == "f32" and scan_ms_step > 0) else None,
        # SURVEY 8(d) defines the fp32 roofline on N*D*4 bytes per pass: the same scan time against those bytes
        "f32_hbm_roofline_frac": (shard_rows * a.dim * 4 / (scan_ms_step * 1e-3) / 1e9 / hbm_peak) if (a.dtype == "f32" and scan_ms_step > 0 and hbm_peak) else None,
        "launches_per_step": scan_launches, "kernel_ms_per_step": scan_ms_step,
        "algorithmic_bytes_per_step": alg_bytes, "algorithmic_ops_per_step": ops,
        "achieved_gbs": achieved_gbs, "achieved_tops": tput,
        "hbm_frac": achieved_gbs / hbm_peak if hbm_peak else None,
        "tensor_frac": tput / tensor_peak if tensor_peak else None,
        "peak_source": peak_src + ("; tensor peak = measured cuBLAS bf16 burst" +
                                   (" x2 (the int8 pipe runs at twice the bf16 rate; nominal 4500)" if int8_pipe else "")
                                   if tensor_bound else ""),
    }
    # DRAM traffic: ncu --set full on the main-chunk launches measured dram__bytes_read+write = 1.0005x
    # (scan_float_tc2, 11.587 GB vs 11.581 GB algorithmic) and 1.0024x (scan_i8_tc) of the bytes of the rows
    # the launch covers (profiles/r01_ncu_*.txt): no re-reads inside a pass.  A batch wider than one query
    # tile makes one pass per tile, and every pass streams the image again.
    sub = min(a.batch, 1024)
    tiles = (-(-sub // 256) if sub > 128 else 1) if kind in (3, 4, 6, 7, 8) else -(-sub // 8)
    ratio, note = 1.002, "the ratio ncu measured on the main-chunk launches (profiles/r01_ncu_*.txt)"
    if kind in (3, 8) and sub > 128 and not any(o.startswith("tc_ts=0") for o in a.opt):
        # pkv_scan_ts.cu: one launch serves up to 4 groups of 256 queries; the groups share row tiles through L2.
        # ncu on the main-chunk launch: dram bytes = 1.06x (2 groups) / 1.94x (4 groups) of the rows' bytes
        # (profiles/r01_ncu_i8_b1024_scan_i8_ts.txt) instead of 2x / 4x for separate passes.
        groups = min(tiles, 4)
        tiles = 1
        ratio = {1: 1.002, 2: 1.06, 3: 1.5, 4: 1.94}[groups]
        note = f"{groups} query groups per launch share row tiles through L2: ncu measured {ratio}x the rows' bytes per launch"
    roofline["passes_per_step"] = tiles * -(-a.batch // 1024)
    roofline["traffic"] = roofline["passes_per_step"] * alg_bytes * ratio
    roofline["traffic_note"] = "estimated: passes_per_step x algorithmic bytes x " + note
    roofline["frac"] = roofline["achieved"] / roofline["peak"] if roofline["peak"] else None

    line = {
        "metric": "queries/sec", "value": value, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": max(a.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
        "config": {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k,
                   "metric": a.metric, "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
                   "l2_policy": f"scanned image {alg_bytes / 1e9:.2f} GB per GPU streams through L2 every step"
                                + (" (larger than the 126 MB L2)" if alg_bytes > 200e6 else ""),
                   "query_sets": n_q_sets},
        "clocks": clocks,
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "pkv_search (C ABI, host buffers)" if world == 1 else
                       "H2D + pkv_search_sharded_device (scan + pack + ncclAllGather + merge in libpkv.so) + D2H"},
        "gpu_launches": int(c1.kernel_launches - c0.kernel_launches) + merge_launches[0],
        "roofline": roofline,
        "overflow_rescans": int(c1.fallback_queries - c0.fallback_queries),
        # what the in-kernel machinery did per step (device counters of the searches in the timed region)
        "search_stats": {"live_refreshes_per_step": (c1.live_refreshes - c0.live_refreshes) / a.steps,
                         "live_refresh_skips_per_step": (c1.live_refresh_skips - c0.live_refresh_skips) / a.steps,
                         "rescored_rows_per_query": (c1.rescored_pairs - c0.rescored_pairs) / a.steps / a.batch,
                         "deferred_rows_per_query": (c1.deferred_pairs - c0.deferred_pairs) / a.steps / a.batch},
    }

    # ---- full-size properties of the last batch's result (the oracle cannot scan 10M rows in the time budget):
    # sorted, unique, idempotent, and a sampled exactness certificate - the distance of every returned row and of
    # 4096 random rows of this shard is recomputed from the STORED rows in float64 with torch; returned distances
    # must match and no sampled row outside the result may beat the k-th best.
    if world == 1:
        try:
            qd = q_dev[(a.steps - 1) % n_q_sets]
            ids, dst, cnt = (t.clone() for t in ix.search(qd, a.k, metric_code, bitmap=bm_dev))
            ids2, dst2, _ = ix.search(qd, a.k, metric_code, bitmap=bm_dev)
            full = bool((cnt == a.k).all().item())
            props = {"sorted": bool((dst[:, 1:] >= dst[:, :-1]).all().item()) if full else None,
                     "unique_ids": bool(all(len(set(r)) == len(r) for r in ids[:8].cpu().tolist())),
                     "idempotent": bool(torch.equal(ids, ids2) and torch.equal(dst.view(torch.int32), dst2.view(torch.int32)))}
            g = torch.Generator(device=dev)
            g.manual_seed(12345)
            sample_rows = torch.randint(0, r1 - r0, (4096,), generator=g, device=dev)
            nqc = min(a.batch, 64)                                   # certificate on the first 64 queries

            def exact(rows_idx, qq):
                xr = ix.get_rows(rows_idx).to(torch.float64)
                if a.dtype == "i8":
                    qq = qq.to(torch.float64)
                else:
                    qq = qq.to(torch.float64)
                dots = qq @ xr.T
                if a.metric == "dot":
                    return -dots
                if a.metric == "cosine":
                    return 1.0 - dots / (qq.norm(dim=1, keepdim=True) * xr.norm(dim=1)[None, :])
                return torch.cdist(qq, xr)

            qs64 = qd[:nqc]
            d_samp = exact(sample_rows, qs64)                         # [nqc, 4096]
            if bm_dev is not None:
                member = ((bm_dev[sample_rows >> 6] >> (sample_rows & 63)) & 1).bool()
                d_samp[:, ~member] = float("inf")
            kth = dst[:nqc, a.k - 1].to(torch.float64)
            tol = 1e-5 * torch.maximum(kth.abs(), (1.0 - kth).abs() if a.metric == "cosine" else kth.abs()) + 1e-7
            in_result = (sample_rows[None, None, :] == (ids[:nqc] - r0)[:, :, None]).any(dim=1)
            beaten = ((d_samp < (kth - tol)[:, None]) & ~in_result).sum().item()
            d_ret = torch.stack([exact((ids[i] - r0).clamp(min=0), qs64[i:i + 1])[0] for i in range(min(nqc, 8))])
            err = ((d_ret - dst[:min(nqc, 8)].to(torch.float64)).abs() /
                   torch.maximum(d_ret.abs(), (1.0 - d_ret).abs() if a.metric == "cosine" else d_ret.abs()).clamp(min=1e-30))
            props.update({"sampled_rows_beating_kth": int(beaten), "returned_distance_max_rel_err": float(err.max().item()),
                          "checked": f"{nqc} queries x 4096 random rows + the returned rows of 8 queries, float64 from the stored rows"})
            line["full_size_properties"] = props
        except Exception as e:  # the properties are a report, never a reason to lose the bench line
            line["full_size_properties"] = {"error": repr(e)}

    # ---- CPU baseline on a bounded sample of the same corpus + parity of the GPU path on that sample
    if not a.no_cpu and sample_host and a.bitmap_density == 0:
        from oracle import oracle as orc

        orc.build()
        sample = np.concatenate(sample_host) if len(sample_host) > 1 else sample_host[0]
        cores = os.cpu_count() or 1
        qs = q_dev[0][: min(cores, a.batch)].cpu().numpy()
        omc = {"l2": orc.L2, "cosine": orc.COSINE, "dot": orc.DOT}[a.metric]
        base, rows_used, want = cpu_baseline(a, sample, qs, omc, a.cpu_seconds)
        line["cpu_baseline"] = base
        sub = pk.VectorIndex(a.dim, dtype_code, device=local_rank)
        if scale is not None:
            sub.set_scale_artifact(pk.scale_artifact(scale))
        sub.append(torch.from_numpy(sample[:rows_used]).to(dev))
        sub.seal()
        got = sub.search(qs, a.k, metric_code)
        if a.dtype == "i8":
            ok = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)))
            line["parity"] = {"checked": f"{qs.shape[0]} queries x {rows_used} rows vs oracle", "bit_exact": ok}
        else:
            rel = float(np.nanmax(np.abs(got[1] - want[1]) / np.maximum(np.abs(want[1]), 1e-30)))
            same = float(np.mean(got[0] == want[0]))
            line["parity"] = {"checked": f"{qs.shape[0]} queries x {rows_used} rows vs oracle",
                              "max_rel_err": rel, "ids_equal_frac": same, "within_1e-5": rel <= 1e-5}
        sub.close()
    print(json.dumps(line), flush=True)
    if world > 1:
        comm.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
