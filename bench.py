#!/usr/bin/env python3
"""bench.py — queries/sec of the vector-similarity hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic queries: B queries against
the whole N x D corpus (resident in HBM), exact top-k per query.  With --gpus G the corpus is
row-sharded over G ranks (one process per GPU, torchrun), every rank scans its shard for the
same batch, one NCCL all-gather collects the per-shard top-k and a merge kernel produces the
global top-k (strong scaling: the corpus is fixed, per-GPU work shrinks).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--rows N] [--dim D] [--batch B] [--k K] [--dtype f32|i8|f16] [--metric cosine|l2|dot]

Prints ONE JSON line (see DESIGN.md "Measurement" for every key).  The default single-GPU run also measures
BASELINE configs 2 and 3 (1M x 768 f32 batch 256; 10M x 768 int8 dot batch 1024) as sub-records under "configs", the
int8 tensor peak of this box (torch._int_mm) that the tensor roofline is quoted against, and a sustained (>= 3 s) figure
beside the burst one.
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
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 2 / 3 sub-records")
    ap.add_argument("--sustain-seconds", type=float, default=3.0, help="length of the sustained-throughput leg (0 = skip)")
    return ap.parse_args()


def config_dict(a, world: int, row_bytes: int = 0) -> dict:
    """The `config` object of both arms (identical keys and values for identical flags)."""
    pad = lambda n, m: (n + m - 1) // m * m
    if not row_bytes:   # the image the default kernels scan: int8 image of f32/f16 rows, the codes of an int8 index
        row_bytes = a.dim + 4 if a.dtype == "i8" else pad(a.dim, 128) + 16
    shard_rows = (a.rows + world - 1) // world
    gb = shard_rows * row_bytes / 1e9
    return {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "batch": a.batch, "k": a.k, "metric": a.metric,
            "parallelism": f"row-shard x{world}" if world > 1 else "single GPU",
            "l2_policy": f"scanned image {gb:.2f} GB per GPU streams through L2 every step"
                         + (" (larger than the 126 MB L2)" if gb > 0.2 else ""),
            "query_sets": 4}


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


# ------------------------------------------------------------------ reference arm (CPU)
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
        "config": config_dict(a, max(int(os.environ.get("WORLD_SIZE", "1")), 1) if a.gpus > 1 else 1),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ peaks measured on this box
_PEAKS_CACHE = {}


def measured_tensor_peaks(torch, dev):
    """int8 tensor peak of THIS GPU: torch._int_mm 8192^3 (cuBLASLt IMMA), best of 10 (burst) and back to
    back for ~2 s (sustained), CUDA events.  2*N^3 ops per call."""
    if "int8" in _PEAKS_CACHE:
        return _PEAKS_CACHE["int8"]
    out = {"int8_tops_burst": None, "int8_tops_sustained": None, "how": "torch._int_mm 8192^3: best of 10 / back to back for 2 s"}
    try:
        n = 8192
        a8 = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
        b8 = torch.randint(-128, 127, (n, n), dtype=torch.int8, device=dev)
        for _ in range(3):
            torch._int_mm(a8, b8)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a8, b8)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["int8_tops_burst"] = 2.0 * n ** 3 / (best * 1e-3) / 1e12
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(10, int(2000.0 / best))
        e0.record()
        for _ in range(reps):
            torch._int_mm(a8, b8)
        e1.record()
        torch.cuda.synchronize()
        out["int8_tops_sustained"] = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del a8, b8
    except Exception as e:   # the peak is a denominator, never a reason to lose the line
        out["error"] = repr(e)
    _PEAKS_CACHE["int8"] = out
    return out


KERNEL_NAMES = {
    1: "scan_f32_simt", 2: "scan_i8_simt", 3: "scan_i8_ts / scan_i8_tc (tcgen05 kind::i8; > 128 queries: resident in TMEM)",
    4: "scan_float_tc (tcgen05 kind::tf32 filter on f32 rows + exact rescore)", 5: "scan_f16_simt",
    6: "scan_float_tc (tcgen05 kind::f16 on f16 rows + exact rescore)",
    7: "scan_float_tc (tcgen05 kind::f16 filter on the fp16 image of the f32 rows + exact rescore)",
    8: "scan_img8 (tcgen05 kind::i8 filter on the per-row-scaled int8 image of the rows, queries in TMEM; likely "
       "candidates re-scored by warps of the same kernel, band pairs after the final thresholds)"}


# ------------------------------------------------------------------ one workload on this rank's GPU
class Workload:
    """The resident index of one (rows, dim, dtype) corpus shard + its query sets, and the timed legs over it."""

    def __init__(self, torch, pk, dist, a, world, rank, local_rank):
        self.torch, self.pk, self.dist, self.a = torch, pk, dist, a
        self.world, self.rank, self.local_rank = world, rank, local_rank
        self.dev = torch.device("cuda", local_rank)
        self.dtype_code = {"f32": pk.F32, "i8": pk.I8, "f16": pk.F16}[a.dtype]
        self.metric_code = {"l2": pk.L2, "cosine": pk.COSINE, "dot": pk.DOT}[a.metric]
        self.elem = {"f32": 4, "i8": 1, "f16": 2}[a.dtype]
        self.comm = None
        self.merge_launches = 0
        self._build()

    # ---- corpus shard of this rank, generated on the device and appended through the C ABI
    def _build(self):
        torch, pk, dist, a, dev = self.torch, self.pk, self.dist, self.a, self.dev
        world, rank, local_rank = self.world, self.rank, self.local_rank
        per = (a.rows + world - 1) // world
        self.r0, self.r1 = min(rank * per, a.rows), min((rank + 1) * per, a.rows)
        r0, r1 = self.r0, self.r1
        ix = pk.VectorIndex(a.dim, self.dtype_code, device=local_rank)
        if a.force_simt:
            ix.set_option("force_simt", 1)
        for kv in a.opt:   # before the first allocation: image_mask decides which filter images exist
            name, value = kv.split("=")
            ix.set_option(name, int(value))
        ix.reserve(max(r1 - r0, 1))
        ix.set_row_base(r0)
        self.scale = None
        if a.dtype == "i8":
            # global symmetric absmax scale over the whole corpus (docs/vector-int8-quant.md:11-49)
            absmax = 0.0
            for b, off, n in corpus_blocks(r0, r1):
                absmax = max(absmax, pk.blob_absmax(gen_block(torch, b, BLOCK_ROWS, a.dim, dev)[off:off + n].contiguous(),
                                                    device=local_rank))
            t = torch.tensor([absmax], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            self.scale = pk.scale_from_absmax(float(t.item()))
            ix.set_scale_artifact(pk.scale_artifact(self.scale))
        sample_rows = min(r1 - r0, a.sample_rows) if rank == 0 else 0
        self.sample_host = []
        for b, off, n in corpus_blocks(r0, r1):
            x = gen_block(torch, b, BLOCK_ROWS, a.dim, dev)[off:off + n].contiguous()
            if a.dtype == "i8":
                x = pk.quantize_int8(x, self.scale, device=local_rank)
            elif a.dtype == "f16":
                x = x.half()
            ix.append(x)
            have = sum(s.shape[0] for s in self.sample_host)
            if have < sample_rows:
                self.sample_host.append(x[: sample_rows - have].cpu().numpy())
            del x
        ix.seal()
        torch.cuda.synchronize()
        self.ix = ix
        self.n_q_sets = 4
        self.q_dev = [self.make_queries(s) for s in range(self.n_q_sets)]
        self.q_pinned = [q.cpu().pin_memory() for q in self.q_dev]
        mk = lambda: (torch.empty((a.batch, a.k), dtype=torch.int64, device=dev),
                      torch.empty((a.batch, a.k), dtype=torch.float32, device=dev),
                      torch.empty(a.batch, dtype=torch.int32, device=dev))
        self.out_dev = mk()
        if world > 1:
            self.merged_out = mk()
            # the exchange lives in libpkv.so (pkv_comm_* / pkv_search_sharded_device: scan + pack + ONE ncclAllGather +
            # merge on one stream); torch.distributed only carries the 128-byte NCCL id and the timing barriers
            uid = torch.tensor(list(pk.Comm.unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device=dev)
            dist.broadcast(uid, 0)
            self.comm = pk.Comm(local_rank, rank, world, bytes(uid.cpu().numpy().tolist()))
        self.bm_dev = self.bm_host = None
        if a.bitmap_density > 0:
            self.set_bitmap(a.bitmap_density)

    def set_bitmap(self, density: float):
        """config 5: membership bits over this shard's rows (Bernoulli(p), seed 0x5EED+2 over GLOBAL rows: SURVEY 8d),
        packed LSB-first into u64 words; the same bitmap for every query of the batch (context of an AND filter).
        density 0 removes it."""
        self.bm_dev = self.bm_host = None
        self.a.bitmap_density = density
        if density <= 0:
            return
        a, r0, r1 = self.a, self.r0, self.r1
        member = np.random.default_rng(CORPUS_SEED + 2).random(a.rows) < density
        local = np.zeros(((r1 - r0 + 63) // 64) * 64, dtype=bool)
        local[: r1 - r0] = member[r0:r1]
        self.bm_host = np.packbits(local, bitorder="little").view(np.uint64).copy()
        self.bm_dev = self.torch.from_numpy(self.bm_host.view(np.int64)).to(self.dev)

    def make_queries(self, step):
        q = gen_queries(self.torch, self.a.batch, self.a.dim, self.dev, step)
        if self.a.dtype == "i8":
            return self.pk.quantize_int8(q, self.scale, device=self.local_rank)
        if self.a.dtype == "f16":
            return q.half()
        return q

    def close(self):
        if self.comm is not None:
            self.comm.close()
        self.ix.close()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def step_device(self, q):
        a = self.a
        if self.world == 1:
            return self.ix.search(q, a.k, self.metric_code, out=self.out_dev, bitmap=self.bm_dev)
        # ONE all-gather of the per-shard candidates: a pack kernel (12-byte entries), the collective, and a merge
        # kernel that reads the gathered buffer directly - all inside the library call
        self.merge_launches += 2
        return self.comm.search(self.ix, q, a.k, self.metric_code, bitmap=self.bm_dev, out=self.merged_out)

    def _allmax(self, values):
        t = self.torch.tensor(values, device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    # ---- device-resident timing (inputs already in HBM): `value`
    def time_device(self, steps, warmup, sample_clocks=True):
        torch = self.torch
        for i in range(max(warmup, 3)):
            self.step_device(self.q_dev[i % self.n_q_sets])
        self.barrier()
        c0 = self.ix.counters()
        ml0 = self.merge_launches
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0 and sample_clocks:
            sampler.start()
        scan_ms = 0.0
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            self.step_device(self.q_dev[i % self.n_q_sets])
            scan_ms += self.ix.counters().last_scan_ms
        ev1.record()
        self.barrier()
        clocks = sampler.stop() if (self.rank == 0 and sample_clocks) else None
        c1 = self.ix.counters()
        ms_total, scan_ms_max = self._allmax([ev0.elapsed_time(ev1), scan_ms])
        return {"ms_per_step": ms_total / steps, "scan_ms_per_step": scan_ms_max / steps, "clocks": clocks, "c0": c0, "c1": c1,
                "steps": steps, "merge_launches": self.merge_launches - ml0}

    # ---- the same loop for >= `seconds`: what the chip sustains once the power cap has settled
    def time_sustained(self, seconds):
        torch = self.torch
        steps, t_done = 0, 0.0
        sampler = ClockSampler(self.local_rank)
        if self.rank == 0:
            sampler.start()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        budget = 64
        while True:
            for _ in range(budget):
                self.step_device(self.q_dev[steps % self.n_q_sets])
                steps += 1
            t_done = time.perf_counter() - t0
            # every rank must run the same number of steps (the exchange is collective): rank 0 decides
            flag = self._allmax([1.0 if t_done >= seconds else 0.0]) if self.world > 1 else [1.0 if t_done >= seconds else 0.0]
            if flag[0] >= 1.0:
                break
        ev1.record()
        self.barrier()
        clocks = sampler.stop() if self.rank == 0 else None
        (ms_total,) = self._allmax([ev0.elapsed_time(ev1)])
        return {"value": self.a.batch / (ms_total / steps * 1e-3), "unit": "queries/s", "seconds": ms_total * 1e-3, "steps": steps,
                "ms_per_step": ms_total / steps, "clocks": clocks}

    # ---- end to end through the public API with HOST buffers: `e2e`
    def time_e2e(self, steps):
        torch, a = self.torch, self.a
        out_host = (np.empty((a.batch, a.k), np.int64), np.empty((a.batch, a.k), np.float32), np.empty(a.batch, np.int32))
        out_pinned = [torch.from_numpy(o).pin_memory() for o in out_host]
        out_pinned_np = tuple(o.numpy() for o in out_pinned)

        def step_e2e(qp):
            if self.world == 1:
                # the C-ABI host call: H2D of the queries, scan, D2H of ids/dist/counts, all inside
                return self.ix.search(qp.numpy(), a.k, self.metric_code, out=out_pinned_np, bitmap=self.bm_host)
            q = qp.to(self.dev, non_blocking=True)
            ids, dst, cnt = self.step_device(q)
            return ids.cpu(), dst.cpu(), cnt.cpu()

        for i in range(3):
            step_e2e(self.q_pinned[i % self.n_q_sets])
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            step_e2e(self.q_pinned[i % self.n_q_sets])
        self.barrier()
        (e2e_s,) = self._allmax([time.perf_counter() - t0])
        h2d = a.batch * a.dim * self.elem + (self.bm_host.nbytes if self.bm_host is not None else 0)
        d2h = a.batch * a.k * 12 + a.batch * 4
        return {"value": a.batch / (e2e_s / steps), "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "pkv_search (C ABI, host buffers)" if self.world == 1 else
                       "H2D + pkv_search_sharded_device (scan + pack + ncclAllGather + merge in libpkv.so) + D2H"}

    # ---- roofline of the dominant (scan) kernel, from the library's own CUDA-event timing of its scan launches on the
    # stream they run on; algorithmic bytes/ops as DESIGN.md states them
    def roofline(self, timed):
        a, torch = self.a, self.torch
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        c0, c1, steps = timed["c0"], timed["c1"], timed["steps"]
        scan_ms_step = timed["scan_ms_per_step"]
        shard_rows = self.r1 - self.r0
        kind = c1.last_scan_kind
        # bytes one pass of the dominant kernel must read per row: the image it scans (+4 B row norm on the paths that
        # read it, +16 B row figures of the int8 image)
        pad = lambda n, m: (n + m - 1) // m * m
        row_bytes = {1: a.dim * 4, 2: a.dim + 4, 3: a.dim + 4, 4: a.dim * 4 + 4, 5: a.dim * 2, 6: a.dim * 2 + 4,
                     7: pad(a.dim, 64) * 2 + 4, 8: pad(a.dim, 128) + 16}.get(kind, a.dim * self.elem)
        stored_row_bytes = a.dim * self.elem
        alg_bytes = shard_rows * row_bytes
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved_gbs = alg_bytes / (scan_ms_step * 1e-3) / 1e9 if scan_ms_step > 0 else 0.0
        ops = 2.0 * a.batch * shard_rows * a.dim
        tput = ops / (scan_ms_step * 1e-3) / 1e12 if scan_ms_step > 0 else 0.0
        # Which roof binds: passes of <= 128 queries stream the scanned image at HBM speed (ncu: ~85 % DRAM throughput,
        # tensor pipe ~40 % active); wider passes keep the tensor pipe busiest (profiles/r0*_ncu_*): tensor-bound.
        tensor_bound = kind in (3, 4, 6, 7, 8) and a.batch > 128
        int8_pipe = a.dtype == "i8" or kind == 8
        bf16_peak = peaks.get("bf16_tflops", 1590.0)
        tp = measured_tensor_peaks(torch, self.dev) if (tensor_bound and int8_pipe) else {}
        if tensor_bound and int8_pipe and tp.get("int8_tops_burst"):
            tensor_peak = tp["int8_tops_burst"]
            peak_src = "measured in this run: torch._int_mm 8192^3 int8 (best of 10)"
        else:
            tensor_peak = 2.0 * bf16_peak if int8_pipe else (0.5 * bf16_peak if kind == 4 else bf16_peak)
            peak_src = "MEASURED_PEAKS.json cuBLAS bf16 burst" + (" x2 for the int8 pipe" if int8_pipe else "")
        # DRAM traffic per step, COUNTED from what the launches read (the ncu captures under profiles/ agree within a
        # few %: r02_ncu_*): every pass streams the scanned image (query groups of one launch share row tiles through
        # L2: ncu measured 1.06x / 1.44x the image bytes at 2 / 4 groups), plus one stored row (and the query, from L2)
        # per re-scored pair, plus the candidate / parked-pair lists
        sub = min(a.batch, 1024)
        img_passes = -(-a.batch // 1024) * ((1 if sub <= 128 or kind in (3, 8) else -(-sub // 256)) if kind in (3, 4, 6, 7, 8) else -(-sub // 8))
        groups = min(-(-sub // 256), 4) if (kind in (3, 8) and sub > 128) else 1
        # (4 groups: 1.44 on the round-1 chunked schedule, 2.30 in the round-2 live launch, where the groups drift apart:
        # profiles/r02_ncu_i8_b1024_live.txt, dram__bytes_read 17.75 GB for 7.72 GB of codes)
        share = {1: 1.002, 2: 1.06, 3: 1.25, 4: 2.30}[groups]
        rescored = (c1.rescored_pairs - c0.rescored_pairs + c1.deferred_pairs - c0.deferred_pairs) / steps
        traffic = img_passes * alg_bytes * share + rescored * stored_row_bytes
        roof = {
            "bound": "tensor" if tensor_bound else "hbm",
            "achieved": tput if tensor_bound else achieved_gbs,
            "peak": tensor_peak if tensor_bound else hbm_peak,
            "unit": ("TOP/s" if int8_pipe else "TFLOP/s") if tensor_bound else "GB/s",
            "frac": None,
            "traffic": traffic,
            "traffic_note": f"counted: {img_passes} image pass(es) x {alg_bytes} B x {share} (L2 tile sharing of {groups} query "
                            f"group(s), ncu-measured) + {rescored:.0f} re-scored pairs x {stored_row_bytes} B stored row",
            "kernel": KERNEL_NAMES.get(kind, str(kind)),
            "algorithmic_bytes_per_row": row_bytes,
            "algorithmic_bytes_per_step": alg_bytes, "algorithmic_ops_per_step": ops,
            "launches_per_step": (c1.scan_launches - c0.scan_launches) / steps, "kernel_ms_per_step": scan_ms_step,
            "achieved_gbs": achieved_gbs, "achieved_tops": tput,
            "hbm_frac": achieved_gbs / hbm_peak if hbm_peak else None,
            "tensor_frac": tput / tensor_peak if tensor_peak else None,
            "tensor_frac_of_nominal": tput / (4500.0 if int8_pipe else 2250.0),
            "int8_peak_measured": tp or None,
            "peak_source": (peak_src if tensor_bound else
                            ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)")),
            # an f32 index is scanned through a 3.9x smaller int8 image, so the scan outruns what streaming the f32 rows
            # allows; this is a speed-up over that design, NOT a roofline fraction
            "speedup_over_f32_streaming_roofline": (shard_rows * a.dim * 4 / (scan_ms_step * 1e-3) / 1e9 / hbm_peak)
            if (a.dtype == "f32" and kind in (7, 8) and scan_ms_step > 0 and hbm_peak) else None,
        }
        roof["frac"] = roof["achieved"] / roof["peak"] if roof["peak"] else None
        return roof

    def search_stats(self, timed):
        c0, c1, steps, a = timed["c0"], timed["c1"], timed["steps"], self.a
        return {"live_refreshes_per_step": (c1.live_refreshes - c0.live_refreshes) / steps,
                "live_refresh_skips_per_step": (c1.live_refresh_skips - c0.live_refresh_skips) / steps,
                "rescored_rows_per_query": (c1.rescored_pairs - c0.rescored_pairs) / steps / a.batch,
                "deferred_rows_per_query": (c1.deferred_pairs - c0.deferred_pairs) / steps / a.batch}

    # ---- full-size properties of the last batch's result (the oracle cannot scan 10M rows in the time budget):
    # sorted, unique, idempotent, and a sampled exactness certificate - the distance of every returned row and of 4096
    # random rows of this shard is recomputed from the STORED rows in float64 with torch; returned distances must match
    # and no sampled row outside the result may beat the k-th best.
    def full_size_properties(self, steps):
        torch, a, ix, dev, r0, r1 = self.torch, self.a, self.ix, self.dev, self.r0, self.r1
        try:
            qd = self.q_dev[(steps - 1) % self.n_q_sets]
            ids, dst, cnt = (t.clone() for t in ix.search(qd, a.k, self.metric_code, bitmap=self.bm_dev))
            ids2, dst2, _ = ix.search(qd, a.k, self.metric_code, bitmap=self.bm_dev)
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
                qq = qq.to(torch.float64)
                dots = qq @ xr.T
                if a.metric == "dot":
                    return -dots
                if a.metric == "cosine":
                    return 1.0 - dots / (qq.norm(dim=1, keepdim=True) * xr.norm(dim=1)[None, :])
                return torch.cdist(qq, xr)

            qs64 = qd[:nqc]
            d_samp = exact(sample_rows, qs64)                         # [nqc, 4096]
            if self.bm_dev is not None:
                member = ((self.bm_dev[sample_rows >> 6] >> (sample_rows & 63)) & 1).bool()
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
            return props
        except Exception as e:  # the properties are a report, never a reason to lose the bench line
            return {"error": repr(e)}

    # ---- CPU baseline on a bounded sample of the same corpus + parity of the GPU path on that sample: the WHOLE batch
    # is searched on the GPU (so the kernels that were timed are the ones checked) and compared with the oracle on 64
    # queries spread over it
    def cpu_baseline_and_parity(self, seconds):
        from oracle import oracle as orc

        torch, pk, a = self.torch, self.pk, self.a
        orc.build()
        sample = np.concatenate(self.sample_host) if len(self.sample_host) > 1 else self.sample_host[0]
        sel = np.unique(np.linspace(0, a.batch - 1, min(64, a.batch)).astype(np.int64))
        q_all = self.q_dev[0].cpu().numpy()
        qs = np.ascontiguousarray(q_all[sel])
        omc = {"l2": orc.L2, "cosine": orc.COSINE, "dot": orc.DOT}[a.metric]
        base, rows_used, want = cpu_baseline(a, sample, qs, omc, seconds)
        sub = pk.VectorIndex(a.dim, self.dtype_code, device=self.local_rank)
        if self.scale is not None:
            sub.set_scale_artifact(pk.scale_artifact(self.scale))
        sub.append(torch.from_numpy(sample[:rows_used]).to(self.dev))
        sub.seal()
        got_all = sub.search(self.q_dev[0], a.k, self.metric_code)
        kind = sub.counters().last_scan_kind
        got = tuple(t.cpu().numpy()[sel] for t in got_all)
        checked = (f"GPU: all {a.batch} queries of the timed batch shape x {rows_used} rows (scan kind {kind}); oracle: "
                   f"{len(sel)} of them, spread over the batch")
        if a.dtype == "i8":
            ok = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)))
            parity = {"checked": checked, "bit_exact": ok}
        else:
            rel = float(np.nanmax(np.abs(got[1] - want[1]) / np.maximum(np.maximum(np.abs(want[1]), np.abs(1.0 - want[1])
                                                                                   if a.metric == "cosine" else 0.0), 1e-30)))
            parity = {"checked": checked, "max_rel_err": rel, "ids_equal_frac": float(np.mean(got[0] == want[0])),
                      "within_1e-5": rel <= 1e-5,
                      "tolerance": "relative to max(|d|, |1-d|) for cosine (the score is 1-d), to |d| otherwise"}
        sub.close()
        return base, parity


def cpu_baseline(a, corpus_sample: np.ndarray, queries: np.ndarray, metric_code: int, seconds: float):
    """Times the oracle's restatement of the reference scan (one query per thread at a time, like one SQLite connection
    per thread) on a bounded sample: Q queries x R rows on all host cores; QPS at the full N rows is extrapolated
    linearly (the scan is linear in rows)."""
    from oracle import oracle as orc

    cores = os.cpu_count() or 1
    threads = min(cores, queries.shape[0])
    nq = queries.shape[0]
    probe = min(20_000, corpus_sample.shape[0])
    t0 = time.perf_counter()
    orc.topk(corpus_sample[:probe], queries, metric_code, a.k, threads=threads)
    rate = probe * nq / max(time.perf_counter() - t0, 1e-6)   # (query, row) pairs per second on all threads
    rows = int(min(corpus_sample.shape[0], max(probe, rate * seconds / nq)))
    t0 = time.perf_counter()
    out = orc.topk(corpus_sample[:rows], queries, metric_code, a.k, threads=threads)
    dt = time.perf_counter() - t0
    qps_full = (nq * rows / dt) / a.rows
    return {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{nq} queries ({threads} threads, one query per thread at a time) x first {rows} rows in {dt:.2f} s, "
                      f"linearly extrapolated to {a.rows} rows",
            "pairs_per_s_per_core": nq * rows / dt / threads}, rows, out


def _record(w, a, timed, e2e, target):
    roof = w.roofline(timed)
    return {"workload": workload_name(a), "n_gpus": w.world, "value": a.batch / (timed["ms_per_step"] * 1e-3), "unit": "queries/s",
            "ms_per_step": timed["ms_per_step"], "steps": timed["steps"], "e2e": e2e["value"] if e2e else None, "target": target,
            "roofline": {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms_per_step", "hbm_frac",
                                              "tensor_frac", "tensor_frac_of_nominal", "peak_source",
                                              "speedup_over_f32_streaming_roofline")},
            "gpu_launches_per_step": (timed["c1"].kernel_launches - timed["c0"].kernel_launches + timed["merge_launches"])
            / timed["steps"],
            "search_stats": w.search_stats(timed)}


def sub_config(torch, pk, base_args, name, target, dist=None, world=1, rank=0, **over):
    """One BASELINE config as a sub-record of the default line: its own index (row-sharded over `world` ranks: every
    rank calls this, rank 0 gets the record), device-resident value, e2e, roofline."""
    import copy

    a = copy.copy(base_args)
    for k, v in over.items():
        setattr(a, k, v)
    a.opt, a.bitmap_density, a.force_simt, a.sample_rows = [], 0.0, False, 0
    try:
        w = Workload(torch, pk, dist, a, world, rank, int(os.environ.get("LOCAL_RANK", "0")))
        timed = w.time_device(a.steps, 3, sample_clocks=False)
        e2e = w.time_e2e(max(5, a.steps // 2))
        rec = _record(w, a, timed, e2e, target) if rank == 0 else None
        w.close()
        del w
        torch.cuda.empty_cache()
        return rec
    except Exception as e:
        return {"workload": name, "error": repr(e)}


def bitmap_leg(w, density, steps, target):
    """BASELINE config 5 on the resident corpus of the main workload: the same searches AND a tag-filter row bitmap
    (collective when sharded: every rank calls this)."""
    import copy

    keep = w.a
    try:
        w.a = copy.copy(keep)
        w.set_bitmap(density)
        timed = w.time_device(steps, 3, sample_clocks=False)
        e2e = w.time_e2e(max(5, steps // 2))
        return _record(w, w.a, timed, e2e, target) if w.rank == 0 else None
    except Exception as e:
        return {"workload": "C5", "error": repr(e)}
    finally:
        w.bm_dev = w.bm_host = None
        w.a = keep


def sharded_parity(torch, pk, dist, base_args, world, rank, local_rank, rows=2_000_000, nq_oracle=32):
    """N-rank parity inside the driver's own scaling run: the first `rows` rows of the corpus row-sharded over the
    `world` ranks, the WHOLE batch searched through pkv_search_sharded_device (scan + pack + ONE ncclAllGather + merge), and
    `nq_oracle` of its queries compared with the oracle scanning the same rows on rank 0's host cores."""
    import copy

    a = copy.copy(base_args)
    a.rows, a.opt, a.bitmap_density, a.force_simt, a.sample_rows = min(rows, base_args.rows), [], 0.0, False, 0
    try:
        w = Workload(torch, pk, dist, a, world, rank, local_rank)
        got_all = tuple(t.clone() for t in w.step_device(w.q_dev[0]))
        torch.cuda.synchronize()
        shard_rows, kind = w.r1 - w.r0, w.ix.counters().last_scan_kind
        rec = None
        if rank == 0:
            from oracle import oracle as orc

            orc.build()
            x = np.concatenate([gen_block(torch, b, BLOCK_ROWS, a.dim, w.dev)[off:off + n].cpu().numpy()
                                for b, off, n in corpus_blocks(0, a.rows)])
            sel = np.unique(np.linspace(0, a.batch - 1, min(nq_oracle, a.batch)).astype(np.int64))
            qs = np.ascontiguousarray(w.q_dev[0].cpu().numpy()[sel])
            omc = {"l2": orc.L2, "cosine": orc.COSINE, "dot": orc.DOT}[a.metric]
            t0 = time.perf_counter()
            want = orc.topk(x, qs, omc, a.k, threads=min(os.cpu_count() or 1, len(sel)))
            dt = time.perf_counter() - t0
            got = tuple(t.cpu().numpy()[sel] for t in got_all)
            den = np.maximum(np.maximum(np.abs(want[1]), np.abs(1.0 - want[1]) if a.metric == "cosine" else 0.0), 1e-30)
            rel = float(np.nanmax(np.abs(got[1] - want[1]) / den))
            rec = {"checked": f"{world} ranks x {shard_rows} rows (first {a.rows} rows of the corpus), all {a.batch} queries through "
                              f"pkv_search_sharded_device (scan kind {kind}); oracle: {len(sel)} of them over the same rows in {dt:.1f} s",
                   "max_rel_err": rel, "within_1e-5": rel <= 1e-5, "ids_equal_frac": float(np.mean(got[0] == want[0])),
                   "counts_equal": bool(np.array_equal(got[2], want[2])),
                   "tolerance": "relative to max(|d|, |1-d|) for cosine (the score is 1-d), to |d| otherwise"}
        w.close()
        del w
        torch.cuda.empty_cache()
        return rec
    except Exception as e:
        return {"error": repr(e)}


class Watchdog:
    """The multi-rank sub-records are extras: if one of them hangs a collective (a rank died, an exchange never
    completed), rank 0 still prints the main line - marked - and every rank leaves."""

    def __init__(self, seconds, rank, line):
        self.line, self.rank = line, rank
        self._t = threading.Timer(seconds + (0.0 if rank == 0 else 5.0), self._fire)
        self._t.daemon = True
        self._t.start()

    def _fire(self):
        if self.rank == 0:
            self.line.setdefault("configs", {})["error"] = "multi-rank sub-records timed out; main line is complete"
            print(json.dumps(self.line), flush=True)
        os._exit(0)

    def cancel(self):
        self._t.cancel()


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
    a.sample_rows = 2_000_000

    w = Workload(torch, pk, dist, a, world, rank, local_rank)
    timed = w.time_device(a.steps, a.warmup)
    e2e = w.time_e2e(a.steps)
    sustained = w.time_sustained(a.sustain_seconds) if a.sustain_seconds > 0 else None

    default_corpus = (a.rows == 10_000_000 and a.dim == 768 and a.batch == 256 and a.dtype == "f32" and a.metric == "cosine"
                      and a.bitmap_density == 0 and not a.opt and not a.force_simt)
    line = {}
    if rank == 0:
        roof = w.roofline(timed)
        line = {
            "metric": "queries/sec", "value": a.batch / (timed["ms_per_step"] * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": timed["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": a.dtype, "data": "synthetic",
            "config": config_dict(a, world, roof["algorithmic_bytes_per_row"]),
            "clocks": timed["clocks"],
            "e2e": e2e,
            "gpu_launches": int(timed["c1"].kernel_launches - timed["c0"].kernel_launches) + timed["merge_launches"],
            "roofline": roof,
            "sustained": sustained,
            # queries searched a second time: candidate-list overflows (whole batch redone) + guessed thresholds that proved
            # too tight (that query alone redone)
            "overflow_rescans": int(timed["c1"].fallback_queries - timed["c0"].fallback_queries),
            # what the in-kernel machinery did per step (device counters of the searches in the timed region)
            "search_stats": w.search_stats(timed),
        }
        if world == 1:
            line["full_size_properties"] = w.full_size_properties(a.steps)
        if world == 1 and not a.no_cpu and w.sample_host and a.bitmap_density == 0:
            line["cpu_baseline"], line["parity"] = w.cpu_baseline_and_parity(a.cpu_seconds)

    # ---- more than one GPU: BASELINE configs 4 and 5 (the ones quoted on 8 GPUs) and an N-rank parity leg as
    # sub-records of the scaling line, so that the driver's own 2/4/8-GPU runs hold them.  Collective: every rank runs
    # them; a watchdog keeps a stuck extra from costing the main line.
    if world > 1 and default_corpus and not a.no_configs:
        dog = Watchdog(420.0, rank, line)
        steps2 = max(10, min(a.steps, 30))
        c5 = bitmap_leg(w, 0.1, steps2, "config 5: vector top-k AND tag-filter bitmap (density 0.1), 10M corpus")
        w.close()
        del w
        torch.cuda.empty_cache()
        par = sharded_parity(torch, pk, dist, a, world, rank, local_rank)
        c4 = sub_config(torch, pk, a, "C4", "config 4: 50Mx512 fp16 (OpenCLIP ViT-B/32 shape), batch 4096, row-sharded",
                        dist=dist, world=world, rank=rank, rows=50_000_000, dim=512, dtype="f16", metric="cosine", batch=4096,
                        steps=5)
        dog.cancel()
        if rank == 0:
            line["parity"] = par
            line["configs"] = {"C4": c4, "C5": c5}
    else:
        w.close()
        del w
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- BASELINE configs 2 and 3 as sub-records of the default line (their targets: >= 70 % of the HBM roofline of an
    # f32-streaming scan = 382 k queries/s; >= 40 % of the int8 tensor peak)
    if world == 1 and default_corpus and not a.no_configs:
        a.sample_rows = 0
        steps2 = max(10, min(a.steps, 40))
        line["configs"] = {
            "C2": sub_config(torch, pk, a, "C2", "382000 queries/s (70 % of the f32-streaming HBM roofline, SURVEY 8d)",
                             rows=1_000_000, dtype="f32", metric="cosine", batch=256, steps=steps2),
            "C3": sub_config(torch, pk, a, "C3", "tensor_frac_of_nominal >= 0.40 (117 k queries/s at 4.5 POPS nominal)",
                             rows=10_000_000, dtype="i8", metric="dot", batch=1024, steps=max(10, steps2 // 2)),
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
