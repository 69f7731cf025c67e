"""ctypes front-end to oracle/libpkv_oracle.so plus an independent NumPy mirror.

TEST INFRASTRUCTURE ONLY — see oracle/pkv_oracle.h.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package panoptikon_b200 never does.

The NumPy mirror (`np_*`) restates the same arithmetic a second time, column by
column, so that the sequential f32 accumulation order of the sqlite-vec scalar
loops is reproduced exactly while staying vectorised over rows.  It exists to
cross-check the C restatement, not to be fast.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpkv_oracle.so")

F32, I8, F16 = 0, 1, 2
L2, COSINE, DOT = 0, 1, 2
AGG_MIN, AGG_MAX, AGG_AVG = 0, 1, 2

_NP_DTYPE = {F32: np.float32, I8: np.int8, F16: np.float16}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (idempotent)."""
    src = os.path.join(_HERE, "pkv_oracle.c")
    hdr = os.path.join(_HERE, "pkv_oracle.h")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libpkv_oracle.so"])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_scale_from_absmax.restype = C.c_float
        L.orc_scale_from_absmax.argtypes = [C.c_float]
        L.orc_scale_artifact.argtypes = [C.c_float, C.c_void_p]
        L.orc_artifact_scale.restype = C.c_int
        L.orc_artifact_scale.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_float)]
        L.orc_blob_absmax.restype = C.c_float
        L.orc_blob_absmax.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_quantize_int8.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
        for name in (
            "orc_distance_cosine_f32",
            "orc_distance_l2_f32",
            "orc_distance_cosine_i8",
            "orc_distance_l2_i8",
            "orc_distance_dot_f32",
            "orc_distance_dot_i8",
        ):
            fn = getattr(L, name)
            fn.restype = C.c_float
            fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_topk.restype = C.c_int
        L.orc_topk.argtypes = [
            C.c_void_p, C.c_int64, C.c_int, C.c_int,
            C.c_void_p, C.c_int, C.c_int, C.c_int,
            C.c_void_p, C.c_int64, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p,
        ]
        L.orc_distances.restype = C.c_int
        L.orc_distances.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_aggregate.restype = C.c_int
        L.orc_aggregate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ codec (C)

def scale_from_absmax(absmax: float) -> float:
    return float(lib().orc_scale_from_absmax(C.c_float(absmax)))


def scale_artifact(scale: float) -> bytes:
    buf = (C.c_uint8 * 4)()
    lib().orc_scale_artifact(C.c_float(scale), buf)
    return bytes(buf)


def artifact_scale(artifact: bytes):
    out = C.c_float()
    buf = (C.c_uint8 * max(len(artifact), 1)).from_buffer_copy(artifact.ljust(1, b"\0"))
    ok = lib().orc_artifact_scale(buf, len(artifact), C.byref(out))
    return float(out.value) if ok else None


def blob_absmax(blob: bytes) -> float:
    buf = np.frombuffer(blob, dtype=np.uint8)
    return float(lib().orc_blob_absmax(_ptr(buf), len(blob)))


def quantize_int8(blob: bytes, scale: float) -> bytes:
    buf = np.frombuffer(blob, dtype=np.uint8)
    out = np.empty(len(blob) // 4, dtype=np.uint8)
    lib().orc_quantize_int8(_ptr(buf), len(blob), C.c_float(scale), _ptr(out))
    return out.tobytes()


def quantize_rows(x: np.ndarray, scale: float) -> np.ndarray:
    """quantize_int8 applied to a C-contiguous f32 matrix; returns int8 of the same shape."""
    x = np.ascontiguousarray(x, dtype="<f4")
    out = np.empty(x.shape, dtype=np.int8)
    lib().orc_quantize_int8(_ptr(x), x.size * 4, C.c_float(scale), _ptr(out))
    return out


# -------------------------------------------------------------- distances (C)

def distance(a: np.ndarray, b: np.ndarray, metric: int) -> float:
    assert a.dtype == b.dtype and a.shape == b.shape and a.ndim == 1
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    kind = "f32" if a.dtype == np.float32 else "i8"
    assert a.dtype in (np.float32, np.int8)
    name = {L2: "l2", COSINE: "cosine", DOT: "dot"}[metric]
    fn = getattr(lib(), f"orc_distance_{name}_{kind}")
    return float(fn(_ptr(a), _ptr(b), a.shape[0]))


def _dtype_code(a: np.ndarray) -> int:
    if a.dtype == np.float32:
        return F32
    if a.dtype == np.int8:
        return I8
    if a.dtype == np.float16:
        return F16
    raise TypeError(a.dtype)


def distances(corpus: np.ndarray, query: np.ndarray, metric: int) -> np.ndarray:
    corpus = np.ascontiguousarray(corpus)
    query = np.ascontiguousarray(query)
    n, d = corpus.shape
    out = np.empty(n, dtype=np.float32)
    rc = lib().orc_distances(_ptr(corpus), n, d, _dtype_code(corpus), _ptr(query), metric, _ptr(out))
    assert rc == 0
    return out


def topk(corpus, queries, metric, k, bitmap=None, bitmap_stride=0, threads=1):
    """Returns (rows[nq,k] int64, dist[nq,k] f32, counts[nq] int32)."""
    corpus = np.ascontiguousarray(corpus)
    queries = np.ascontiguousarray(queries)
    assert corpus.dtype == queries.dtype
    n, d = corpus.shape
    nq = queries.shape[0]
    rows = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float32)
    counts = np.empty(nq, dtype=np.int32)
    if bitmap is not None:
        bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
    rc = lib().orc_topk(
        _ptr(corpus), n, d, _dtype_code(corpus), _ptr(queries), nq, metric, k,
        _ptr(bitmap), bitmap_stride, threads, _ptr(rows), _ptr(dist), _ptr(counts),
    )
    assert rc == 0
    return rows, dist, counts


def aggregate(dist, item_of_row, n_items, agg, weights=None):
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    item_of_row = np.ascontiguousarray(item_of_row, dtype=np.int64)
    if weights is not None:
        weights = np.ascontiguousarray(weights, dtype=np.float32)
    out = np.empty(n_items, dtype=np.float64)
    rc = lib().orc_aggregate(_ptr(dist), _ptr(item_of_row), _ptr(weights), dist.shape[0], n_items, agg, _ptr(out))
    assert rc == 0
    return out


# ------------------------------------------------------------- NumPy mirror

def np_scale_from_absmax(absmax) -> np.float32:
    absmax = np.float32(absmax)
    if absmax > 0 and np.isfinite(absmax):
        return np.float32(absmax / np.float32(127.0))
    return np.float32(1.0)


def np_quantize_int8(x: np.ndarray, scale) -> np.ndarray:
    """vector_quants.rs:1489-1497 in NumPy: np.rint is round-half-to-even."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(all="ignore"):
        r = np.rint(x / np.float32(scale)).astype(np.float32)
    nan = np.isnan(r)
    r = np.clip(r, np.float32(-128.0), np.float32(127.0))
    r = np.where(nan, np.float32(0.0), r)
    return r.astype(np.int8)


def np_distances(corpus: np.ndarray, query: np.ndarray, metric: int) -> np.ndarray:
    """All-rows distances with the scalar loops' sequential f32 accumulation order.

    Column i is added to every row's running f32 sum before column i+1, which
    is exactly `for i in 0..d { acc += ... }` executed independently per row.
    """
    if corpus.dtype == np.float16:
        corpus = corpus.astype(np.float32)
        query = query.astype(np.float32)
    n, d = corpus.shape
    is_i8 = corpus.dtype == np.int8
    a = corpus.astype(np.int32) if is_i8 else corpus
    q = query.astype(np.int32) if is_i8 else query
    f32 = np.float32
    with np.errstate(all="ignore"):
        if metric == L2:
            res = np.zeros(n, dtype=f32)
            for i in range(d):
                t = (a[:, i] - q[i]).astype(f32)
                res = (res + (t * t).astype(f32)).astype(f32)
            return np.sqrt(res.astype(np.float64)).astype(f32)
        dot = np.zeros(n, dtype=f32)
        amag = np.zeros(n, dtype=f32)
        bmag = f32(0)
        for i in range(d):
            dot = (dot + (a[:, i] * q[i]).astype(f32)).astype(f32)
            if metric == COSINE:
                amag = (amag + (a[:, i] * a[:, i]).astype(f32)).astype(f32)
                bmag = f32(bmag + f32(q[i] * q[i]))
        if metric == DOT:
            return (f32(0) - dot).astype(f32)
        den = np.sqrt(amag.astype(np.float64)) * np.sqrt(np.float64(bmag))
        return (1.0 - dot.astype(np.float64) / den).astype(f32)


def np_order(dist: np.ndarray) -> np.ndarray:
    """Row order under the contract: ascending distance, NaN last, ties by row."""
    nan = np.isnan(dist)
    key = np.where(nan, np.float32(np.inf), dist)
    return np.lexsort((np.arange(dist.shape[0]), key, nan))


def np_topk(corpus, queries, metric, k, bitmap=None, bitmap_stride=0):
    nq = queries.shape[0]
    n = corpus.shape[0]
    rows = np.full((nq, k), -1, dtype=np.int64)
    dist = np.full((nq, k), np.nan, dtype=np.float32)
    counts = np.zeros(nq, dtype=np.int32)
    for qi in range(nq):
        d = np_distances(corpus, queries[qi], metric)
        order = np_order(d)
        if bitmap is not None:
            bm = bitmap[qi * bitmap_stride: qi * bitmap_stride + (n + 63) // 64] if bitmap_stride else bitmap
            member = ((bm[order >> 6] >> (order & 63).astype(np.uint64)) & np.uint64(1)).astype(bool)
            order = order[member]
        order = order[:k]
        counts[qi] = len(order)
        rows[qi, : len(order)] = order
        dist[qi, : len(order)] = d[order]
    return rows, dist, counts


# ------------------------------------------------------ synthetic workloads

CORPUS_SEED = 0x5EED


def synthetic(n: int, d: int, seed: int = CORPUS_SEED, normalise: bool = True) -> np.ndarray:
    """SURVEY §8d recipe: default_rng(seed).standard_normal((n,d), f32), L2-normalised rows
    (mirrors tools/pql-equivalence/run_suite.py:532-542)."""
    x = np.random.default_rng(seed).standard_normal((n, d), dtype=np.float32)
    if normalise:
        for b in range(0, n, 65536):  # chunked: keeps the f64 temporary small
            blk = x[b:b + 65536]
            blk /= np.linalg.norm(blk.astype(np.float64), axis=1, keepdims=True).astype(np.float32)
    return x
