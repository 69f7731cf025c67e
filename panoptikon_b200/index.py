"""VectorIndex: the corpus of one embedding setter resident in one B200's HBM, searched by
libpkv's sm_100a kernels.  Thin wrapper over the C ABI (include/pkv.h); accepts NumPy arrays
(host path: H2D/D2H inside the call) or torch CUDA tensors (device path)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _native as N

_NP = {N.F32: np.float32, N.I8: np.int8, N.F16: np.float16}
_CODE = {np.dtype(np.float32): N.F32, np.dtype(np.int8): N.I8, np.dtype(np.float16): N.F16}


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _torch_code(t) -> int:
    import torch

    return {torch.float32: N.F32, torch.int8: N.I8, torch.float16: N.F16}[t.dtype]


class VectorIndex:
    def __init__(self, dim: int, dtype: int = N.F32, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().pkv_index_create(device, dim, dtype, C.byref(self._h)))
        self.dim, self.dtype, self.device = dim, dtype, device

    # -- lifecycle -----------------------------------------------------------
    def close(self) -> None:
        if self._h:
            N.lib().pkv_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reserve(self, rows: int) -> None:
        N.check(N.lib().pkv_index_reserve(self._h, rows))

    def append(self, rows, row_ids=None) -> None:
        """rows: [n, dim] array of the index dtype — NumPy (host) or torch CUDA tensor."""
        if _is_torch(rows):
            assert rows.is_cuda and rows.is_contiguous() and rows.dim() == 2 and rows.shape[1] == self.dim
            assert _torch_code(rows) == self.dtype
            ids_ptr = None
            if row_ids is not None:
                import torch

                assert row_ids.is_cuda and row_ids.dtype == torch.int64 and row_ids.is_contiguous()
                ids_ptr = C.c_void_p(row_ids.data_ptr())
            N.check(N.lib().pkv_index_append_device(self._h, C.c_void_p(rows.data_ptr()), ids_ptr, rows.shape[0]))
            return
        rows = np.ascontiguousarray(rows, dtype=_NP[self.dtype])
        assert rows.ndim == 2 and rows.shape[1] == self.dim
        ids = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int64)
        assert ids is None or ids.shape == (rows.shape[0],)
        N.check(N.lib().pkv_index_append(self._h, _np_ptr(rows), _np_ptr(ids), rows.shape[0]))

    def append_blobs(self, blob: bytes, row_ids=None) -> None:
        """Concatenated SQLite blobs exactly as stored (n * dim * elem_size bytes)."""
        a = np.frombuffer(blob, dtype=_NP[self.dtype])
        self.append(a.reshape(-1, self.dim), row_ids)

    def set_scale_artifact(self, artifact: bytes) -> None:
        buf = (C.c_uint8 * max(len(artifact), 1)).from_buffer_copy(artifact.ljust(1, b"\0"))
        N.check(N.lib().pkv_index_set_scale(self._h, buf, len(artifact)))

    def set_row_base(self, base: int) -> None:
        N.check(N.lib().pkv_index_set_row_base(self._h, base))

    def seal(self) -> None:
        N.check(N.lib().pkv_index_seal(self._h))

    def info(self) -> N.IndexInfo:
        out = N.IndexInfo()
        N.check(N.lib().pkv_index_get_info(self._h, C.byref(out)))
        return out

    def counters(self) -> N.Counters:
        out = N.Counters()
        N.check(N.lib().pkv_index_counters(self._h, C.byref(out)))
        return out

    def set_option(self, name: str, value: int) -> None:
        N.check(N.lib().pkv_index_set_option(self._h, name.encode(), value))

    @property
    def rows(self) -> int:
        return int(self.info().rows)

    # -- search ----------------------------------------------------------------
    def search(self, queries, k: int, metric: int = N.COSINE, bitmap=None, bitmap_stride_words: int = 0,
               out=None, stream: int = 0):
        """Top-k per query under (distance asc, row asc, NaN last).

        NumPy queries -> (ids[nq,k] int64, dist[nq,k] f32, counts[nq] int32) NumPy arrays.
        torch CUDA queries -> the same as CUDA tensors (device path, no host copies)."""
        if _is_torch(queries):
            return self._search_device(queries, k, metric, bitmap, bitmap_stride_words, out, stream)
        queries = np.ascontiguousarray(queries)
        if queries.ndim == 1:
            queries = queries[None, :]
        qd = _CODE.get(queries.dtype)
        if qd is None:
            raise TypeError(f"unsupported query dtype {queries.dtype}")
        if queries.shape[1] != self.dim:
            raise N.PkvError(N.ERR_DIM_MISMATCH, f"query dimension {queries.shape[1]} != index dimension {self.dim}")
        nq = queries.shape[0]
        p = N.SearchParams(metric=metric, k=k, query_dtype=qd)
        if bitmap is not None:
            bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
            p.bitmap = bitmap.ctypes.data
            p.bitmap_stride_words = bitmap_stride_words
        kk = max(k, 1)
        if out is None:
            ids = np.empty((nq, kk), np.int64)
            dist = np.empty((nq, kk), np.float32)
            counts = np.empty(nq, np.int32)
        else:
            ids, dist, counts = out
        N.check(N.lib().pkv_search(self._h, _np_ptr(queries), nq, C.byref(p), _np_ptr(ids), _np_ptr(dist),
                                   _np_ptr(counts)))
        return ids, dist, counts

    def distances(self, queries, metric: int = N.COSINE):
        """Exact distance of EVERY stored row to each query (torch CUDA queries -> [nq, rows] f32 CUDA
        tensor): the reference's untruncated scoring, input of the per-item aggregation."""
        import torch

        assert queries.is_cuda and queries.is_contiguous() and queries.dim() == 2
        if queries.shape[1] != self.dim:
            raise N.PkvError(N.ERR_DIM_MISMATCH, f"query dimension {queries.shape[1]} != index dimension {self.dim}")
        out = torch.empty((queries.shape[0], self.rows), dtype=torch.float32, device=queries.device)
        stream = torch.cuda.current_stream(queries.device).cuda_stream
        N.check(N.lib().pkv_distances_device(self._h, C.c_void_p(queries.data_ptr()), queries.shape[0], metric,
                                             _torch_code(queries), C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out

    def get_rows(self, rows):
        """Stored vectors at the given positions (torch int64 CUDA tensor) -> [n, dim] CUDA tensor."""
        import torch

        dt = {N.F32: torch.float32, N.I8: torch.int8, N.F16: torch.float16}[self.dtype]
        out = torch.empty((rows.numel(), self.dim), dtype=dt, device=rows.device)
        stream = torch.cuda.current_stream(rows.device).cuda_stream
        N.check(N.lib().pkv_index_get_rows_device(self._h, C.c_void_p(rows.data_ptr()), rows.numel(),
                                                  C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out

    def rank_groups(self, queries, group_of_row, n_groups: int, aggregation: int = N.AGG_MIN, metric: int = N.COSINE,
                    weights=None, offset: int = 0, limit: int = 320):
        """The grouped vector operator (score every row, aggregate per group, rank): torch CUDA tensors in,
        (groups[limit] int64, aggregates[limit] f64, count) CUDA tensors out; order_rank of entry i is
        offset + i + 1."""
        import torch

        assert queries.is_cuda and queries.is_contiguous() and queries.dim() == 2
        assert group_of_row.is_cuda and group_of_row.dtype == torch.int64 and group_of_row.numel() == self.rows
        p = N.RankParams(metric=metric, aggregation=aggregation, query_dtype=_torch_code(queries), offset=offset,
                         limit=limit, d_group_of_row=group_of_row.data_ptr(), n_groups=n_groups,
                         d_weights=None if weights is None else weights.data_ptr())
        groups = torch.empty(limit, dtype=torch.int64, device=queries.device)
        agg = torch.empty(limit, dtype=torch.float64, device=queries.device)
        count = torch.empty(1, dtype=torch.int32, device=queries.device)
        stream = torch.cuda.current_stream(queries.device).cuda_stream
        N.check(N.lib().pkv_rank_groups_device(self._h, C.c_void_p(queries.data_ptr()), queries.shape[0], C.byref(p),
                                               C.c_void_p(groups.data_ptr()), C.c_void_p(agg.data_ptr()),
                                               C.c_void_p(count.data_ptr()), C.c_void_p(stream)))
        return groups, agg, int(count.item())

    def _search_device(self, queries, k, metric, bitmap, bitmap_stride_words, out, stream):
        import torch

        assert queries.is_cuda and queries.is_contiguous() and queries.dim() == 2
        if queries.shape[1] != self.dim:
            raise N.PkvError(N.ERR_DIM_MISMATCH, f"query dimension {queries.shape[1]} != index dimension {self.dim}")
        nq = queries.shape[0]
        p = N.SearchParams(metric=metric, k=k, query_dtype=_torch_code(queries))
        if bitmap is not None:
            assert bitmap.is_cuda and bitmap.dtype in (torch.int64, torch.uint64) and bitmap.is_contiguous()
            p.bitmap = bitmap.data_ptr()
            p.bitmap_stride_words = bitmap_stride_words
        kk = max(k, 1)
        if out is None:
            ids = torch.empty((nq, kk), dtype=torch.int64, device=queries.device)
            dist = torch.empty((nq, kk), dtype=torch.float32, device=queries.device)
            counts = torch.empty(nq, dtype=torch.int32, device=queries.device)
        else:
            ids, dist, counts = out
        if not stream:
            stream = torch.cuda.current_stream(queries.device).cuda_stream
        N.check(N.lib().pkv_search_device(self._h, C.c_void_p(queries.data_ptr()), nq, C.byref(p),
                                          C.c_void_p(ids.data_ptr()), C.c_void_p(dist.data_ptr()),
                                          C.c_void_p(counts.data_ptr()), C.c_void_p(stream)))
        return ids, dist, counts


class ShardedIndex:
    """The corpus row-sharded over GPUs inside ONE process (pkv_sharded_*): contiguous row ranges in insertion
    order, one shard per listed device, per-shard scans run concurrently, lists merged on the first device."""

    def __init__(self, dim: int, dtype: int = N.F32, devices=(0,), total_rows: int = 0):
        self._h = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        N.check(N.lib().pkv_sharded_create(devs, len(devices), dim, dtype, C.byref(self._h)))
        self.dim, self.dtype, self.devices = dim, dtype, tuple(devices)
        if total_rows:
            self.reserve(total_rows)

    def close(self) -> None:
        if self._h:
            N.lib().pkv_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reserve(self, total_rows: int) -> None:
        N.check(N.lib().pkv_sharded_reserve(self._h, total_rows))

    def append(self, rows, row_ids=None) -> None:
        rows = np.ascontiguousarray(rows, dtype=_NP[self.dtype])
        assert rows.ndim == 2 and rows.shape[1] == self.dim
        ids = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.int64)
        N.check(N.lib().pkv_sharded_append(self._h, _np_ptr(rows), _np_ptr(ids), rows.shape[0]))

    def set_scale_artifact(self, artifact: bytes) -> None:
        buf = (C.c_uint8 * max(len(artifact), 1)).from_buffer_copy(artifact.ljust(1, b"\0"))
        N.check(N.lib().pkv_sharded_set_scale(self._h, buf, len(artifact)))

    def set_option(self, name: str, value: int) -> None:
        N.check(N.lib().pkv_sharded_set_option(self._h, name.encode(), value))

    def seal(self) -> None:
        N.check(N.lib().pkv_sharded_seal(self._h))

    @property
    def rows(self) -> int:
        out = C.c_int64()
        N.check(N.lib().pkv_sharded_rows(self._h, C.byref(out)))
        return int(out.value)

    def shard_rows(self):
        out = []
        for i in range(N.lib().pkv_sharded_shard_count(self._h)):
            h = C.c_void_p()
            N.check(N.lib().pkv_sharded_shard(self._h, i, C.byref(h)))
            info = N.IndexInfo()
            N.check(N.lib().pkv_index_get_info(h, C.byref(info)))
            out.append(int(info.rows))
        return out

    def search(self, queries: np.ndarray, k: int, metric: int = N.COSINE, bitmap=None):
        queries = np.ascontiguousarray(queries)
        if queries.ndim == 1:
            queries = queries[None, :]
        if queries.shape[1] != self.dim:
            raise N.PkvError(N.ERR_DIM_MISMATCH, f"query dimension {queries.shape[1]} != index dimension {self.dim}")
        nq = queries.shape[0]
        p = N.SearchParams(metric=metric, k=k, query_dtype=_CODE[queries.dtype])
        if bitmap is not None:
            bitmap = np.ascontiguousarray(bitmap, dtype=np.uint64)
            p.bitmap = bitmap.ctypes.data
        ids = np.empty((nq, max(k, 1)), np.int64)
        dist = np.empty((nq, max(k, 1)), np.float32)
        counts = np.empty(nq, np.int32)
        N.check(N.lib().pkv_sharded_search(self._h, _np_ptr(queries), nq, C.byref(p), _np_ptr(ids), _np_ptr(dist),
                                           _np_ptr(counts)))
        return ids, dist, counts


class Comm:
    """One process per GPU: the library's NCCL communicator (pkv_comm_*).  `unique_id()` on rank 0, distribute the
    128 bytes by any means, then Comm(device, rank, nranks, uid) on every rank."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        N.check(N.lib().pkv_comm_unique_id(buf, 128))
        return bytes(buf)

    def __init__(self, device: int, rank: int, nranks: int, uid: bytes):
        self._h = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        N.check(N.lib().pkv_comm_create(device, rank, nranks, buf, 128, C.byref(self._h)))
        self.rank, self.nranks = rank, nranks

    def close(self) -> None:
        if self._h:
            N.lib().pkv_comm_destroy(self._h)
            self._h = C.c_void_p()

    def search(self, index: "VectorIndex", queries, k: int, metric: int = N.COSINE, bitmap=None, out=None):
        """Every rank: its shard + the same torch CUDA queries -> the global top-k (CUDA tensors) on every rank.
        The exchange is enqueued on the current stream (torch orders later work on that stream behind it)."""
        import torch

        assert queries.is_cuda and queries.is_contiguous() and queries.dim() == 2
        nq = queries.shape[0]
        p = N.SearchParams(metric=metric, k=k, query_dtype=_torch_code(queries))
        if bitmap is not None:
            p.bitmap = bitmap.data_ptr()
        if out is None:
            out = (torch.empty((nq, k), dtype=torch.int64, device=queries.device),
                   torch.empty((nq, k), dtype=torch.float32, device=queries.device),
                   torch.empty(nq, dtype=torch.int32, device=queries.device))
        stream = torch.cuda.current_stream(queries.device).cuda_stream
        N.check(N.lib().pkv_search_sharded_device(index._h, self._h, C.c_void_p(queries.data_ptr()), nq, C.byref(p),
                                                  C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                                  C.c_void_p(out[2].data_ptr()), C.c_void_p(stream)))
        return out


class Corpus:
    """Lifecycle of the HBM replica of one stored payload (pkv_corpus_*): pending -> building -> ready, filled by
    idempotent chunks at one artifact_rev, kept in sync by the inline hook, usable only at the (rev, epoch) asked for."""
    PENDING, BUILDING, READY = 0, 1, 2

    def __init__(self, dim: int, dtype: int, index_db: str, space: str, profile_id: int = -1, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().pkv_corpus_create(device, dim, dtype, index_db.encode(), space.encode(), profile_id,
                                          C.byref(self._h)))
        self.dim, self.dtype = dim, dtype

    def close(self) -> None:
        if self._h:
            N.lib().pkv_corpus_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, st: int) -> None:
        if st != N.OK:
            msg = N.lib().pkv_corpus_last_error(self._h).decode("utf-8", "replace") or N.last_error()
            raise N.PkvError(st, msg)

    def begin(self, artifact_rev: int, artifact: Optional[bytes], index_epoch: int) -> None:
        a = artifact or b""
        buf = (C.c_uint8 * max(len(a), 1)).from_buffer_copy(a.ljust(1, b"\0"))
        self._check(N.lib().pkv_corpus_begin(self._h, artifact_rev, buf, len(a), index_epoch))

    def upload_chunk(self, artifact_rev: int, ids, blobs: np.ndarray):
        """-> (rows written, cursor)."""
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        blobs = np.ascontiguousarray(blobs)
        written, cursor = C.c_int64(), C.c_int64()
        self._check(N.lib().pkv_corpus_upload_chunk(self._h, artifact_rev, _np_ptr(ids), _np_ptr(blobs),
                                                    _CODE[blobs.dtype], len(ids), C.byref(written), C.byref(cursor)))
        return int(written.value), int(cursor.value)

    def append_inline(self, data_id: int, blob: np.ndarray, index_epoch: int) -> None:
        blob = np.ascontiguousarray(blob)
        self._check(N.lib().pkv_corpus_append_inline(self._h, data_id, _np_ptr(blob), blob.nbytes, _CODE[blob.dtype],
                                                     index_epoch))

    def finish(self, artifact_rev: int, index_epoch: int) -> None:
        self._check(N.lib().pkv_corpus_finish(self._h, artifact_rev, index_epoch))

    def invalidate(self) -> None:
        self._check(N.lib().pkv_corpus_invalidate(self._h))

    def info(self) -> N.CorpusInfo:
        out = N.CorpusInfo()
        self._check(N.lib().pkv_corpus_get_info(self._h, C.byref(out)))
        return out

    def ready(self, artifact_rev: int, index_epoch: int):
        """The searchable VectorIndex view (borrowed) and the ReadyPair, or None when not ready at (rev, epoch)."""
        pair = N.ReadyPair()
        h = C.c_void_p()
        st = N.lib().pkv_corpus_ready(self._h, artifact_rev, index_epoch, C.byref(pair), C.byref(h))
        if st == N.ERR_NOT_READY:
            return None
        self._check(st)
        view = VectorIndex.__new__(VectorIndex)
        view._h, view.dim, view.dtype, view.device = h, self.dim, self.dtype, 0
        view.close = lambda: None          # borrowed: the corpus owns the index
        return view, (int(pair.profile_id), float(pair.scale), int(pair.dim))


# ---- codec (db/vector_quants.rs:1446-1503) -------------------------------------

def scale_from_absmax(absmax: float) -> float:
    return float(N.lib().pkv_scale_from_absmax(C.c_float(absmax)))


def scale_artifact(scale: float) -> bytes:
    buf = (C.c_uint8 * 4)()
    N.lib().pkv_scale_artifact(C.c_float(scale), buf)
    return bytes(buf)


def artifact_scale(artifact: bytes) -> Optional[float]:
    out = C.c_float()
    buf = (C.c_uint8 * max(len(artifact), 1)).from_buffer_copy(artifact.ljust(1, b"\0"))
    st = N.lib().pkv_artifact_scale(buf, len(artifact), C.byref(out))
    return float(out.value) if st == N.OK else None


def blob_absmax(values, device: int = 0) -> float:
    out = C.c_float()
    if _is_torch(values):
        assert values.is_cuda and values.is_contiguous()
        N.check(N.lib().pkv_blob_absmax_device(device, C.c_void_p(values.data_ptr()), values.numel(), C.byref(out), None))
    else:
        values = np.ascontiguousarray(values, dtype=np.float32)
        N.check(N.lib().pkv_blob_absmax(device, _np_ptr(values), values.size, C.byref(out)))
    return float(out.value)


def quantize_int8(values, scale: float, device: int = 0):
    """GPU batch codec, bit-exact with the Rust quantize_int8."""
    if _is_torch(values):
        import torch

        assert values.is_cuda and values.is_contiguous() and values.dtype == torch.float32
        codes = torch.empty(values.shape, dtype=torch.int8, device=values.device)
        N.check(N.lib().pkv_quantize_int8_device(device, C.c_void_p(values.data_ptr()), values.numel(),
                                                 C.c_float(scale), C.c_void_p(codes.data_ptr()), None))
        return codes
    values = np.ascontiguousarray(values, dtype=np.float32)
    codes = np.empty(values.shape, dtype=np.int8)
    N.check(N.lib().pkv_quantize_int8(device, _np_ptr(values), values.size, C.c_float(scale), _np_ptr(codes)))
    return codes


def merge_topk(ids, dist, device: int = 0):
    """ids/dist: torch CUDA tensors [parts, nq, k] as gathered from the row shards."""
    import torch

    parts, nq, k = ids.shape
    o_ids = torch.empty((nq, k), dtype=torch.int64, device=ids.device)
    o_dist = torch.empty((nq, k), dtype=torch.float32, device=ids.device)
    o_cnt = torch.empty(nq, dtype=torch.int32, device=ids.device)
    stream = torch.cuda.current_stream(ids.device).cuda_stream
    N.check(N.lib().pkv_merge_topk_device(device, C.c_void_p(ids.data_ptr()), C.c_void_p(dist.data_ptr()), parts, nq, k,
                                          C.c_void_p(o_ids.data_ptr()), C.c_void_p(o_dist.data_ptr()),
                                          C.c_void_p(o_cnt.data_ptr()), C.c_void_p(stream)))
    return o_ids, o_dist, o_cnt


def pack_topk(ids, dist, out=None, device: int = 0):
    """[nq,k] int64 ids + [nq,k] f32 distances (torch CUDA) -> [nq,k,3] int32 packed entries, the payload of the
    one all-gather.  Enqueued on the current stream."""
    import torch

    nq, k = ids.shape
    if out is None:
        out = torch.empty((nq, k, 3), dtype=torch.int32, device=ids.device)
    stream = torch.cuda.current_stream(ids.device).cuda_stream
    N.check(N.lib().pkv_pack_topk_device(device, C.c_void_p(ids.data_ptr()), C.c_void_p(dist.data_ptr()), nq * k,
                                         C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
    return out


def merge_packed(packed, out=None, device: int = 0):
    """packed: [parts,nq,k,3] int32 as gathered from the row shards -> (ids, dist, counts) of the global top-k.
    Enqueued on the current stream."""
    import torch

    parts, nq, k, _ = packed.shape
    if out is None:
        out = (torch.empty((nq, k), dtype=torch.int64, device=packed.device),
               torch.empty((nq, k), dtype=torch.float32, device=packed.device),
               torch.empty(nq, dtype=torch.int32, device=packed.device))
    stream = torch.cuda.current_stream(packed.device).cuda_stream
    N.check(N.lib().pkv_merge_packed_device(device, C.c_void_p(packed.data_ptr()), parts, nq, k,
                                            C.c_void_p(out[0].data_ptr()), C.c_void_p(out[1].data_ptr()),
                                            C.c_void_p(out[2].data_ptr()), C.c_void_p(stream)))
    return out


def aggregate(dist, item_of_row, n_items: int, agg: int, weights=None, device: int = 0):
    """Per-item MIN/MAX/AVG (or weighted mean) of row distances; torch CUDA tensors."""
    import torch

    out = torch.empty(n_items, dtype=torch.float64, device=dist.device)
    stream = torch.cuda.current_stream(dist.device).cuda_stream
    N.check(N.lib().pkv_aggregate_device(device, C.c_void_p(dist.data_ptr()), C.c_void_p(item_of_row.data_ptr()),
                                         None if weights is None else C.c_void_p(weights.data_ptr()),
                                         dist.numel(), n_items, agg, C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
    return out


def fuse_ranks(lists, mode: str = "rrf", weights=None, ks=None, cap: int = 4096):
    """Rank fusion over several filters' (groups, ranks) lists (builder.rs:1284-1317).
    mode "rrf": sum_i w_i / (k_i + rank_i), best = largest; "min"/"max": coalesced min / max rank."""
    m = {"rrf": 0, "min": 1, "max": 2}[mode]
    n = len(lists)
    g_arr = [np.ascontiguousarray(g, dtype=np.int64) for g, _ in lists]
    r_arr = [np.ascontiguousarray(r, dtype=np.int64) for _, r in lists]
    gp = (C.c_void_p * n)(*[a.ctypes.data for a in g_arr])
    rp = (C.c_void_p * n)(*[a.ctypes.data for a in r_arr])
    lens = (C.c_int32 * n)(*[len(a) for a in g_arr])
    w = (C.c_double * n)(*(weights if weights is not None else [1.0] * n))
    k = (C.c_int32 * n)(*(ks if ks is not None else [60] * n))
    out_g = np.empty(cap, np.int64)
    out_s = np.empty(cap, np.float64)
    cnt = C.c_int32()
    N.check(N.lib().pkv_fuse_ranks(m, n, gp, rp, lens, w, k, _np_ptr(out_g), _np_ptr(out_s), cap, C.byref(cnt)))
    return out_g[: cnt.value], out_s[: cnt.value]
