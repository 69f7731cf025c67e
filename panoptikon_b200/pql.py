"""Host-side mirror of the PQL vector-filter surface (thin wrappers over libpkv's C++ policy
layer, csrc/pkv_policy.cpp).  Names follow the reference:
pql/builder/filters/embedding_types.rs, pql/preprocess.rs:314-465."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _native as N
from .index import VectorIndex


class PqlError(ValueError):
    """PqlError::invalid — what the server maps to HTTP 400 (api/search.rs:1664)."""


def _invalid(status: int) -> PqlError:
    return PqlError(N.last_error())


def parse_index_mode(name: str) -> int:
    out = C.c_int()
    if N.lib().pkv_parse_index_mode(name.encode(), C.byref(out)) != N.OK:
        raise _invalid(0)
    return out.value


def parse_distance_function(name: str, from_override: bool = False) -> int:
    out = C.c_int()
    if N.lib().pkv_parse_distance_function(name.encode(), int(from_override), C.byref(out)) != N.OK:
        raise _invalid(0)
    return out.value


def parse_distance_aggregation(name: str) -> int:
    out = C.c_int()
    if N.lib().pkv_parse_distance_aggregation(name.encode(), C.byref(out)) != N.OK:
        raise _invalid(0)
    return out.value


def _opt(variant: Optional[str]):
    return None if variant is None else variant.encode()


def validate_quant_args(index: int, k: int) -> None:
    if N.lib().pkv_validate_quant_args(index, k) != N.OK:
        raise _invalid(0)


def quant_requested(index: int, variant: Optional[str]) -> bool:
    return bool(N.lib().pkv_quant_requested(index, _opt(variant)))


def quant_strict(index: int, variant: Optional[str]) -> bool:
    return bool(N.lib().pkv_quant_strict(index, _opt(variant)))


@dataclass
class ReadyPair:
    """db/vector_quants.rs:1784-1788"""
    profile_id: int
    scale: float
    dim: int


class Space:
    """One setter's searchable space: exact f32 index + optional ready int8 profile(s)."""

    def __init__(self, model: str, exact: Optional[VectorIndex]):
        self._h = C.c_void_p()
        self._keep = [exact]
        N.check(N.lib().pkv_space_create(model.encode(), exact._h if exact else None, C.byref(self._h)))

    def close(self):
        if self._h:
            N.lib().pkv_space_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_quant(self, profile_name: str, pair: Optional[ReadyPair], index: Optional[VectorIndex],
                  is_default: bool = True) -> None:
        rp = None
        if pair is not None:
            rp = N.ReadyPair(profile_id=pair.profile_id, scale=pair.scale, dim=pair.dim)
        self._keep.append(index)
        N.check(N.lib().pkv_space_set_quant(self._h, profile_name.encode(), int(is_default),
                                            C.byref(rp) if rp is not None else None,
                                            index._h if index is not None else None))

    def set_modality(self, row_modality: Optional[np.ndarray]) -> None:
        """Per stored row 0 = image setter, 1 = its "t"-prefixed text sibling (db/vector_quants.rs:480-510)."""
        if row_modality is None:
            N.check(N.lib().pkv_space_set_modality(self._h, None, 0))
            return
        m = np.ascontiguousarray(row_modality, dtype=np.uint8)
        st = N.lib().pkv_space_set_modality(self._h, m.ctypes.data_as(C.c_void_p), m.size)
        if st == N.ERR_INVALID:
            raise PqlError(N.last_error())
        N.check(st)

    def search(self, queries: np.ndarray, distance_function: int, index: int = N.INDEX_AUTO,
               variant: Optional[str] = None, k: int = N.DEFAULT_K, depth: int = 100, clip_xmodal: bool = False):
        """Returns (ids, dist, counts, used_profile_id).  Validation failures raise PqlError with
        the reference's message text; device failures raise PkvError."""
        queries = np.ascontiguousarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries[None, :]
        nq, qdim = queries.shape
        d = max(1, min(depth, N.MAX_K))
        ids = np.empty((nq, d), np.int64)
        dist = np.empty((nq, d), np.float32)
        counts = np.empty(nq, np.int32)
        used = C.c_int64(-1)
        st = N.lib().pkv_space_search_xmodal(self._h, queries.ctypes.data_as(C.c_void_p), nq, qdim, distance_function,
                                             index, _opt(variant), k, depth, int(clip_xmodal),
                                             ids.ctypes.data_as(C.c_void_p), dist.ctypes.data_as(C.c_void_p),
                                             counts.ctypes.data_as(C.c_void_p), C.byref(used))
        if st in (N.ERR_INVALID, N.ERR_DIM_MISMATCH, N.ERR_NOT_READY):
            raise PqlError(N.last_error())
        N.check(st)
        return ids, dist, counts, used.value


def xmodal_text_sibling_name(model: str) -> str:
    """db/vector_quants.rs:51-53"""
    buf = C.create_string_buffer(len(model.encode()) + 2)
    N.check(N.lib().pkv_xmodal_text_sibling_name(model.encode(), buf, len(buf)))
    return buf.value.decode()


def resolve_ready_pair(pairs) -> Optional[ReadyPair]:
    """The loop of resolve_ready_pair (db/vector_quants.rs:1817-1867).  pairs: one entry per setter the query
    involves - None = no such setter (skipped), "not-ready" = the setter exists without a ready pair, or a ReadyPair.
    Returns the shared pair, or None (the `auto` fallback contract)."""
    n = len(pairs)
    arr = (N.ReadyPair * max(n, 1))()
    states = (C.c_int32 * max(n, 1))()
    for i, p in enumerate(pairs):
        if p is None:
            states[i] = 0
        elif isinstance(p, ReadyPair):
            states[i] = 1
            arr[i] = N.ReadyPair(profile_id=p.profile_id, scale=p.scale, dim=p.dim)
        else:
            states[i] = 2
    out = N.ReadyPair()
    st = N.lib().pkv_resolve_ready_pair(arr, states, n, C.byref(out))
    if st == N.ERR_NOT_READY:
        return None
    N.check(st)
    return ReadyPair(int(out.profile_id), float(out.scale), int(out.dim))


def similar_to(index: VectorIndex, target_rows, group_of_row, n_groups: int, distance_function: int = N.L2,
               distance_aggregation: int = N.AGG_AVG, weights=None, modality=None, clip_xmodal: bool = False,
               xmodal_i2i: bool = True, xmodal_t2t: bool = True, offset: int = 0, limit: int = 320):
    """`similar_to` (pql/builder/filters/item_similarity.rs:432-581) through pkv_similar_to_device: the target item's
    stored vectors (positions `target_rows`) are the query side, the aggregate (AVG by default) runs over every admitted
    (target vector, candidate vector) pair, grouped by `group_of_row` (-1 = not a candidate; the caller gives the target
    item's own rows -1).  All arrays are torch CUDA tensors.  Returns (groups, aggregates, count)."""
    import torch

    assert target_rows.is_cuda and target_rows.dtype == torch.int64 and target_rows.is_contiguous()
    assert group_of_row.is_cuda and group_of_row.dtype == torch.int64 and group_of_row.numel() == index.rows
    if target_rows.numel() == 0:
        raise PqlError("the target item has no embeddings for this model")
    p = N.SimilarParams(metric=distance_function, aggregation=distance_aggregation, offset=offset, limit=limit,
                        clip_xmodal=int(clip_xmodal), xmodal_i2i=int(xmodal_i2i), xmodal_t2t=int(xmodal_t2t),
                        n_targets=target_rows.numel(), d_target_rows=target_rows.data_ptr(),
                        d_group_of_row=group_of_row.data_ptr(), n_groups=n_groups,
                        d_modality=None if modality is None else modality.data_ptr(),
                        d_weights=None if weights is None else weights.data_ptr())
    dev = group_of_row.device
    groups = torch.empty(limit, dtype=torch.int64, device=dev)
    agg = torch.empty(limit, dtype=torch.float64, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream
    st = N.lib().pkv_similar_to_device(index._h, C.byref(p), C.c_void_p(groups.data_ptr()), C.c_void_p(agg.data_ptr()),
                                       C.c_void_p(count.data_ptr()), C.c_void_p(stream))
    if st == N.ERR_INVALID:
        raise PqlError(N.last_error())
    N.check(st)
    return groups, agg, int(count.item())
