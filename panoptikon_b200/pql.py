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

    def search(self, queries: np.ndarray, distance_function: int, index: int = N.INDEX_AUTO,
               variant: Optional[str] = None, k: int = N.DEFAULT_K, depth: int = 100):
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
        st = N.lib().pkv_space_search(self._h, queries.ctypes.data_as(C.c_void_p), nq, qdim, distance_function, index,
                                      _opt(variant), k, depth, ids.ctypes.data_as(C.c_void_p),
                                      dist.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p),
                                      C.byref(used))
        if st in (N.ERR_INVALID, N.ERR_DIM_MISMATCH, N.ERR_NOT_READY):
            raise PqlError(N.last_error())
        N.check(st)
        return ids, dist, counts, used.value


def similar_to(index: VectorIndex, item_of_row, n_items: int, target_item: int, distance_function: int,
               distance_aggregation: int = N.AGG_AVG, weights=None, offset: int = 0, limit: int = 320):
    """`similar_to` (pql/builder/filters/item_similarity.rs:432-581): the target item's stored vectors are the
    queries, every other item's vectors the candidates, the aggregate (AVG by default,
    item_similarity.rs:127-130) runs over all (target vector, candidate vector) pairs, the target itself
    is excluded.  item_of_row: torch int64 CUDA tensor, dense item index per stored row.
    Returns (items, aggregates, count)."""
    import torch

    target_rows = torch.nonzero(item_of_row == target_item).flatten()
    if target_rows.numel() == 0:
        raise PqlError(f"item {target_item} has no embeddings for this model")
    queries = index.get_rows(target_rows.contiguous())
    groups = torch.where(item_of_row == target_item, torch.full_like(item_of_row, -1), item_of_row).contiguous()
    return index.rank_groups(queries, groups, n_items, distance_aggregation, distance_function, weights, offset, limit)
