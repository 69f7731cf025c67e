"""ctypes binding of include/pkv.h.  Loading fails loudly when libpkv.so is missing: there is
no Python or CPU fallback for any compute entry point."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PKV_LIB_PATH") or os.path.join(HERE, "libpkv.so")   # (PKV_LIB_PATH: instrumented debug builds)

OK, ERR_INVALID, ERR_DIM_MISMATCH, ERR_NOT_READY, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED = range(7)
F32, I8, F16 = 0, 1, 2
L2, COSINE, DOT = 0, 1, 2
INDEX_AUTO, INDEX_EXACT, INDEX_QUANT, INDEX_ANN = 0, 1, 2, 3
AGG_MIN, AGG_MAX, AGG_AVG = 0, 1, 2
MAX_K = 4096
DEFAULT_K = 10000


class PkvError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"pkv status {status}: {message}")
        self.status = status
        self.message = message


class IndexInfo(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("dim", C.c_int32), ("dtype", C.c_int32),
        ("sealed", C.c_int32), ("has_scale", C.c_int32), ("scale", C.c_float),
        ("rows", C.c_int64), ("capacity_rows", C.c_int64), ("device_bytes", C.c_int64), ("row_base", C.c_int64),
    ]


class SearchParams(C.Structure):
    _fields_ = [
        ("metric", C.c_int32), ("k", C.c_int32), ("query_dtype", C.c_int32), ("reserved", C.c_int32),
        ("bitmap", C.c_void_p), ("bitmap_stride_words", C.c_int64),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("searches", C.c_int64), ("queries", C.c_int64), ("kernel_launches", C.c_int64),
        ("scan_launches", C.c_int64), ("fallback_queries", C.c_int64),
        ("last_scan_ms", C.c_double), ("last_total_ms", C.c_double),
        ("last_scan_kind", C.c_int32), ("reserved", C.c_int32), ("combined_searches", C.c_int64),
        ("live_refreshes", C.c_int64), ("live_refresh_skips", C.c_int64), ("rescored_pairs", C.c_int64),
        ("deferred_pairs", C.c_int64),
    ]


class RankParams(C.Structure):
    _fields_ = [
        ("metric", C.c_int32), ("aggregation", C.c_int32), ("query_dtype", C.c_int32), ("offset", C.c_int32),
        ("limit", C.c_int32), ("reserved", C.c_int32), ("d_group_of_row", C.c_void_p), ("n_groups", C.c_int64),
        ("d_weights", C.c_void_p),
    ]


class CorpusInfo(C.Structure):
    _fields_ = [
        ("state", C.c_int32), ("dtype", C.c_int32), ("dim", C.c_int32), ("reserved", C.c_int32),
        ("profile_id", C.c_int64), ("artifact_rev", C.c_int64), ("index_epoch", C.c_uint64), ("rows", C.c_int64),
        ("cursor", C.c_int64),
    ]


class SimilarParams(C.Structure):
    _fields_ = [
        ("metric", C.c_int32), ("aggregation", C.c_int32), ("offset", C.c_int32), ("limit", C.c_int32),
        ("clip_xmodal", C.c_int32), ("xmodal_i2i", C.c_int32), ("xmodal_t2t", C.c_int32), ("n_targets", C.c_int32),
        ("d_target_rows", C.c_void_p), ("d_group_of_row", C.c_void_p), ("n_groups", C.c_int64),
        ("d_modality", C.c_void_p), ("d_weights", C.c_void_p),
    ]


class ReadyPair(C.Structure):
    _fields_ = [("profile_id", C.c_int64), ("scale", C.c_float), ("dim", C.c_int64)]


# every symbol include/pkv.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "pkv_abi_version": (C.c_int, []),
    "pkv_last_error": (C.c_char_p, []),
    "pkv_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "pkv_scale_from_absmax": (C.c_float, [C.c_float]),
    "pkv_guess_rank": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int]),
    "pkv_scale_artifact": (None, [C.c_float, _P]),
    "pkv_artifact_scale": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_float)]),
    "pkv_blob_absmax": (C.c_int, [C.c_int, _P, C.c_int64, C.POINTER(C.c_float)]),
    "pkv_quantize_int8": (C.c_int, [C.c_int, _P, C.c_int64, C.c_float, _P]),
    "pkv_blob_absmax_device": (C.c_int, [C.c_int, _P, C.c_int64, C.POINTER(C.c_float), _P]),
    "pkv_quantize_int8_device": (C.c_int, [C.c_int, _P, C.c_int64, C.c_float, _P, _P]),
    "pkv_index_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "pkv_index_destroy": (C.c_int, [_P]),
    "pkv_index_reserve": (C.c_int, [_P, C.c_int64]),
    "pkv_index_append": (C.c_int, [_P, _P, _P, C.c_int64]),
    "pkv_index_append_device": (C.c_int, [_P, _P, _P, C.c_int64]),
    "pkv_index_set_scale": (C.c_int, [_P, _P, C.c_size_t]),
    "pkv_index_set_row_base": (C.c_int, [_P, C.c_int64]),
    "pkv_index_seal": (C.c_int, [_P]),
    "pkv_index_get_info": (C.c_int, [_P, C.POINTER(IndexInfo)]),
    "pkv_search": (C.c_int, [_P, _P, C.c_int, C.POINTER(SearchParams), _P, _P, _P]),
    "pkv_search_device": (C.c_int, [_P, _P, C.c_int, C.POINTER(SearchParams), _P, _P, _P, _P]),
    "pkv_distances_device": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "pkv_rank_groups_device": (C.c_int, [_P, _P, C.c_int, C.POINTER(RankParams), _P, _P, _P, _P]),
    "pkv_index_get_rows_device": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "pkv_fuse_ranks": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "pkv_merge_topk_device": (C.c_int, [C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "pkv_pack_topk_device": (C.c_int, [C.c_int, _P, _P, C.c_int64, _P, _P]),
    "pkv_merge_packed_device": (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "pkv_sharded_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "pkv_sharded_destroy": (C.c_int, [_P]),
    "pkv_sharded_shard_count": (C.c_int, [_P]),
    "pkv_sharded_shard": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "pkv_sharded_reserve": (C.c_int, [_P, C.c_int64]),
    "pkv_sharded_append": (C.c_int, [_P, _P, _P, C.c_int64]),
    "pkv_sharded_set_scale": (C.c_int, [_P, _P, C.c_size_t]),
    "pkv_sharded_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "pkv_sharded_seal": (C.c_int, [_P]),
    "pkv_sharded_rows": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "pkv_sharded_search": (C.c_int, [_P, _P, C.c_int, C.POINTER(SearchParams), _P, _P, _P]),
    "pkv_comm_unique_id": (C.c_int, [_P, C.c_size_t]),
    "pkv_comm_create": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, C.c_size_t, C.POINTER(_P)]),
    "pkv_comm_destroy": (C.c_int, [_P]),
    "pkv_comm_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "pkv_search_sharded_device": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(SearchParams), _P, _P, _P, _P]),
    "pkv_similar_to_device": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "pkv_xmodal_text_sibling_name": (C.c_int, [C.c_char_p, C.c_char_p, C.c_size_t]),
    "pkv_resolve_ready_pair": (C.c_int, [_P, _P, C.c_int, _P]),
    "pkv_space_set_modality": (C.c_int, [_P, _P, C.c_int64]),
    "pkv_space_search_xmodal": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64, C.c_int,
                                          C.c_int, _P, _P, _P, C.POINTER(C.c_int64)]),
    "pkv_corpus_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int64, C.POINTER(_P)]),
    "pkv_corpus_destroy": (C.c_int, [_P]),
    "pkv_corpus_last_error": (C.c_char_p, [_P]),
    "pkv_corpus_begin": (C.c_int, [_P, C.c_int64, _P, C.c_size_t, C.c_uint64]),
    "pkv_corpus_upload_chunk": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int, C.c_int64, C.POINTER(C.c_int64),
                                          C.POINTER(C.c_int64)]),
    "pkv_corpus_append_inline": (C.c_int, [_P, C.c_int64, _P, C.c_size_t, C.c_int, C.c_uint64]),
    "pkv_corpus_finish": (C.c_int, [_P, C.c_int64, C.c_uint64]),
    "pkv_corpus_invalidate": (C.c_int, [_P]),
    "pkv_corpus_ready": (C.c_int, [_P, C.c_int64, C.c_uint64, _P, C.POINTER(_P)]),
    "pkv_corpus_get_info": (C.c_int, [_P, _P]),
    "pkv_sqlite_register_index": (C.c_int, [C.c_char_p, _P]),
    "pkv_aggregate_device": (C.c_int, [C.c_int, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, _P, _P]),
    "pkv_index_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "pkv_index_set_option": (C.c_int, [_P, C.c_char_p, C.c_int64]),
    "pkv_parse_index_mode": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "pkv_parse_distance_function": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "pkv_parse_distance_aggregation": (C.c_int, [C.c_char_p, C.POINTER(C.c_int)]),
    "pkv_validate_quant_args": (C.c_int, [C.c_int, C.c_int64]),
    "pkv_quant_requested": (C.c_int, [C.c_int, C.c_char_p]),
    "pkv_quant_strict": (C.c_int, [C.c_int, C.c_char_p]),
    "pkv_space_create": (C.c_int, [C.c_char_p, _P, C.POINTER(_P)]),
    "pkv_space_destroy": (C.c_int, [_P]),
    "pkv_space_set_quant": (C.c_int, [_P, C.c_char_p, C.c_int, C.POINTER(ReadyPair), _P]),
    "pkv_space_search": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int64, C.c_int,
                                   _P, _P, _P, C.POINTER(C.c_int64)]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built (python -m panoptikon_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m panoptikon_b200.build` "
                "(nvcc, sm_100a). panoptikon_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        if L.pkv_abi_version() != 1:
            raise ImportError("libpkv.so ABI version mismatch")
        _lib = L
    return _lib


def last_error() -> str:
    return lib().pkv_last_error().decode("utf-8", "replace")


def check(status: int) -> None:
    if status != OK:
        raise PkvError(status, last_error())
