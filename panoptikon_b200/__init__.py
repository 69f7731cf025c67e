"""panoptikon_b200 — B200-native (sm_100a) implementation of Panoptikon's vector-similarity hot
path: the PQL vector filters' brute-force scan + int8 codec behind a C ABI (include/pkv.h).

Everything that computes lives in libpkv.so (hand-written CUDA).  There is no CPU or Python
fallback: importing the compute wrappers without the built library raises ImportError."""
from ._native import (AGG_AVG, AGG_MAX, AGG_MIN, COSINE, DEFAULT_K, DOT, F16, F32, I8, INDEX_ANN, INDEX_AUTO,
                      INDEX_EXACT, INDEX_QUANT, L2, MAX_K, PkvError, lib)
from .index import (Comm, Corpus, ShardedIndex, VectorIndex, aggregate, artifact_scale, blob_absmax, fuse_ranks, merge_packed, merge_topk, pack_topk, quantize_int8,
                    scale_artifact, scale_from_absmax)
from .pql import (PqlError, ReadyPair, Space, parse_distance_aggregation, parse_distance_function, parse_index_mode,
                  quant_requested, quant_strict, resolve_ready_pair, similar_to, validate_quant_args,
                  xmodal_text_sibling_name)

__all__ = [n for n in dir() if not n.startswith("_")]
