#!/usr/bin/env python3
"""Writes tests/golden/reference_kats.json.

The reference is Rust + an un-vendored C crate and cannot be executed in this
container (no cargo/rustc, no sqlite-vec), so these fixtures are TRANSCRIBED from
the reference's own unit tests: the inputs are copied value for value, and the
expected outputs are either the literals the reference asserts or — where the
reference computes its expectation with an f64 formula inside the test — the same
formula evaluated here in Python floats (IEEE binary64).  Every entry cites the
reference lines it was taken from (paths relative to /root/reference/panoptikon/src).

Run from the repo root:  python tests/golden/make_reference_kats.py
"""
import json
import math
import os
import struct


def f32(x: float) -> float:
    return struct.unpack("<f", struct.pack("<f", x))[0]


def main() -> None:
    kats = {}

    # db/vector_quants.rs:3588-3626  int8_codec_rounds_ties_to_even_and_clamps
    kats["codec"] = {
        "source": "db/vector_quants.rs:3588-3626",
        "quantize": [
            {"values": [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 2.4999, -2.4999], "scale": 1.0,
             "codes": [0, 2, 2, 0, -2, -2, 2, -2]},
            {"values": [11.0, -11.0, 1000.0, -1000.0], "scale_from_absmax": 11.0,
             "codes": [127, -127, 127, -128]},
            {"values": [1.0, 2.0, 3.0], "scale": 1.0, "codes": [1, 2, 3]},
            {"values": [0.0, 0.0], "scale": 1.0, "codes": [0, 0]},
        ],
        "scale_from_absmax": [
            {"absmax": 0.0, "scale": 1.0},
            {"absmax": "nan", "scale": 1.0},
        ],
        "artifact_roundtrip_absmax": 3.5,
        "artifact_rejects_hex": ["", "0000000000", struct.pack("<f", 0.0).hex(),
                                 struct.pack("<f", -1.0).hex(), struct.pack("<f", float("nan")).hex()],
    }

    # db/vector_quants.rs:2194-2244  build_uses_absmax_scale_artifact: |s - 11/127| < 1e-6
    kats["build_scale"] = {"source": "db/vector_quants.rs:2194-2244", "absmax": 11.0,
                           "expected_scale": 11.0 / 127.0, "tol": 1e-6}

    # db/vector_quants.rs:3632-3687  sqlite_vec_int8_distances_match_a_rust_reference
    cases = [
        ([127, -128, 0, 3, -3, 64, -64, 1], [1, 2, 3, 4, 5, 6, 7, 8]),
        ([1, 1, 1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1]),
        ([-5, 20, -33, 44, 0, 12, -7, 100], [100, -7, 12, 0, 44, -33, 20, -5]),
    ]
    out = []
    for left, right in cases:
        dot = sum(float(a) * float(b) for a, b in zip(left, right))
        norm = lambda v: math.sqrt(sum(float(x) * float(x) for x in v))
        l2 = math.sqrt(sum((float(a) - float(b)) ** 2 for a, b in zip(left, right)))
        cosine = 1.0 - dot / (norm(left) * norm(right))
        out.append({"left": left, "right": right, "l2": l2, "cosine": cosine})
    kats["int8_distances"] = {"source": "db/vector_quants.rs:3632-3687", "cases": out,
                              "l2_rel_tol": 1e-4, "cosine_abs_tol": 1e-4}

    # db/vector_quants.rs:3254-3278 disagreeing_vectors + QUERY_VECTOR,
    # fillers vec8(0.5,-0.5) up to ARTIFACT_MIN_VECTORS=1024 (:34, :1947-1949, :2036-2081)
    vectors = []
    for idx in range(6):
        v = [f32(0.02)] * 8
        v[idx] = 11.0
        v[(idx + 1) % 8] = f32(f32(0.6) + f32(f32(0.5) * f32(idx)))
        vectors.append(v)
    for idx in range(6):
        v = [4.0] * 8
        for flip in range(0, (idx % 3) + 1):
            v[7 - flip] = f32(f32(-0.5) - f32(f32(0.2) * f32(idx)))
        vectors.append(v)
    kats["order_parity"] = {
        "source": "db/vector_quants.rs:3254-3278,3324-3382",
        "vectors": vectors,
        "filler": [0.5, -0.5, 1.0, -1.0, 2.0, -2.0, 3.0, -3.0],
        "total_vectors": 1024,
        "query": [1.0] * 8,
        "metric": "COSINE",
        "expect": "int8 ordering of the 12 seeded vectors == f32 ordering; repeatable; independent of k",
    }

    # db/vector_quants.rs:3532-3582 similar_to_quant_matches_exact
    sim = [[1.0, f32(f32(0.2) + f32(f32(idx) * f32(0.4))), 1.0, -1.0, 2.0, -2.0, 3.0, -3.0] for idx in range(8)]
    kats["similar_to"] = {
        "source": "db/vector_quants.rs:3532-3582",
        "vectors": sim,
        "filler": [0.5, -0.5, 1.0, -1.0, 2.0, -2.0, 3.0, -3.0],
        "total_vectors": 1024,
        "target_index": 0,
        "metric": "L2",
        "aggregation": "AVG",
        "expect": "7 results (target excluded); int8 order == f32 order",
    }

    # pql/embedding_utils.rs:377-381,427-436 : f32_1d.npy -> [0.0,1.5,-2.25,3.0] as LE f32 bytes
    kats["query_blob"] = {"source": "pql/embedding_utils.rs:15-21,377-381",
                          "values": [0.0, 1.5, -2.25, 3.0],
                          "le_hex": struct.pack("<4f", 0.0, 1.5, -2.25, 3.0).hex()}

    # pql/preprocess.rs:1268-1314 quant_policy_tests + :436-446 validate_quant_args
    kats["policy"] = {
        "source": "pql/preprocess.rs:413-446,1268-1314",
        "quant_requested": [
            {"index": "auto", "variant": None, "expect": True},
            {"index": "auto", "variant": "", "expect": True},
            {"index": "auto", "variant": "   ", "expect": True},
            {"index": "auto", "variant": "plain", "expect": True},
            {"index": "quant", "variant": None, "expect": True},
            {"index": "quant", "variant": "plain", "expect": True},
            {"index": "exact", "variant": None, "expect": False},
            {"index": "exact", "variant": "plain", "expect": False},
        ],
        "strict": [
            {"index": "auto", "variant": None, "expect": False},
            {"index": "auto", "variant": "  ", "expect": False},
            {"index": "auto", "variant": "plain", "expect": True},
            {"index": "quant", "variant": None, "expect": True},
        ],
        "errors": {
            "ann": "index \"ann\" is reserved and not yet available",
            "k": "k must be a positive integer",
        },
    }

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(path, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
