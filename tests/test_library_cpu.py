"""CPU-side checks of the C ABI library: it loads, exports every symbol include/pkv.h declares,
its host-side policy layer mirrors the reference, and compute entry points fail loudly without a GPU."""
import json
import os
import re
import struct

import numpy as np
import pytest

import panoptikon_b200 as pk
from panoptikon_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_kats.json")))


def _has_cuda():
    import torch

    return torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "pkv.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(pkv_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = pk.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in pkv.h but not exported"
    assert declared == set(N.SIGNATURES), declared ^ set(N.SIGNATURES)
    assert lib.pkv_abi_version() == 1
    assert hasattr(lib, "sqlite3_pkv_init"), "the SQLite extension entry point is not exported"


def test_sqlite_extension_loads_and_registers_functions():
    """a15: libpkv.so loads into SQLite the way sqlite-vec does (extension entry point), and the table-valued function is
    planned and reports errors as SQL errors.  Python's stdlib sqlite3 stands in for the Rust server's libsqlite3."""
    import sqlite3

    con = sqlite3.connect(":memory:")
    con.enable_load_extension(True)
    con.load_extension(N.LIB_PATH)   # default entry point of libpkv.so: sqlite3_pkv_init
    assert con.execute("SELECT pkv_version()").fetchone()[0].startswith("libpkv abi 1")
    got = con.execute("SELECT pkv_scale_from_absmax(11.0), pkv_scale_from_absmax(0.0), pkv_scale_from_absmax(NULL)").fetchone()
    assert got[0] == pk.scale_from_absmax(11.0) and got[1] == 1.0 and got[2] is None
    assert con.execute("SELECT pkv_last_execute_ms()").fetchone()[0] == 0.0
    blob = np.zeros(8, np.float32).tobytes()
    with pytest.raises(sqlite3.OperationalError, match="no resident index is registered for setter 'nope'"):
        con.execute("SELECT id, d FROM pkv_topk('nope', ?, 10)", (blob,)).fetchall()
    with pytest.raises(sqlite3.OperationalError):
        con.execute("SELECT id FROM pkv_topk('nope')").fetchall()      # the query blob is mandatory
    # the plan joins it like a table (the shape of dist_{cte}, exact.rs:106-165)
    con.execute("CREATE TABLE item_data(id INTEGER PRIMARY KEY, item_id INTEGER)")
    plan = con.execute("EXPLAIN QUERY PLAN SELECT item_data.item_id, t.d FROM pkv_topk('m', ?, 5, 'cosine') AS t "
                       "JOIN item_data ON item_data.id = t.id ORDER BY t.d", (blob,)).fetchall()
    text = " ".join(str(r) for r in plan)
    assert "VIRTUAL TABLE INDEX 15" in text.upper()          # all four arguments reached xBestIndex
    assert "USING INTEGER PRIMARY KEY" in text.upper()       # item_data is probed by id, per returned row
    assert "TEMP B-TREE" not in text.upper()                 # rows arrive best first: ORDER BY d needs no sort


def test_host_scalar_codec_matches_reference_kats():
    g = GOLD["codec"]
    assert pk.scale_from_absmax(0.0) == 1.0
    assert pk.scale_from_absmax(float("nan")) == 1.0
    assert pk.scale_from_absmax(float("inf")) == 1.0
    assert pk.scale_from_absmax(11.0) == struct.unpack("<f", struct.pack("<f", np.float32(11.0) / np.float32(127.0)))[0]
    s = pk.scale_from_absmax(g["artifact_roundtrip_absmax"])
    assert pk.artifact_scale(pk.scale_artifact(s)) == s
    for hexed in g["artifact_rejects_hex"]:
        assert pk.artifact_scale(bytes.fromhex(hexed)) is None
    assert "4 bytes" in N.last_error() or "positive finite" in N.last_error()


def test_policy_truth_tables_match_reference():
    g = GOLD["policy"]
    for row in g["quant_requested"]:
        assert pk.quant_requested(pk.parse_index_mode(row["index"]), row["variant"]) == row["expect"], row
    for row in g["strict"]:
        assert pk.quant_strict(pk.parse_index_mode(row["index"]), row["variant"]) == row["expect"], row
    with pytest.raises(pk.PqlError, match=re.escape(g["errors"]["ann"])):
        pk.validate_quant_args(pk.INDEX_ANN, 10)
    with pytest.raises(pk.PqlError, match=re.escape(g["errors"]["k"])):
        pk.validate_quant_args(pk.INDEX_AUTO, 0)
    pk.validate_quant_args(pk.INDEX_QUANT, 1)
    assert pk.DEFAULT_K == 10000


def test_enum_names_follow_serde():
    assert [pk.parse_index_mode(n) for n in ("auto", "exact", "quant", "ann")] == [0, 1, 2, 3]
    with pytest.raises(pk.PqlError):
        pk.parse_index_mode("Auto")
    assert pk.parse_distance_function("L2") == pk.L2 and pk.parse_distance_function("COSINE") == pk.COSINE
    with pytest.raises(pk.PqlError):
        pk.parse_distance_function("cosine")
    assert pk.parse_distance_function("CoSiNe", from_override=True) == pk.COSINE  # from_override lowercases
    assert [pk.parse_distance_aggregation(n) for n in ("MIN", "MAX", "AVG")] == [0, 1, 2]


@pytest.mark.skipif(_has_cuda(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu():
    with pytest.raises(pk.PkvError) as e:
        pk.VectorIndex(8)
    assert e.value.status == N.ERR_CUDA and "no CPU fallback" in e.value.message
    with pytest.raises(pk.PkvError):
        pk.quantize_int8(np.zeros(4, np.float32), 1.0)
    with pytest.raises(pk.PkvError):
        pk.blob_absmax(np.zeros(4, np.float32))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "panoptikon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f
                assert not re.search(r"#include\s+[\"<].*oracle", text), f


def test_rank_fusion_matches_the_sql_expression():
    """build_coalesced_expr (pql/builder.rs:1284-1317): RRF = sum_i 1.0/(k_i + coalesce(rank_i, HUGE)) * w_i;
    otherwise min / max over the filters of the coalesced ranks."""
    rng = np.random.default_rng(0)
    HUGE = 9223372036854775805.0
    lists, ks, ws = [], [60, 0, 10], [1.0, 0.5, 2.0]
    for _ in range(3):
        g = rng.choice(500, size=200, replace=False)
        lists.append((g, np.arange(1, 201)))
    groups = sorted(set(np.concatenate([g for g, _ in lists]).tolist()))
    rank = [{int(g): int(r) for g, r in zip(*l)} for l in lists]
    want = {g: sum(1.0 / (ks[i] + rank[i].get(g, HUGE)) * ws[i] for i in range(3)) for g in groups}
    order = sorted(groups, key=lambda g: (-want[g], g))
    got_g, got_s = pk.fuse_ranks(lists, "rrf", ws, ks)
    assert list(got_g) == order
    assert np.allclose(got_s, [want[g] for g in order], rtol=1e-15)
    mn = {g: min(rank[i].get(g, HUGE) for i in range(3)) for g in groups}
    got_g, got_s = pk.fuse_ranks(lists, "min")
    assert list(got_g) == sorted(groups, key=lambda g: (mn[g], g))
    mx = {g: max(rank[i].get(g, -HUGE) for i in range(3)) for g in groups}
    got_g, got_s = pk.fuse_ranks(lists, "max")
    assert list(got_g) == sorted(groups, key=lambda g: (-mx[g], g))


def test_cross_modal_host_policy():
    """xmodal_text_sibling_name (db/vector_quants.rs:51-53) and the sibling rule of resolve_ready_pair (:1817-1867)."""
    assert pk.xmodal_text_sibling_name("ViT-H-14-378-quickgelu/dfn5b") == "tViT-H-14-378-quickgelu/dfn5b"
    s1 = float(np.float32(0.0125))                  # scales are f32 (the 4-byte artifact)
    a = pk.ReadyPair(3, s1, 1024)
    assert pk.resolve_ready_pair([a]) == a
    assert pk.resolve_ready_pair([a, None]) == a                                   # the sibling setter does not exist: skipped
    assert pk.resolve_ready_pair([a, pk.ReadyPair(3, s1, 1024)]) == a          # siblings share one artifact
    assert pk.resolve_ready_pair([a, pk.ReadyPair(3, 0.02, 1024)]) is None         # scale differs: rebuild pending
    assert pk.resolve_ready_pair([a, pk.ReadyPair(3, s1, 768)]) is None        # dim differs
    assert pk.resolve_ready_pair([a, "not-ready"]) is None                        # an existing setter without a ready pair
    assert pk.resolve_ready_pair([None, None]) is None and pk.resolve_ready_pair([]) is None
    assert pk.resolve_ready_pair([pk.ReadyPair(3, float("nan"), 8)]) is None       # no usable scale


def test_guess_rank_bounds_the_miss_probability():
    """The guessed start takes the r-th best distance of a 3968-row sample as a query's starting
    threshold; r must make a too-tight guess rarer than guess_miss_ppm (Poisson tail) without admitting more rows than
    the candidate lists survive."""
    from scipy.stats import poisson

    L = pk.lib()
    S = 3968
    for rows, k, want in ((10_000_000, 100, 3), (1_000_000, 100, 5), (100_000, 100, None), (70_000, 10, None), (5_000_000, 1, None)):
        r = L.pkv_guess_rank(k, S, rows, 100)
        x = k * S / rows
        assert r >= 2 and poisson.sf(r - 1, x) <= 1e-4, (rows, k, r)
        assert r == 2 or poisson.sf(r - 2, x) > 1e-4, "not the smallest such rank"
        assert r * rows / S <= 16000
        if want is not None:
            assert r == want
    assert L.pkv_guess_rank(100, S, 1_000_000, 1000) < L.pkv_guess_rank(100, S, 1_000_000, 1)      # looser is safer
    assert L.pkv_guess_rank(100, S, 40_000_000, 100) == 0      # even the 2nd best of the sample admits too many rows
    assert L.pkv_guess_rank(0, S, 1000, 100) == 0 and L.pkv_guess_rank(10, 0, 1000, 100) == 0


def test_rust_ffi_binding_is_generated_from_the_header_and_complete():
    """include/pkv_ffi.rs is the `extern "C"` block the Rust server adds: generated from pkv.h
    (fresh), one `pub fn` per exported entry point, and #[repr(C)] structs whose fields are those of the ctypes
    structures the GPU tests drive the library with (same names, same order, same sizes)."""
    import ctypes as C
    import subprocess
    import sys

    rc = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_rust_ffi.py"), "--check"]).returncode
    assert rc == 0, "include/pkv_ffi.rs is stale: run tools/gen_rust_ffi.py"
    rs = open(os.path.join(ROOT, "include", "pkv_ffi.rs")).read()
    fns = set(re.findall(r"pub fn ([a-z0-9_]+)\(", rs))
    assert fns == set(N.SIGNATURES) | {"sqlite3_pkv_init"}, fns ^ (set(N.SIGNATURES) | {"sqlite3_pkv_init"})
    size = {"i8": 1, "u8": 1, "i32": 4, "u32": 4, "c_int": 4, "f32": 4, "i64": 8, "u64": 8, "f64": 8, "usize": 8}
    pairs = {"PkvIndexInfo": N.IndexInfo, "PkvSearchParams": N.SearchParams, "PkvCounters": N.Counters,
             "PkvRankParams": N.RankParams, "PkvCorpusInfo": N.CorpusInfo, "PkvSimilarParams": N.SimilarParams,
             "PkvReadyPair": N.ReadyPair}
    for rname, ct in pairs.items():
        body = re.search(r"pub struct %s \{(.*?)\n\}" % rname, rs, flags=re.S).group(1)
        fields = re.findall(r"pub ([a-z0-9_]+): ([^,]+),", body)
        assert [f for f, _ in fields] == [f[0] for f in ct._fields_], rname
        for (fname, rtype), cf in zip(fields, ct._fields_):
            want = C.sizeof(cf[1])
            got = 8 if rtype.startswith("*") else size[rtype]
            assert got == want, (rname, fname, rtype)
