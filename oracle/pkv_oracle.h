/*
 * pkv_oracle.h — CPU restatement of Panoptikon's vector-similarity hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or the timed CPU baseline.  The product path (panoptikon_b200/libpkv.so)
 * never links or calls this file.
 *
 * What it restates (paths relative to /root/reference):
 *   - the int8 codec: panoptikon/src/db/vector_quants.rs:1446-1503
 *   - sqlite-vec 0.1.9 vec_distance_{L2,cosine} for f32 and vec_int8 blobs.
 *     sqlite-vec is a crates.io dependency (panoptikon/Cargo.toml:84,
 *     Cargo.lock:6122-6128) whose C source is NOT vendored in the reference
 *     tree; the scalar loops below restate its published algorithm
 *     (sequential f32 accumulators, double sqrt/divide, f32 result) and are
 *     anchored on the reference's own call sites
 *     (pql/builder/filters/image_embeddings.rs:321-362,
 *      text_embeddings.rs:386-418, item_similarity.rs:503-521)
 *     and on the reference's known-answer test
 *     db/vector_quants.rs:3632-3687.
 *   - the ordering contract: ascending f32 distance,
 *     ties by ascending row, NaN last — the determinism rule the reference's
 *     own golden harness applies (pql/quant_ab.rs:32-42).
 *
 * Parity status: codec and int8 distances are PINNED by the reference's
 * known-answer tests (tests/golden/reference_kats.json); f32 distance VALUES
 * are "parity unpinned" — no reference test asserts a numeric f32
 * vec_distance result.
 */
#ifndef PKV_ORACLE_H
#define PKV_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_F32 = 0, ORC_I8 = 1, ORC_F16 = 2 };
enum { ORC_L2 = 0, ORC_COSINE = 1, ORC_DOT = 2 };
enum { ORC_AGG_MIN = 0, ORC_AGG_MAX = 1, ORC_AGG_AVG = 2 };

/* ---- int8 codec (vector_quants.rs:1446-1503) ---- */
float orc_scale_from_absmax(float absmax);
void orc_scale_artifact(float scale, uint8_t out[4]);
/* returns 1 and writes *scale when the artifact is usable, else 0 */
int orc_artifact_scale(const uint8_t *artifact, size_t len, float *scale);
float orc_blob_absmax(const uint8_t *blob, size_t len);
/* writes len/4 codes */
void orc_quantize_int8(const uint8_t *blob, size_t len, float scale, uint8_t *out);

/* ---- sqlite-vec 0.1.9 scalar distances ---- */
float orc_distance_cosine_f32(const float *a, const float *b, size_t d);
float orc_distance_l2_f32(const float *a, const float *b, size_t d);
float orc_distance_cosine_i8(const int8_t *a, const int8_t *b, size_t d);
float orc_distance_l2_i8(const int8_t *a, const int8_t *b, size_t d);
/* benchmark-only metric (BASELINE config 3): distance = 0 - dot */
float orc_distance_dot_f32(const float *a, const float *b, size_t d);
float orc_distance_dot_i8(const int8_t *a, const int8_t *b, size_t d);
/* IEEE binary16 bits -> f32 (the fp16 corpus extension widens before scoring) */
float orc_half_to_float(uint16_t h);

/*
 * Brute-force scan + top-k, one query per thread (the reference runs one
 * query per SQLite connection thread: db/connection.rs:235,328-329).
 *   corpus : n rows of `dim` components, dtype ORC_F32 / ORC_I8 / ORC_F16, row-major
 *   queries: nq rows, same dtype (F16 corpus takes F16 queries)
 *   bitmap : NULL, or LSB-first u64 words; bit r set <=> row r is a member.
 *            bitmap_stride = words per query (0: one bitmap shared by all).
 *   out_rows/out_dist: nq*k, padded with -1 / NaN; out_counts: nq
 * Returns 0, or -1 on bad arguments.
 */
int orc_topk(const void *corpus, int64_t n, int dim, int dtype,
             const void *queries, int nq, int metric, int k,
             const uint64_t *bitmap, int64_t bitmap_stride,
             int threads,
             int64_t *out_rows, float *out_dist, int32_t *out_counts);

/* All n distances of one query (for tests that need the full ordering). */
int orc_distances(const void *corpus, int64_t n, int dim, int dtype,
                  const void *query, int metric, float *out);

/*
 * Per-item aggregation of row distances (builder/filters/exact.rs:67-80):
 * MIN / MAX / AVG(d) grouped by item; weights==NULL, else SUM(d*w)/SUM(w).
 * SQLite aggregates in double and skips NULL (NaN) inputs.  Items are the
 * dense ids 0..n_items-1; out[i] is NaN for an item with no non-NULL row.
 */
int orc_aggregate(const float *dist, const int64_t *item_of_row, const float *weights,
                  int64_t n, int64_t n_items, int agg, double *out);

#ifdef __cplusplus
}
#endif
#endif
