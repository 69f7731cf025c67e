/*
 * pkv.h — C ABI of the B200-native vector-similarity operator for Panoptikon.
 *
 * This is the drop-in boundary for ONE path of reasv/panoptikon: the PQL vector
 * filters' brute-force scan, which the reference executes row-at-a-time inside
 * SQLite through sqlite-vec's vec_distance_L2 / vec_distance_cosine / vec_int8
 * plus a small Rust int8 codec.  Paths below are relative to
 * /root/reference/panoptikon/src unless they start with docs/.
 *
 * Conventions
 *   - plain C types only; every function returns a pkv_status (0 = ok) unless
 *     stated otherwise; the message for the calling thread's last failure is
 *     pkv_last_error().  No C++ exception or abort crosses this boundary.
 *   - inputs are borrowed for the duration of the call (the Rust side hands
 *     `&[u8]` / `Vec<u8>` blobs: pql/embedding_utils.rs:15-21,
 *     builder/filters/embedding_types.rs:71-76); outputs are caller-allocated.
 *   - blobs use the reference's storage layout: one vector = D little-endian
 *     f32 (`embeddings.embedding`, migrations/index/20250117193000_init.sql:29-33)
 *     or D int8 codes (`embedding_quants.quant`,
 *     migrations/index/20260730150000_embedding_quants_rowid.sql:32-48), no header.
 *   - all entry points are re-entrant; searches may run concurrently from many
 *     threads (the reference's read pool is 16 threads: db/connection.rs:235),
 *     append/seal take the index exclusively.  Concurrent small pkv_search calls with the
 *     same metric/k are combined into one pass over the corpus (no added wait when idle).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     PKV_ERR_CUDA when no sm_100 device is usable.
 *
 * Ordering contract (the reference leaves ties to SQLite's plan order,
 * builder.rs:1188-1205; its own golden harness appends a stable key,
 * pql/quant_ab.rs:32-42): ascending f32 distance, ties by ascending row id
 * position (insertion order), NaN distances (SQL NULL, "NULLS LAST") last.
 */
#ifndef PKV_H
#define PKV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PKV_ABI_VERSION 1

typedef enum {
    PKV_OK = 0,
    PKV_ERR_INVALID = 1,      /* bad argument (maps to PqlError::invalid -> HTTP 400, api/search.rs:1664) */
    PKV_ERR_DIM_MISMATCH = 2, /* pql/preprocess.rs:372-381 */
    PKV_ERR_NOT_READY = 3,    /* quant pair not ready (db/vector_quants.rs:1829-1850) / index not sealed */
    PKV_ERR_CUDA = 4,         /* device failure (maps to ApiError::internal, db/pql.rs:18-21) */
    PKV_ERR_OOM = 5,
    PKV_ERR_UNSUPPORTED = 6
} pkv_status;

/* element type of the stored rows */
typedef enum {
    PKV_F32 = 0, /* embeddings.embedding: D x LE f32 */
    PKV_I8 = 1,  /* embedding_quants.quant: D x int8 (vec_int8 blobs, docs/vector-int8-quant.md:11-49) */
    PKV_F16 = 2  /* framework extension (BASELINE config 4); the reference always stores f32 */
} pkv_dtype;

/* distance function: builder/filters/embedding_types.rs:20-42 (serde names "L2"/"COSINE").
 * PKV_DOT (distance = -dot) is the benchmark-only metric of BASELINE config 3. */
typedef enum { PKV_L2 = 0, PKV_COSINE = 1, PKV_DOT = 2 } pkv_metric;

typedef struct pkv_index pkv_index; /* opaque; owned by the library */

typedef struct {
    int32_t abi_version;
    int32_t device;
    int32_t dim;
    int32_t dtype;    /* pkv_dtype */
    int32_t sealed;   /* 1 after pkv_index_seal: searchable */
    int32_t has_scale;
    float scale;      /* int8 scale (valid when has_scale) */
    int64_t rows;
    int64_t capacity_rows;
    int64_t device_bytes; /* HBM held by the corpus + per-row side arrays */
    int64_t row_base;
} pkv_index_info;

/* -- library ---------------------------------------------------------------- */
int pkv_abi_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
const char *pkv_last_error(void);
/* Number of usable CUDA devices; 0 (and PKV_ERR_CUDA) when there is none. */
int pkv_device_count(int *count);

/* -- int8 codec: db/vector_quants.rs:1446-1503 ------------------------------- */
/* scale_from_absmax (:1465-1471): absmax/127, or 1.0 when absmax is not a positive finite value. */
float pkv_scale_from_absmax(float absmax);
/* scale_artifact (:1449-1451): the 4-byte little-endian f32 payload. */
void pkv_scale_artifact(float scale, uint8_t out[4]);
/* artifact_scale (:1456-1460): PKV_OK and *scale, or PKV_ERR_INVALID for anything that
 * is not exactly 4 bytes holding a positive finite f32. */
int pkv_artifact_scale(const uint8_t *artifact, size_t len, float *scale);
/* blob_absmax over `n` LE f32 components in HOST memory, reduced on the GPU
 * (:1474-1483; NaN never replaces the running max). */
int pkv_blob_absmax(int device, const float *values, int64_t n, float *absmax);
/* quantize_int8 / compute_query_quant (:1489-1503) for `n` components in HOST memory,
 * computed on the GPU, bit-exact with the Rust codec:
 * clamp(round_ties_even(x / scale), -128, 127) as i8, NaN -> 0. */
int pkv_quantize_int8(int device, const float *values, int64_t n, float scale, int8_t *codes);
/* Same two operations on DEVICE pointers (bulk corpus build, replaces the 5000-row
 * backfill chunks of db/vector_quants.rs:1085-1163).  stream is a cudaStream_t or NULL. */
int pkv_blob_absmax_device(int device, const float *d_values, int64_t n, float *absmax, void *stream);
int pkv_quantize_int8_device(int device, const float *d_values, int64_t n, float scale, int8_t *d_codes,
                             void *stream);

/* -- index lifecycle --------------------------------------------------------- */
/* One index = one (setter space, payload) pair resident in one GPU's HBM: the
 * `embeddings` rows of a setter (PKV_F32) or its `embedding_quants` rows for one
 * (profile_id, artifact_rev) (PKV_I8). */
int pkv_index_create(int device, int dim, int dtype, pkv_index **out);
int pkv_index_destroy(pkv_index *h);
int pkv_index_reserve(pkv_index *h, int64_t rows);
/* Appends `n` rows given as concatenated blobs exactly as SQLite stores them
 * (n * dim * elem_size bytes, HOST memory).  row_ids: the n `item_data.id` values
 * (embeddings.id), or NULL for row_base + position.  Copies; never retains pointers. */
int pkv_index_append(pkv_index *h, const void *rows, const int64_t *row_ids, int64_t n);
/* Same with DEVICE pointers (rows and row_ids already in this GPU's HBM). */
int pkv_index_append_device(pkv_index *h, const void *d_rows, const int64_t *d_row_ids, int64_t n);
/* int8 only: the frozen scale artifact of the pair (vector_quant_coverage.artifact,
 * migrations/index/20260720130000_vector_quants.sql:12-29).  Rejected like artifact_scale. */
int pkv_index_set_scale(pkv_index *h, const uint8_t *artifact, size_t len);
/* First global row number of this shard when the corpus is row-sharded over GPUs;
 * only affects ids reported for rows appended with row_ids == NULL. */
int pkv_index_set_row_base(pkv_index *h, int64_t row_base);
/* Makes appended rows searchable (computes the per-row side arrays). Idempotent;
 * appending after a seal un-seals the new rows until the next seal. */
int pkv_index_seal(pkv_index *h);
int pkv_index_get_info(const pkv_index *h, pkv_index_info *info);

/* -- search: replaces the materialised distance CTE of builder/filters/exact.rs:106-165
 *    (`dist_{cte}(.., d)` = vec_distance_*(stored, ?) over every candidate row),
 *    truncated to retrieval depth k — the reserved `index`/`k` seam of
 *    builder/filters/embedding_types.rs:44-66 / docs/vector-int8-quant.md:86-88. */
typedef struct {
    int32_t metric;       /* pkv_metric */
    int32_t k;            /* retrieval depth, 1..PKV_MAX_K */
    int32_t query_dtype;  /* PKV_F32, or PKV_I8 codes (QuantResolved.query_quant) for an int8
                             index, or PKV_F16 for an f16 index.  f32 queries against an int8
                             index are quantised on the GPU with the index scale
                             (compute_query_quant, db/vector_quants.rs:1501-1503). */
    int32_t reserved;
    /* Optional membership filter ("∩ context", image_embeddings.rs:140-199): LSB-first
     * u64 words, bit r set <=> stored row r may be returned.  NULL = all rows.
     * bitmap_stride_words = 0: one bitmap for every query; else words per query. */
    const uint64_t *bitmap;
    int64_t bitmap_stride_words;
} pkv_search_params;

#define PKV_MAX_K 4096 /* the server clamps LIMIT to 4096 (api/search.rs:51) */

/* HOST buffers.  queries: nq blobs of dim components.  out_ids / out_dist: nq*k,
 * best first; out_counts: nq (number of valid entries; the tail is id -1 / NaN). */
int pkv_search(pkv_index *h, const void *queries, int nq, const pkv_search_params *params,
               int64_t *out_ids, float *out_dist, int32_t *out_counts);
/* DEVICE buffers (queries, bitmap and outputs in this GPU's HBM), work enqueued on
 * `stream` (cudaStream_t or NULL); returns after the results are complete. */
int pkv_search_device(pkv_index *h, const void *d_queries, int nq, const pkv_search_params *params,
                      int64_t *d_out_ids, float *d_out_dist, int32_t *d_out_counts, void *stream);

/* Untruncated scoring: the exact distance of every stored row to each of nq queries, i.e. the whole
 * materialised `dist_{cte}` of builder/filters/exact.rs:106-165 (the reference scores every
 * candidate and only LIMITs at the very end, docs/vector-index-design.md:84-89).
 * d_out: [nq][rows] f32 in DEVICE memory, row order = stored order; feed it to
 * pkv_aggregate_device for the per-item MIN/MAX/AVG of builder/filters/exact.rs:67-80.
 * Meant for interactive batches (nq <= 1024 per call). */
int pkv_distances_device(pkv_index *h, const void *d_queries, int nq, int metric, int query_dtype, float *d_out,
                         void *stream);

/* The grouped vector operator: score EVERY stored row against nq query vectors, aggregate per group,
 * rank the groups — what the filter compilers render as
 *   SELECT ..., row_number() OVER (ORDER BY AGG(d) ASC) AS order_rank FROM dist_{cte} GROUP BY file_id
 * (builder/filters/exact.rs:67-80,106-165; builder.rs:757-771).  nq = 1 for image_embeddings /
 * text_embeddings; nq = the target item's own vectors for similar_to, whose AGG runs over every
 * (target vector, candidate vector) pair and whose target rows are excluded by group -1
 * (pql/builder/filters/item_similarity.rs:432-581).
 *   d_group_of_row : per stored row, the dense group index (file / item, or text data row), or -1
 *                    when the row is not a candidate (context CTE, the target itself)
 *   d_weights      : NULL, or per-row w = pow(coalesce(conf,1),wc)*pow(coalesce(lang_conf,1),wl)
 *                    (exact.rs:37-60) => SUM(d*w)/SUM(w) instead of `aggregation`
 *   offset / limit : pagination of the ranked list (builder.rs:578-582); offset+limit <= 2048
 * Output (DEVICE): out_groups[limit] (-1 padded), out_agg[limit] (the aggregate, NaN = SQL NULL;
 * order_rank of entry i is offset+i+1), *out_count.  Groups without any candidate row are absent;
 * NULL aggregates rank last; ties by ascending group index. */
typedef struct {
    int32_t metric;       /* pkv_metric */
    int32_t aggregation;  /* pkv_distance_aggregation */
    int32_t query_dtype;  /* as pkv_search_params */
    int32_t offset;
    int32_t limit;
    int32_t reserved;
    const int64_t *d_group_of_row;
    int64_t n_groups;
    const float *d_weights;
} pkv_rank_params;
int pkv_rank_groups_device(pkv_index *h, const void *d_queries, int nq, const pkv_rank_params *params,
                           int64_t *d_out_groups, double *d_out_agg, int32_t *d_out_count, void *stream);
/* similar_to (pql/builder/filters/item_similarity.rs:83-142 arguments, :432-581 the rendered self-join): the stored
 * vectors of the TARGET item are the query side ("main_embeddings"), every other stored row a candidate
 * ("other_embeddings"); AGG(d) - or SUM(d*w)/SUM(w) with w = pow(conf_m*conf_o, wc) * pow(lang_m*lang_o, wl), which
 * factorises into a per-row weight on each side - runs over every admitted (target vector, candidate vector) pair,
 * grouped by the candidate's group (file / data row), ranked ascending.
 *   d_target_rows : positions of the target item's rows in this index (they are read on the device: no host copy)
 *   d_group_of_row: dense group per stored row, -1 = not a candidate; the caller gives the target item's own rows -1
 *                   (`other.sha256 != target`)
 *   d_modality    : per stored row 0 = image setter ("clip"), 1 = its "t"-prefixed text sibling ("text-embedding");
 *                   one index holds the whole space (db/vector_quants.rs:480-510).  NULL = a plain single-setter index.
 *   clip_xmodal=0 : only image rows take part on either side; =1: all pairs except image-image when !xmodal_i2i and
 *                   text-text when !xmodal_t2t (:468-488)
 *   d_weights     : per stored row pow(coalesce(conf,1),wc)*pow(coalesce(lang_conf,1),wl), or NULL */
typedef struct {
    int32_t metric;       /* PKV_L2 (the default of SimilarityArgs) or PKV_COSINE */
    int32_t aggregation;  /* pkv_distance_aggregation; similar_to defaults to AVG */
    int32_t offset;
    int32_t limit;
    int32_t clip_xmodal;
    int32_t xmodal_i2i;
    int32_t xmodal_t2t;
    int32_t n_targets;
    const int64_t *d_target_rows;
    const int64_t *d_group_of_row;
    int64_t n_groups;
    const uint8_t *d_modality;
    const float *d_weights;
} pkv_similar_params;
int pkv_similar_to_device(pkv_index *h, const pkv_similar_params *params, int64_t *d_out_groups, double *d_out_agg,
                          int32_t *d_out_count, void *stream);

/* Copies stored rows (by position) into d_out[n][dim] in the index dtype: similar_to reads the target's
 * own stored vectors as its queries (item_similarity.rs:352-426). */
int pkv_index_get_rows_device(pkv_index *h, const int64_t *d_rows, int n, void *d_out, void *stream);

/* Rank fusion of several filters' ranked lists over the same group ids (HOST, lists are <= LIMIT long):
 * mode 0: RRF, score = SUM_i weight_i * 1.0 / (k_i + coalesce(rank_i, 9223372036854775805)), best = largest
 *         (build_coalesced_expr, builder.rs:1284-1302);
 * mode 1 / 2: min / max over the filters of coalesce(rank_i, very large / very small) (builder.rs:1304-1317).
 * Writes the union of groups ordered best first (ties by group id). */
int pkv_fuse_ranks(int mode, int n_lists, const int64_t *const *groups, const int64_t *const *ranks, const int32_t *lens,
                   const double *weights, const int32_t *ks, int64_t *out_groups, double *out_scores, int32_t cap,
                   int32_t *out_count);

/* Merge `parts` per-shard result lists (each nq*k, laid out [part][nq][k], as gathered
 * by one NCCL all-gather) into the global top-k under the same total order.
 * DEVICE buffers. */
int pkv_merge_topk_device(int device, const int64_t *d_ids, const float *d_dist, int parts, int nq, int k,
                          int64_t *d_out_ids, float *d_out_dist, int32_t *d_out_counts, void *stream);

/* The same exchange without intermediate copies: pack a shard's nq*k results into 12-byte entries
 * {id low word, id high word, distance bits} (the payload of the ONE all-gather), and merge the gathered
 * [parts][nq][k] packed buffer directly.  Both only ENQUEUE on `stream` (no host synchronisation). */
int pkv_pack_topk_device(int device, const int64_t *d_ids, const float *d_dist, int64_t n, void *d_packed, void *stream);
int pkv_merge_packed_device(int device, const void *d_packed, int parts, int nq, int k, int64_t *d_out_ids,
                            float *d_out_dist, int32_t *d_out_counts, void *stream);

/* -- sharding: the corpus row-sharded over the GPUs of one box (the reference has no multi-device path) ----
 * Contiguous row ranges by insertion order, one shard per listed device; every shard scans for the same batch and the
 * per-shard top-k lists are merged under the same total order (distance, then global row).  Two shapes:
 *
 * (1) ONE process, S shards - what the single-process Rust server binds.  pkv_sharded_search is a drop-in for
 *     pkv_search: host buffers in, global top-k out; the per-shard scans run concurrently, the lists are gathered on the
 *     first device and merged by the merge kernel.  The same device may be listed several times (tests). */
typedef struct pkv_sharded pkv_sharded;
int pkv_sharded_create(const int *devices, int n_shards, int dim, int dtype, pkv_sharded **out);
int pkv_sharded_destroy(pkv_sharded *h);
int pkv_sharded_shard_count(const pkv_sharded *h);
/* Borrows shard i (options, counters, info); the handle stays owned by the sharded index. */
int pkv_sharded_shard(pkv_sharded *h, int i, pkv_index **out);
/* Fixes the shard boundaries: ceil(total_rows / S) rows per shard, rounded up to a multiple of 64 so that a global
 * membership bitmap slices on word boundaries; shard i reports rows i*per .. as its ids (or the appended row_ids). */
int pkv_sharded_reserve(pkv_sharded *h, int64_t total_rows);
/* Appends in global insertion order; rows spill over into the next shard at the boundary. */
int pkv_sharded_append(pkv_sharded *h, const void *rows, const int64_t *row_ids, int64_t n);
int pkv_sharded_set_scale(pkv_sharded *h, const uint8_t *artifact, size_t len);
int pkv_sharded_set_option(pkv_sharded *h, const char *name, int64_t value);
int pkv_sharded_seal(pkv_sharded *h);
int pkv_sharded_rows(const pkv_sharded *h, int64_t *rows);
/* As pkv_search.  params->bitmap: ONE bitmap over the global rows (bitmap_stride_words must be 0). */
int pkv_sharded_search(pkv_sharded *h, const void *queries, int nq, const pkv_search_params *params, int64_t *out_ids,
                       float *out_dist, int32_t *out_counts);

/* (2) ONE process per GPU (torchrun / MPI style): a communicator over NCCL, resolved at run time (dlopen of
 *     libnccl.so.2; PKV_ERR_UNSUPPORTED when absent).  Rank 0 makes the 128-byte id, the host distributes it (a file,
 *     an env var, a torch.distributed broadcast), every rank calls pkv_comm_create. */
typedef struct pkv_comm pkv_comm;
int pkv_comm_unique_id(uint8_t *out, size_t len /* 128 */);
int pkv_comm_create(int device, int rank, int nranks, const uint8_t *unique_id, size_t len /* 128 */, pkv_comm **out);
int pkv_comm_destroy(pkv_comm *h);
int pkv_comm_info(const pkv_comm *h, int *rank, int *nranks);
/* Every rank calls this with ITS shard and the SAME queries (device buffers); every rank ends up with the global
 * top-k.  The exchange is ONE ncclAllGather of the packed per-shard lists (nq*k*12 bytes per rank) enqueued behind the
 * scan on `stream`, then the merge kernel reading the gathered buffer.  The shard's scan has completed on return; the
 * exchange is only enqueued: the outputs are complete once `stream` is synchronised.  Collective: all ranks, same
 * order; one call at a time per communicator and stream. */
int pkv_search_sharded_device(pkv_index *shard, pkv_comm *comm, const void *d_queries, int nq,
                              const pkv_search_params *params, int64_t *d_out_ids, float *d_out_dist, int32_t *d_out_counts,
                              void *stream);

/* -- SQLite seam (db/sql_functions.rs:105-128 registers sqlite-vec the same way) ---------------------------------------
 * sqlite3_pkv_init is an extension entry point - int(sqlite3*, char**, const sqlite3_api_routines*) - to hand to
 * sqlite3_auto_extension() next to / instead of sqlite3_vec_init, or to load_extension().  Every connection then has
 *   pkv_topk(model, query_blob, k [, metric])  table-valued: rows (id = item_data.id, d = distance, rank), best first -
 *                                              substitutable for the `dist_{cte}` CTE of
 *                                              builder/filters/exact.rs:106-165 truncated to retrieval depth k
 *   pkv_version(), pkv_scale_from_absmax(x), pkv_last_execute_ms()
 * `model` (the setter name the filter compilers bind, image_embeddings.rs:140-199) selects an index the host registered
 * with pkv_sqlite_register_index (NULL handle = withdraw).  The index stays owned by the caller. */
int pkv_sqlite_register_index(const char *model, pkv_index *h);
int sqlite3_pkv_init(void *db, char **pzErrMsg, const void *pApi);

/* Per-item aggregation of row distances (builder/filters/exact.rs:67-80):
 * agg 0 MIN / 1 MAX / 2 AVG of d grouped by item, or SUM(d*w)/SUM(w) when d_weights
 * is non-NULL.  item ids are dense 0..n_items-1; NaN rows are skipped (SQL NULL);
 * out[i] = NaN for items without rows.  DEVICE buffers. */
int pkv_aggregate_device(int device, const float *d_dist, const int64_t *d_item_of_row, const float *d_weights,
                         int64_t n, int64_t n_items, int agg, double *d_out, void *stream);

/* -- counters for the bench (cumulative since create) ------------------------- */
typedef struct {
    int64_t searches;
    int64_t queries;
    int64_t kernel_launches;    /* our own kernels launched by searches */
    int64_t scan_launches;      /* the dominant (scan) kernels among them */
    int64_t fallback_queries;   /* queries re-run on the small-chunk schedule after a candidate overflow */
    double last_scan_ms;        /* CUDA-event time of the scan kernels of the last search on this thread */
    double last_total_ms;       /* CUDA-event time of the last search's device work */
    int32_t last_scan_kind;     /* which scan kernel family ran: 1 simt-f32, 2 simt-i8, 3 tc-i8, 4 tc-tf32, 5 simt-f16,
                                   6 tc-f16 (f16 rows), 7 tc-f16 on the fp16 image of f32 rows,
                                   8 tc-i8 on the int8 image of f32/f16 rows (pkv_scan_img8.cu) */
    int32_t reserved;
    int64_t combined_searches;  /* host searches that shared one corpus scan with concurrent callers */
    int64_t live_refreshes;     /* in-kernel threshold selections (live mode) */
    int64_t live_refresh_skips; /* ... that gave up because too many keys were still live */
    int64_t rescored_pairs;     /* (row, query) pairs re-scored exactly inside the int8-image scan kernel */
    int64_t deferred_pairs;     /* ... parked by a live launch and re-scored behind it, after the final-threshold check */
} pkv_counters;
int pkv_index_counters(pkv_index *h, pkv_counters *out);
/* Options of an index.  Production: "image_mask" (which filter images an f32/f16 index builds at seal: bit 1 int8,
 * bit 0 fp16; set before the first append), "use_shadow" (which image searches scan: -1 best available, 2 int8, 1 fp16,
 * 0 none), "combine" (merge concurrent small host searches into one scan), "live" (0 never / 1 auto / 2 always: one
 * launch over the bulk of a large corpus with thresholds maintained in-kernel), "time_kernels".  Everything else the
 * call accepts is a test / experiment switch that forces a particular kernel or schedule so that the parity tests can
 * compare them; those are listed in INTEGRATION.md section 4 and are not part of the supported surface.  Unknown names
 * are PKV_ERR_INVALID. */
int pkv_index_set_option(pkv_index *h, const char *name, int64_t value);
/* The sample rank a guessed start uses: with x = k * sample_rows / rows sample rows expected among
 * the corpus' true top-k, the smallest rank r >= 2 with P[Poisson(x) >= r] <= miss_ppm * 1e-6 - the r-th best sample
 * distance is then too tight a threshold for a query with at most that probability; 0 when a guess would admit more rows
 * than the candidate lists survive (the search then learns its thresholds on a chunked prefix).  Pure host arithmetic. */
int pkv_guess_rank(int k, int64_t sample_rows, int64_t rows, int miss_ppm);

/* -- PQL operator policy: pql/preprocess.rs:314-465, builder/filters/embedding_types.rs --- */
typedef enum { PKV_INDEX_AUTO = 0, PKV_INDEX_EXACT = 1, PKV_INDEX_QUANT = 2, PKV_INDEX_ANN = 3 } pkv_index_mode;
typedef enum { PKV_AGG_MIN = 0, PKV_AGG_MAX = 1, PKV_AGG_AVG = 2 } pkv_distance_aggregation;
#define PKV_DEFAULT_K 10000 /* default_k(), embedding_types.rs:64-66 */

/* serde names -> enums; PKV_ERR_INVALID for unknown names.
 * IndexMode: "auto"/"exact"/"quant"/"ann"; DistanceFunction: "L2"/"COSINE"
 * (from_override, embedding_types.rs:28-36, is case-insensitive); aggregation "MIN"/"MAX"/"AVG". */
int pkv_parse_index_mode(const char *name, int *mode);
int pkv_parse_distance_function(const char *name, int case_insensitive, int *metric);
int pkv_parse_distance_aggregation(const char *name, int *agg);
/* validate_quant_args (preprocess.rs:436-446) */
int pkv_validate_quant_args(int index_mode, int64_t k);
/* quant_requested (preprocess.rs:413-421): 1/0 */
int pkv_quant_requested(int index_mode, const char *variant_or_null);
/* strict = index == quant || normalize_variant(variant).is_some() (preprocess.rs:331-332,425-431) */
int pkv_quant_strict(int index_mode, const char *variant_or_null);

/* A ReadyPair (db/vector_quants.rs:1784-1788) as the operator needs it. */
typedef struct {
    int64_t profile_id;
    float scale;
    int64_t dim;
} pkv_ready_pair;

/* A searchable "space": the exact f32 index of a setter and, optionally, the int8 index
 * of its ready quant profile.  Mirrors what resolve_vector_quant + the filter compilers
 * decide per request (preprocess.rs:314-393; image_embeddings.rs:321-362):
 *   exact            -> f32 index
 *   auto             -> quant index of the default profile when a ready pair exists and the
 *                       query dimension matches, else exact (silent fallback)
 *   quant / variant  -> quant index or an error (never a silent fallback)
 *   ann              -> rejected
 * Neither index is owned by the space. */
typedef struct pkv_space pkv_space;
int pkv_space_create(const char *model, pkv_index *exact_f32, pkv_space **out);
int pkv_space_destroy(pkv_space *s);
/* Registers (or, with quant == NULL, withdraws) the ready pair of profile `profile_name`;
 * is_default marks it as the configured default profile. */
int pkv_space_set_quant(pkv_space *s, const char *profile_name, int is_default, const pkv_ready_pair *pair,
                        pkv_index *quant_i8);
/* One vector filter evaluation. query: nq f32 blobs of query_dim components (HOST).
 * *used_profile_id = the profile searched, or -1 when the exact index was used. */
int pkv_space_search(pkv_space *s, const float *queries, int nq, int query_dim, int metric, int index_mode,
                     const char *variant_or_null, int64_t k_arg, int depth, int64_t *out_ids, float *out_dist,
                     int32_t *out_counts, int64_t *used_profile_id);

/* -- corpus lifecycle: the HBM replica of one stored payload (db/vector_quants.rs:1085-1163 chunked backfill,
 *    :1347-1438 inline hook, :1829-1850 readiness; db/epochs.rs:38-44 invalidation) ----------------------------------
 * A pkv_corpus is the reference's pair state machine (pending -> building -> ready) around a pkv_index, keyed by
 * (index_db, space, profile_id): filled by idempotent chunks at one artifact_rev, kept in sync by the inline hook,
 * usable only while ready at the (artifact_rev, index epoch) the host reads from its DB.  f32 source blobs of an int8
 * corpus are quantised on the GPU with the frozen scale (quantize_int8 of backfill_chunk / write_inline_quants). */
typedef struct pkv_corpus pkv_corpus;
typedef enum { PKV_CORPUS_PENDING = 0, PKV_CORPUS_BUILDING = 1, PKV_CORPUS_READY = 2 } pkv_corpus_state;
typedef struct {
    int32_t state; /* pkv_corpus_state */
    int32_t dtype;
    int32_t dim;
    int32_t reserved;
    int64_t profile_id;   /* -1: the exact f32 payload (`embeddings`) */
    int64_t artifact_rev;
    uint64_t index_epoch;
    int64_t rows;
    int64_t cursor;       /* largest item_data.id uploaded by chunks: the `d.id > ?` resume point of the backfill */
} pkv_corpus_info;
int pkv_corpus_create(int device, int dim, int dtype, const char *index_db, const char *space, int64_t profile_id,
                      pkv_corpus **out);
int pkv_corpus_destroy(pkv_corpus *c);
const char *pkv_corpus_last_error(const pkv_corpus *c);
/* -> building at (artifact_rev, epoch).  A new revision (or scale) drops the rows held; the same one resumes.  int8:
 * the 4-byte scale artifact is mandatory and validated like artifact_scale. */
int pkv_corpus_begin(pkv_corpus *c, int64_t artifact_rev, const uint8_t *artifact, size_t artifact_len, uint64_t index_epoch);
/* One backfill chunk: n rows keyed by ascending item_data.id; rows at or below the cursor, or already added by the
 * inline hook, are skipped (replaying a chunk writes nothing twice).  Zero rows are written when the pair is no
 * longer building at this revision.  src_dtype: the corpus dtype, or PKV_F32 for an int8 corpus (GPU codec). */
int pkv_corpus_upload_chunk(pkv_corpus *c, int64_t artifact_rev, const int64_t *ids, const void *blobs, int src_dtype,
                            int64_t n, int64_t *written, int64_t *cursor);
/* The inline hook: one freshly written embedding while building or ready (searchable at once when ready).  A blob of
 * the wrong size downgrades the replica to pending and returns PKV_ERR_DIM_MISMATCH. */
int pkv_corpus_append_inline(pkv_corpus *c, int64_t data_id, const void *blob, size_t blob_bytes, int src_dtype,
                             uint64_t index_epoch);
int pkv_corpus_finish(pkv_corpus *c, int64_t artifact_rev, uint64_t index_epoch);
/* A write to the index DB that was not mirrored (bump_index_epoch): the replica is stale -> pending. */
int pkv_corpus_invalidate(pkv_corpus *c);
/* PKV_OK (+ the ReadyPair, + the searchable index, borrowed) only while ready at exactly this revision and epoch;
 * PKV_ERR_NOT_READY otherwise - `auto` then falls back to exact, `quant` fails (pql/preprocess.rs:356-362). */
int pkv_corpus_ready(pkv_corpus *c, int64_t artifact_rev, uint64_t index_epoch, pkv_ready_pair *pair, pkv_index **index);
int pkv_corpus_get_info(pkv_corpus *c, pkv_corpus_info *info);

/* -- cross-modal spaces (db/vector_quants.rs:480-510; image_embeddings.rs:140-199) -------------------------------------
 * A CLIP image setter and its "t"-prefixed text sibling form ONE space with one int8 scale; a filter without
 * clip_xmodal sees the image setter's rows only, with it both setters' rows.  One index holds the whole space. */
/* xmodal_text_sibling_name (db/vector_quants.rs:51-53): "t" + model. */
int pkv_xmodal_text_sibling_name(const char *model, char *out, size_t cap);
/* The loop of resolve_ready_pair (db/vector_quants.rs:1817-1867) over the setters a query involves.  states[i]:
 * 0 no such setter (skipped), 1 ready (pairs[i] valid), 2 exists but not ready.  PKV_OK + *out, or PKV_ERR_NOT_READY
 * (a setter not ready, or siblings that do not share scale and dim: "a rebuild is pending"). */
int pkv_resolve_ready_pair(const pkv_ready_pair *pairs, const int32_t *states, int n, pkv_ready_pair *out);
/* Per stored row of the space's indexes (same row order in the exact and the quant index): 0 image setter, 1 text
 * sibling.  HOST array; NULL withdraws it (single-setter space). */
int pkv_space_set_modality(pkv_space *s, const uint8_t *row_modality, int64_t n_rows);
/* pkv_space_search with SemanticImageArgs.clip_xmodal (image_embeddings.rs:21-83): 0 = the image setter's rows only. */
int pkv_space_search_xmodal(pkv_space *s, const float *queries, int nq, int query_dim, int metric, int index_mode,
                            const char *variant_or_null, int64_t k_arg, int depth, int clip_xmodal, int64_t *out_ids,
                            float *out_dist, int32_t *out_counts, int64_t *used_profile_id);

#ifdef __cplusplus
}
#endif
#endif /* PKV_H */
