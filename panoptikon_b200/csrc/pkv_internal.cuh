// pkv_internal.cuh — shared declarations of the libpkv translation units.
//
// Data layout in HBM:
//   corpus    : row-major, one row per stored vector, row pitch = dim padded with zeros to a
//               multiple of 128 bytes (f32: 32 comps, f16: 64, i8: 128) so every row starts on
//               a 128-byte line and a K-chunk is one TMA/UMMA swizzle atom.  Zero padding
//               changes neither dot products nor norms nor L2.
//   row_ids   : optional int64 per row (item_data.id); absent => row_base + position.
//   row_norm  : int32 sum of squares per row for int8 (exact), f32 for the tensor-core f32 path.
//   candidates: per query a buffer of `cap` packed 64-bit keys
//               (ordered f32 distance << 32 | local row); see pack_key().
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/pkv.h"

namespace pkv {

// ------------------------------------------------------------------ errors
void set_error(const char *fmt, ...);
int fail(int code, const char *fmt, ...);

#define PKV_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            return ::pkv::fail(_e == cudaErrorMemoryAllocation ? PKV_ERR_OOM : PKV_ERR_CUDA,        \
                               "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                               __LINE__);                                                           \
        }                                                                                           \
    } while (0)

#define PKV_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != PKV_OK) return _s;  \
    } while (0)

struct DeviceGuard {  // pkv_api.cu
    int prev = -1;
    int use(int device);
    ~DeviceGuard();
};
#define PKV_USE_DEVICE(dev)           \
    ::pkv::DeviceGuard _device_guard; \
    PKV_TRY(_device_guard.use(dev))

// ------------------------------------------------------------- key packing
// Total order of the contract: ascending f32 distance, NaN last, ties by ascending row.
// -0.0 is canonicalised to +0.0 so that equal distances really tie.
__host__ __device__ inline uint32_t ordered_bits(float d) {
    if (d != d) return 0xFFFFFFFFu;
    d += 0.0f;
    uint32_t b;
#ifdef __CUDA_ARCH__
    b = __float_as_uint(d);
#else
    memcpy(&b, &d, 4);
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ inline float unordered_bits(uint32_t o) {
    if (o == 0xFFFFFFFFu) {
#ifdef __CUDA_ARCH__
        return __uint_as_float(0x7FC00000u);
#else
        uint32_t n = 0x7FC00000u;
        float f;
        memcpy(&f, &n, 4);
        return f;
#endif
    }
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, 4);
    return f;
#endif
}
__host__ __device__ inline uint64_t pack_key(float d, uint32_t row) {
    return ((uint64_t)ordered_bits(d) << 32) | row;
}
static constexpr uint64_t KEY_MAX = 0xFFFFFFFFFFFFFFFFull;

// --------------------------------------------------------- filter kinds
// How a scan kernel turns the exact k-th best distance d_k into the cheap
// threshold `thr_f` it compares its own (approximate) per-pair figure against.
// The kernel passes a pair on to the exact key computation iff !(f > thr_f).
enum FilterKind : int {
    FK_EXACT_DIST = 0,  // f is the final distance itself (f32 DOT, int8 DOT): thr_f = d_k
    FK_COS_RATIO = 1,   // f = -dot / sqrt(aMag); thr_f = -(1 - d_k - 2.4e-7) * sqrt(bMag) + slack
    FK_L2_SQUARED = 2,  // f = squared L2; thr_f = d_k^2 * (1 + rel) + abs
};

struct FilterSpec {
    int kind;
    float rel;  // relative slack on the threshold
    float abs;  // absolute slack, in units of the filter figure (scaled by sqrt(bMag) for FK_COS_RATIO)
};

struct SearchStatus {  // device -> pinned host after every chunk
    uint32_t max_raw_cnt;
    uint32_t any_overflow;
    uint32_t min_filled;
    uint32_t sticky_overflow;  // like any_overflow but only cleared at the start of a search
    // cumulative over one search (cleared with sticky_overflow): what the in-kernel machinery did
    uint32_t live_refreshes;   // in-kernel threshold selections that ran to completion
    uint32_t live_skips;       // ... that gave up (more live keys than the register file of a warp holds)
    uint32_t rescored;         // (row, query) pairs re-scored exactly inside the int8-image scan kernel
    uint32_t deferred;         // ... parked until the end of the search and re-scored after the final-threshold check
    uint32_t defer_overflow;   // a query parked more pairs than its list holds: the search is redone without deferral
    uint32_t reserved[3];
};

// ------------------------------------------------------------ device state
struct TopkDev {
    uint64_t *cand;          // [nq][cap]
    uint32_t *cnt;           // [nq] raw number of pushes (may exceed cap: overflow)
    uint64_t *thr_key;       // [nq] exact key of the current k-th best (KEY_MAX while < k found)
    float *thr_f;            // [nq] filter-domain threshold
    uint32_t cap;
    const uint64_t *bitmap;  // optional membership bits over local rows
    int64_t bitmap_stride;   // words per query, 0 = shared
    // ---- live mode: ONE launch scans every remaining row; the thresholds are re-read from
    // global memory tile by tile and tightened in-kernel (live_refresh, pkv_device.cuh) while the scan runs.
    // The candidate buffers are append-only during such a launch and hold KEY_MAX in every unwritten slot.
    int live;                // 0: thresholds are fixed for the duration of a launch (chunked schedule)
    int k;                   // retrieval depth (live_refresh selects the k-th best)
    uint32_t refresh_every;  // a push whose count reaches a multiple of this re-selects the query's threshold
    FilterSpec fs;           // exact k-th distance -> filter-domain threshold (filter_threshold)
    const float *q_mag_f;    // [nq] |q|^2 (FK_COS_RATIO)
    SearchStatus *status;    // device counters of the search
    int defer;               // int8-image path: pairs inside the filter's error band are parked until the end of the search
};

struct ScanArgs {
    const void *data;       // corpus base
    int64_t pitch_bytes;    // row pitch in bytes
    uint32_t row_begin, row_end;
    int dim_pad;            // padded components per row
    int dim;                // true components per row
    const void *queries;    // [nq][dim_pad], index dtype (f16 index: f32 queries)
    const float *q_mag_f;   // [nq] f32 path: sum q^2 (sequential f32, as the reference accumulates bMag)
    const int32_t *q_mag_i; // [nq] int8 path: exact integer sum q^2
    const int32_t *row_mag_i; // [rows] int8 path: exact integer sum a^2
    const float *row_mag_f;   // [rows] tensor-core f32 path
    int nq;
    int metric;
    TopkDev topk;
    float *dense_out;      // non-NULL: write the exact distance of EVERY (query,row) pair here, push nothing
    int64_t dense_stride;  // elements between consecutive queries in dense_out
};


// --------------------------------------------------------------- host side
struct Workspace {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    int nq_cap = 0, dim_pad = 0, cap = 0;
    int pend_cap = 0;                 // filter survivors per query per chunk (a filter passes several rows per true candidate)
    void *d_qraw = nullptr;       // raw queries as the caller laid them out
    size_t qraw_bytes = 0;
    void *d_q = nullptr;          // padded queries in scan dtype
    float *d_q_mag_f = nullptr;
    int32_t *d_q_mag_i = nullptr;
    uint64_t *d_cand = nullptr;
    uint32_t *d_cnt = nullptr;
    uint64_t *d_thr_key = nullptr;
    float *d_thr_f = nullptr;
    uint32_t *d_pend_rows = nullptr;  // floating-point filter paths: rows awaiting exact re-scoring, [nq_cap][pend_cap]
    uint32_t *d_pend_cnt = nullptr;
    // int8-image path: pairs inside the filter's error band, parked for the whole search with the filter's dot
    // product and re-checked against the final thresholds (rescore_deferred_kernel)
    uint32_t *d_defer_rows = nullptr;
    int *d_defer_dots = nullptr;
    uint32_t *d_defer_cnt = nullptr;
    float4 *d_defer_meta = nullptr;
    __half *d_q16 = nullptr;          // fp16-image path: scaled half queries [nq_cap][dim_pad_h]
    int8_t *d_q8 = nullptr;           // int8-image path: per-query quantised codes [nq_cap][dim_pad8]
    float4 *d_q8_meta = nullptr;      // {1/s_q, |e_q|/s_q, |q|/s_q, |q|^2} per query
    bool q8_ready = false;            // codes of the current batch are in d_q8
    float *d_q_scale = nullptr;       // accumulator -> dot factor per query
    SearchStatus *d_status = nullptr;
    SearchStatus *h_status = nullptr;  // pinned
    int64_t *d_out_ids = nullptr;
    float *d_out_dist = nullptr;
    int32_t *d_out_counts = nullptr;
    uint64_t *d_bitmap = nullptr;
    size_t bitmap_bytes = 0;
    size_t out_cap = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> chunk_ev;  // begin/end pairs for chunks enqueued without a host sync
    ~Workspace();
};

struct Options {
    int force_simt = 0;
    int64_t candidate_capacity = 0;  // 0 = auto
    int64_t first_chunk_rows = 0;    // 0 = auto
    int64_t chunk_growth_x100 = 0;   // 0 = auto
    int time_kernels = 1;
    int tc_min_queries_img = 1;      // f32 with fp16 image: tensor-core filter from the first query
    int tc_min_queries_f32 = 9;     // f32: the FFMA kernel is HBM-bound up to ~16 queries
    int combine = 1;                 // merge concurrent small host searches into one scan (see Combiner)
    int optimistic = 1;              // enqueue all chunks after the first without host syncs; verify at the end
    int simt_bootstrap = 1;          // threshold-less first chunk runs on the CUDA-core kernel
    int use_shadow = -1;             // which image the tensor-core filter scans: -1 best available, 0 none (tf32 on
                                     // the f32 rows), 1 fp16 image, 2 int8 image
    int img8_max_queries = 1 << 30;  // best-available choice: batches above this use the fp16 image when both exist
    int img8_peak_sigma_x10 = 30;    // int8 image: rows up to mean + this/10 sigma of peakiness quantise without saturation
    int image_mask = 2;              // images built at seal for f32/f16 indexes: bit 0 fp16 (f32 only), bit 1 int8
    int tc_prefetch_tiles = 0;       // L2 prefetch distance of the TMA producer, in tiles per CTA
    int tc_cta2 = 1;                 // use the 2-CTA (cta_group::2) kernels when the batch is large enough
    int tc_ts = 1;                   // int8, > 128 queries: queries resident in TMEM (pkv_scan_ts.cu)
    int ts_groups = 4;               // 256-query groups served by one launch of that kernel (row tiles shared through L2)
    int ts_stages = 0;               // 0 = as many row stages as fit
    int ts_acc_buffers = 0;          // 0 = three 128-row accumulator buffers when TMEM has room (D <= 512) else two,
                                     // 2 = always two, 3 = three whenever possible (96-row tiles up to D = 896)
    int ts_chunks = 0;               // K-chunks (8 KB boxes) per stage: 0 = auto (3 when they divide the row, else 2)
    int tc_min_queries = 1;          // int8: the TMA/tcgen05 kernel streams at ~95% of HBM peak even for one query
    int live = 1;                    // one launch over all rows behind a short chunked prefix, thresholds maintained in-kernel:
                                     // 0 never, 1 when the scan is long enough to amortise it (live_min_rows), 2 whenever possible
    int live_refresh = 0;            // pushes per query between two in-kernel threshold selections (0 = auto: 32)
    int img8_defer = 1;              // int8-image path: pairs inside the filter's error band wait for the final thresholds
    int img8_fused = 1;              // int8-image path, exact re-scoring by warps of the scan kernel itself: 0 never
                                     // (no live launches either), 1 in live launches, 2 in every launch
    int img8_epi = 1;                // int8-image epilogue: survivors of the warp-wide bound are culled with the exact per-pair
                                     // bound in registers before they are held (0: at flush time, from global row figures)
    int64_t live_start_rows = 0;     // rows scanned on the chunked schedule before the live launch (0 = auto)
    int guess = 1;                   // start the live launch from thresholds GUESSED on a strided sample (verified at the
                                     // end: a query that did not find k rows under its guess is searched again)
    int64_t guess_max_rows = 24000000; // ... for corpora up to this many rows (beyond, even the 2nd best of the sample
                                     // admits more rows than the candidate lists survive: the chunked prefix learns)
    int guess_miss_ppm = 100;        // acceptable probability (parts per million, per query) that a guess is too tight
    int guess_factor = 0;            // > 0: explicit tightness instead - this many times k rows of the corpus beat the guess
    int64_t live_min_rows = 600000;  // live = 1: only when the live launch would cover at least this many rows (2 = always)
};

// Concurrent small searches (the server issues one query per request thread, 16 pool threads:
// db/connection.rs:235) are combined: while one scan runs, later arrivals queue up and the next
// scan serves all of them in one pass over the corpus.  Nobody waits when the index is idle.
struct PendingSearch {
    const void *queries;
    int nq;
    pkv_search_params params;
    int64_t *out_ids;
    float *out_dist;
    int32_t *out_counts;
    int status = PKV_OK;
    std::string error;
    bool done = false;
};
struct Combiner {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<PendingSearch *> queue;
    bool busy = false;
};

struct Index {
    int device = 0;
    int dim = 0, dim_pad = 0;
    int dtype = PKV_F32;
    int elem = 4;
    int64_t pitch = 0;  // bytes
    int64_t rows = 0, cap_rows = 0, sealed_rows = 0;
    int64_t row_base = 0;
    uint8_t *d_data = nullptr;
    int64_t *d_ids = nullptr;   // null until some append carries ids
    int64_t ids_rows = 0;       // rows of d_ids that are initialised
    int32_t *d_mag_i = nullptr; // int8 row sums of squares
    float *d_mag_f = nullptr;   // f32/f16 row sums of squares (tensor-core paths)
    __half *d_shadow = nullptr; // f32 index: power-of-two-scaled fp16 image of the rows (filter operand)
    int dim_pad_h = 0;          // components per fp16 image row (multiple of 64)
    float shadow_scale = 0.f;   // 0 = image not built yet
    int8_t *d_img8 = nullptr;   // f32/f16 index: int8 image, every row quantised with its own scale (filter operand)
    float4 *d_img8_meta = nullptr;  // per row {|a|/s_a, |codes|, |a - s_a codes|/s_a, 1/s_a}
    int dim_pad8 = 0;           // bytes per int8 image row (multiple of 128)
    int64_t image_rows = 0;     // rows whose images are built
    float img8_U = 0.f;         // code length |a|/s_a of every int8-image row (0 = not fixed yet)
    // threshold-guess sample: SAMPLE_ROWS stored rows taken at a constant stride over the sealed corpus (a copy, in
    // the index dtype and pitch) + their int8 norms; searched densely to GUESS each query's starting threshold
    uint8_t *d_sample = nullptr;
    int32_t *d_sample_mag_i = nullptr;
    int64_t sample_rows = 0;    // rows in the block (0 = not built)
    int64_t sample_of_rows = 0; // sealed_rows the block was drawn from
    float shadow_absmax = 0.f;
    bool has_scale = false;
    float scale = 1.0f;
    int sm_count = 148;
    Options opt;
    std::shared_mutex mu;       // searches shared, append/seal exclusive
    std::mutex ws_mu;
    std::vector<Workspace *> ws_free;
    Combiner comb;
    std::atomic<int64_t> n_combined{0};  // searches that shared a scan with another caller
    // counters
    std::atomic<int64_t> n_searches{0}, n_queries{0}, n_launches{0}, n_scan_launches{0}, n_fallback{0};
    std::atomic<int64_t> n_live_refreshes{0}, n_live_skips{0}, n_rescored{0}, n_deferred{0};
};

// rows awaiting exact re-scoring (filter kernels of the floating-point paths)
struct PendDev {
    uint32_t *rows;  // [nq][cap]
    uint32_t *cnt;   // [nq]
    uint32_t cap;
    int *dots;       // [nq][cap] the filter's integer dot product (int8-image path: deferred pairs are re-checked
                     // against the final threshold before their rows are gathered); may be NULL elsewhere
    float4 *meta;    // [nq][cap] the row figures of the parked pair (the re-check then reads its list sequentially
                     // instead of gathering row_meta at random); may be NULL elsewhere
};

// ---- kernels / launchers (each returns a pkv_status) ----
// pkv_scan_simt.cu
int launch_scan_simt(const Index &ix, const ScanArgs &a, cudaStream_t s, int *launches);
// pkv_scan_tc.cu (tcgen05 tensor-core path)
bool scan_tc_supported(const Index &ix, int nq);
int launch_scan_tc(const Index &ix, const ScanArgs &a, cudaStream_t s, int *launches);
// pkv_scan_tc_f32.cu (tf32 filter + exact re-scoring)
bool scan_tc_f32_supported(const Index &ix, int nq);
FilterSpec filter_spec_tc_f32(const Index &ix, int metric, int nq);
int launch_scan_tc_f32(const Index &ix, const ScanArgs &a, Workspace &ws, cudaStream_t s, int *launches);
int scan_tc_f32_kind(const Index &ix, int nq);
int build_shadow(Index &ix, int64_t row_begin, int64_t row_end, cudaStream_t s);
int launch_rescore(const Index &ix, const ScanArgs &a, const PendDev &pend, SearchStatus *status, cudaStream_t s);
// pkv_scan_img8.cu (int8 image of floating-point rows: tcgen05 kind::i8 filter + exact re-scoring)
bool img8_usable(const Index &ix);
int build_img8(Index &ix, int64_t row_begin, int64_t row_end, cudaStream_t s);
int launch_scan_img8(const Index &ix, const ScanArgs &a, Workspace &ws, cudaStream_t s, int *launches);
int launch_rescore_deferred(const Index &ix, const ScanArgs &a, Workspace &ws, cudaStream_t s);

// pkv_operator.cu
struct PairRules {  // which (target row, candidate row) pairs of similar_to's self-join take part, and the target-side weight
    const uint8_t *row_modality;  // per stored row: 0 image ("clip"), 1 text ("text-embedding"); NULL = no rules
    const uint8_t *q_modality;    // per target row
    const float *q_weights;       // per target row, or NULL
    int clip_xmodal, skip_i2i, skip_t2t;
};
int rank_groups(const float *d_dist, int64_t rows, int nq, const int64_t *d_group_of_row, const float *d_weights,
                int64_t n_groups, int agg, int offset, int limit, int64_t *d_out_groups, double *d_out_agg,
                int32_t *d_out_count, cudaStream_t s, const PairRules *rules_or_null = nullptr);
int launch_gather_attrs(const int64_t *d_rows, int n, int64_t n_rows, const uint8_t *d_modality, const float *d_weights,
                        uint8_t *d_q_mod, float *d_q_w, cudaStream_t s);
int launch_gather_rows(const Index &ix, const int64_t *d_rows, int n, void *d_out, cudaStream_t s);
// pkv_topk.cu
int launch_reset_status(Workspace &ws, cudaStream_t s);
int launch_prep_queries(const Index &ix, Workspace &ws, const void *d_qraw, int nq, int query_dtype, cudaStream_t s);
int launch_reset_state(Workspace &ws, int nq, cudaStream_t s);
int launch_select(const Index &ix, Workspace &ws, int nq, int k, FilterSpec fs, bool clear_tail, bool clear_deferred,
                  int64_t *d_ids, float *d_dist, int32_t *d_counts, cudaStream_t s, int guess_rank = 0);
int build_sample(Index &ix, cudaStream_t s);
int launch_finalize(const Index &ix, Workspace &ws, int nq, int k, int64_t *d_ids, float *d_dist, int32_t *d_counts,
                    cudaStream_t s);
int launch_row_mags(Index &ix, int64_t row_begin, int64_t row_end, cudaStream_t s);
int launch_merge(const int64_t *d_ids, const float *d_dist, int parts, int nq, int k, int64_t *o_ids, float *o_dist,
                 int32_t *o_counts, cudaStream_t s);
int launch_pack_topk(const int64_t *d_ids, const float *d_dist, int64_t n, void *d_packed, cudaStream_t s);
int launch_merge_packed(const void *d_packed, int parts, int nq, int k, int64_t *o_ids, float *o_dist, int32_t *o_counts,
                        cudaStream_t s);
int launch_aggregate(const float *d_dist, const int64_t *d_item, const float *d_w, int64_t n, int64_t n_items, int agg,
                     double *d_out, cudaStream_t s);
int launch_absmax(const float *d_values, int64_t n, float *d_out, cudaStream_t s);
int launch_quantize(const float *d_values, int64_t n, float scale, int8_t *d_codes, cudaStream_t s);
int launch_fill_ids(int64_t *d_ids, int64_t begin, int64_t end, int64_t base, cudaStream_t s);

FilterSpec filter_spec_simt(int dtype, int metric);

}  // namespace pkv
