// pkv_api.cu — the extern "C" boundary (include/pkv.h): index lifecycle, the chunked
// scan/select search driver, codec entry points.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pkv_internal.cuh"

namespace pkv {

// ------------------------------------------------------------------ errors
static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

// Makes `device` current for the duration of one API call and restores the caller's device afterwards
// (a server thread or a torch process may hold indexes on several GPUs).
int DeviceGuard::use(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(PKV_ERR_CUDA, "no usable CUDA device (%s); libpkv has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(PKV_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) {
        cudaGetLastError();
        cur = -1;
    }
    if (cur != device) {
        PKV_CUDA(cudaSetDevice(device));
        if (prev < 0) prev = cur;
    }
    return PKV_OK;
}
DeviceGuard::~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
}

// figures of the last search issued by the calling thread (pkv_counters.last_*)
struct LastSearch {
    double scan_ms = 0, total_ms = 0;
    int kind = 0;
};
static thread_local LastSearch g_last;

static int elem_size(int dtype) { return dtype == PKV_F32 ? 4 : (dtype == PKV_I8 ? 1 : 2); }
static int pad_dim(int dim, int dtype) {
    const int q = 128 / elem_size(dtype);  // components per 128-byte line
    return (dim + q - 1) / q * q;
}

Workspace::~Workspace() {
    cudaSetDevice(device);
    cudaFree(d_qraw);
    cudaFree(d_q);
    cudaFree(d_q_mag_f);
    cudaFree(d_q_mag_i);
    cudaFree(d_cand);
    cudaFree(d_cnt);
    cudaFree(d_thr_key);
    cudaFree(d_thr_f);
    cudaFree(d_status);
    cudaFree(d_pend_rows);
    cudaFree(d_pend_cnt);
    cudaFree(d_defer_rows);
    cudaFree(d_defer_dots);
    cudaFree(d_defer_cnt);
    cudaFree(d_defer_meta);
    cudaFree(d_q16);
    cudaFree(d_q8);
    cudaFree(d_q8_meta);
    cudaFree(d_q_scale);
    if (h_status) cudaFreeHost(h_status);
    cudaFree(d_out_ids);
    cudaFree(d_out_dist);
    cudaFree(d_out_counts);
    cudaFree(d_bitmap);
    for (auto &e : ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : chunk_ev)
        if (e) cudaEventDestroy(e);
    if (owns_stream && stream) cudaStreamDestroy(stream);
}

static constexpr int SUB_BATCH = 1024;  // queries per internal pass (bounds the workspace)

static int candidate_capacity(const Index &ix, int k) {
    int64_t cap = 4096;
    while (cap < 4 * (int64_t)k) cap <<= 1;
    if (ix.opt.candidate_capacity > 0) {
        cap = 2;
        while (cap < ix.opt.candidate_capacity) cap <<= 1;
        while (cap < 2 * (int64_t)k) cap <<= 1;
    }
    return (int)cap;
}

static int make_workspace(Index &ix, int nq, int k, size_t out_rows, Workspace **out) {
    Workspace *ws = nullptr;
    {
        std::lock_guard<std::mutex> g(ix.ws_mu);
        if (!ix.ws_free.empty()) {
            ws = ix.ws_free.back();
            ix.ws_free.pop_back();
        }
    }
    const int cap = candidate_capacity(ix, k);
    int nq_cap = 128;  // >= one query tile: TMA boxes over the query buffers never run out of bounds
    const int want = nq < SUB_BATCH ? nq : SUB_BATCH;
    while (nq_cap < want) nq_cap <<= 1;
    if (ws && (ws->nq_cap < nq_cap || ws->cap != cap || ws->dim_pad != ix.dim_pad || ws->out_cap < out_rows)) {
        nq_cap = ws->nq_cap > nq_cap ? ws->nq_cap : nq_cap;
        if (ws->out_cap > out_rows) out_rows = ws->out_cap;
        delete ws;
        ws = nullptr;
    }
    if (!ws) {
        ws = new (std::nothrow) Workspace();
        if (!ws) return fail(PKV_ERR_OOM, "out of host memory");
        ws->device = ix.device;
        ws->nq_cap = nq_cap;
        ws->cap = cap;
        ws->dim_pad = ix.dim_pad;
        ws->out_cap = out_rows;
        cudaError_t e = cudaSuccess;
        auto A = [&](void **p, size_t bytes) {
            if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 1);
        };
        if (cudaStreamCreateWithFlags(&ws->stream, cudaStreamNonBlocking) == cudaSuccess) ws->owns_stream = true;
        ws->qraw_bytes = (size_t)nq_cap * ix.dim * 4;
        A(&ws->d_qraw, ws->qraw_bytes);
        A(&ws->d_q, (size_t)nq_cap * ix.dim_pad * 4);
        A((void **)&ws->d_q_mag_f, sizeof(float) * nq_cap);
        A((void **)&ws->d_q_mag_i, sizeof(int32_t) * nq_cap);
        A((void **)&ws->d_cand, sizeof(uint64_t) * (size_t)nq_cap * cap);
        A((void **)&ws->d_cnt, sizeof(uint32_t) * nq_cap);
        A((void **)&ws->d_thr_key, sizeof(uint64_t) * nq_cap);
        A((void **)&ws->d_thr_f, sizeof(float) * nq_cap);
        A((void **)&ws->d_status, sizeof(SearchStatus));
        if (ix.dtype != PKV_I8) {
            ws->pend_cap = 4 * cap;
            A((void **)&ws->d_pend_rows, sizeof(uint32_t) * (size_t)nq_cap * ws->pend_cap);
            A((void **)&ws->d_pend_cnt, sizeof(uint32_t) * nq_cap);
            A((void **)&ws->d_defer_rows, sizeof(uint32_t) * (size_t)nq_cap * ws->pend_cap);
            A((void **)&ws->d_defer_dots, sizeof(int) * (size_t)nq_cap * ws->pend_cap);
            A((void **)&ws->d_defer_cnt, sizeof(uint32_t) * nq_cap);
            A((void **)&ws->d_defer_meta, sizeof(float4) * (size_t)nq_cap * ws->pend_cap);
            A((void **)&ws->d_q16, sizeof(__half) * (size_t)nq_cap * ix.dim_pad_h);
            A((void **)&ws->d_q_scale, sizeof(float) * nq_cap);
            A((void **)&ws->d_q8, (size_t)nq_cap * ix.dim_pad8);
            A((void **)&ws->d_q8_meta, sizeof(float4) * nq_cap);
        }
        A((void **)&ws->d_out_ids, sizeof(int64_t) * out_rows);
        A((void **)&ws->d_out_dist, sizeof(float) * out_rows);
        A((void **)&ws->d_out_counts, sizeof(int32_t) * (out_rows ? out_rows : 1));
        if (e == cudaSuccess) e = cudaMallocHost((void **)&ws->h_status, sizeof(SearchStatus));
        for (auto &ev : ws->ev)
            if (e == cudaSuccess) e = cudaEventCreate(&ev);
        if (e != cudaSuccess) {
            delete ws;
            cudaGetLastError();
            return fail(e == cudaErrorMemoryAllocation ? PKV_ERR_OOM : PKV_ERR_CUDA, "workspace allocation failed: %s",
                        cudaGetErrorString(e));
        }
    }
    *out = ws;
    return PKV_OK;
}

static void release_workspace(Index &ix, Workspace *ws) {
    std::lock_guard<std::mutex> g(ix.ws_mu);
    ix.ws_free.push_back(ws);
}

// ----------------------------------------------------------- search driver
struct SearchRun {
    Index &ix;
    Workspace &ws;
    cudaStream_t s;
    ScanArgs args;
    FilterSpec fs;
    int nq, k;
    int64_t safe_rows;
    uint32_t min_filled = 0;
    uint32_t last_max_raw = 0;  // largest per-query push count of the last synced chunk
    double scan_ms = 0;
    int launches = 0, scan_launches = 0;
    int depth_overflows = 0;
    bool use_tc = false, use_tc_f32 = false;
    bool live_capable = false;  // the scan kernels of this search can maintain the thresholds themselves
    int unsynced = 0;           // chunks enqueued since the last host sync
    // result buffers: the select behind the LAST chunk writes them itself (no finalize launch)
    int64_t *d_ids = nullptr;
    float *d_dist = nullptr;
    int32_t *d_counts = nullptr;
    bool outputs_written = false;
    bool defer = false;        // int8-image path: band pairs are parked until the final thresholds are known
    bool defer_dirty = false;  // a scan has parked pairs since the last deferred pass
};

// One chunk: scan kernel(s) over rows [b, e), then the select.
//   sync : wait for the chunk and read its status (splits the range when a candidate buffer overflowed)
//   live : the scan maintains the thresholds in-kernel (one launch for all remaining rows)
//   last : no chunk follows: the select writes the result rows
static int scan_range(SearchRun &r, int64_t b, int64_t e, bool sync, bool live, bool last) {
    r.args.row_begin = (uint32_t)b;
    r.args.row_end = (uint32_t)e;
    r.args.topk.live = live ? 1 : 0;
    const bool timed = r.ix.opt.time_kernels != 0;
    cudaEvent_t ev_begin = r.ws.ev[0], ev_end = r.ws.ev[1];
    if (!sync && timed) {
        // chunks enqueued back to back keep their own event pair; read after the final sync
        while (r.ws.chunk_ev.size() < (size_t)(2 * (r.unsynced + 1))) {
            cudaEvent_t ev = nullptr;
            PKV_CUDA(cudaEventCreate(&ev));
            r.ws.chunk_ev.push_back(ev);
        }
        ev_begin = r.ws.chunk_ev[2 * r.unsynced];
        ev_end = r.ws.chunk_ev[2 * r.unsynced + 1];
    }
    if (timed) PKV_CUDA(cudaEventRecord(ev_begin, r.s));
    int n = 0;
    // per-chunk status (max_raw_cnt / any_overflow / min_filled) is only read after a synced chunk
    if (sync) PKV_TRY(launch_reset_status(r.ws, r.s));
    // While some query has no threshold yet every pair is a candidate: that is dense work with one
    // push per pair, which the CUDA-core kernel does far more cheaply than the tensor-core
    // kernels' survivor path (built for rare survivors).
    const bool bootstrap = r.min_filled < (uint32_t)r.k && (e - b) <= r.safe_rows && r.ix.opt.simt_bootstrap;
    if (r.use_tc_f32 && !bootstrap)
        PKV_TRY(launch_scan_tc_f32(r.ix, r.args, r.ws, r.s, &n));
    else if (r.use_tc && !bootstrap)
        PKV_TRY(launch_scan_tc(r.ix, r.args, r.s, &n));
    else
        PKV_TRY(launch_scan_simt(r.ix, r.args, r.s, &n));
    if (r.defer && r.use_tc_f32 && !bootstrap) r.defer_dirty = true;
    bool deferred_now = false;
    if (last && r.defer_dirty) {
        // the pairs parked by every chunk of this search, against the thresholds the whole scan arrived at
        PKV_TRY(launch_rescore_deferred(r.ix, r.args, r.ws, r.s));
        n += 1;
        r.defer_dirty = false;
        deferred_now = true;
    }
    r.launches += n;
    r.scan_launches += n;
    if (timed) PKV_CUDA(cudaEventRecord(ev_end, r.s));
    // a chunk that may be followed by a live launch leaves KEY_MAX behind the kept keys (append-only buffers)
    const bool write_out = last && !sync;  // a synced chunk may still be split and re-scanned
    PKV_TRY(launch_select(r.ix, r.ws, r.nq, r.k, r.fs, /*clear_tail=*/r.live_capable && !last, deferred_now,
                          write_out ? r.d_ids : nullptr, r.d_dist, r.d_counts, r.s));
    if (write_out) r.outputs_written = true;
    r.launches += sync ? 2 : 1;
    if (!sync) {
        r.unsynced++;
        return PKV_OK;
    }
    PKV_CUDA(cudaMemcpyAsync(r.ws.h_status, r.ws.d_status, sizeof(SearchStatus), cudaMemcpyDeviceToHost, r.s));
    PKV_CUDA(cudaStreamSynchronize(r.s));
    if (timed) {
        float ms = 0.f;
        PKV_CUDA(cudaEventElapsedTime(&ms, r.ws.ev[0], r.ws.ev[1]));
        r.scan_ms += ms;
    }
    const SearchStatus st = *r.ws.h_status;
    r.min_filled = st.min_filled;
    r.last_max_raw = st.max_raw_cnt;
    static const bool trace = getenv("PKV_TRACE") != nullptr;
    if (trace && timed) {
        float ms = 0.f, ms2 = 0.f;
        cudaEventElapsedTime(&ms, r.ws.ev[0], r.ws.ev[1]);
        cudaEventRecord(r.ws.ev[0], r.s);
        cudaEventSynchronize(r.ws.ev[0]);
        cudaEventElapsedTime(&ms2, r.ws.ev[1], r.ws.ev[0]);
        fprintf(stderr, "[pkv] rows [%lld,%lld) nq %d%s: scan %.3f ms (%.0f GB/s), select+sync %.3f ms, max_raw_cnt %u, overflow %u\n",
                (long long)b, (long long)e, r.nq, live ? " live" : "", ms, (double)(e - b) * r.ix.pitch / (ms * 1e6), ms2,
                st.max_raw_cnt, st.any_overflow);
    }
    if (st.any_overflow) {
        // Some query pushed more candidates than its buffer holds.  The select kept the best k
        // of what fitted (all real rows, so the thresholds only got tighter); re-scan the range
        // in halves — duplicates are removed by the select — down to a size that cannot overflow.
        if (e - b <= r.safe_rows)
            return fail(PKV_ERR_CUDA, "internal: candidate overflow on a range of %lld rows (cap %d, k %d)",
                        (long long)(e - b), r.ws.cap, r.k);
        r.depth_overflows++;
        int64_t mid = b + (e - b) / 2;
        mid = (mid + 127) / 128 * 128;
        if (mid <= b || mid >= e) mid = b + (e - b) / 2;
        PKV_TRY(scan_range(r, b, mid, true, false, false));
        PKV_TRY(scan_range(r, mid, e, true, false, false));
    }
    return PKV_OK;
}

// The int8 kernels that re-read and re-select thresholds in-kernel: pkv_scan_tc.cu (<= 128 queries) and
// pkv_scan_ts.cu (more); the 2-CTA shared-memory kernel (tc_ts = 0) does not.
static bool live_int8_ok(const Index &ix, int nq) {
    if (nq <= 128) return true;
    return ix.opt.tc_ts && ix.dim_pad <= 1024 && (ix.sm_count % 2) == 0;
}

// Queries already on the device in caller layout; outputs to device buffers.
static int search_batch(Index &ix, Workspace &ws, cudaStream_t s, const void *d_qraw, int nq,
                        const pkv_search_params &p, const uint64_t *d_bitmap, int64_t *d_ids, float *d_dist,
                        int32_t *d_counts, bool allow_guess = true) {
    const int64_t N = ix.sealed_rows;
    const int k = p.k;
    PKV_TRY(launch_prep_queries(ix, ws, d_qraw, nq, p.query_dtype, s));
    ws.q8_ready = false;
    PKV_TRY(launch_reset_state(ws, nq, s));
    SearchRun r{ix, ws, s, ScanArgs{}, filter_spec_simt(ix.dtype, p.metric), nq, k, (int64_t)ws.cap - k};
    r.launches = 2;
    r.use_tc = scan_tc_supported(ix, nq);
    r.use_tc_f32 = scan_tc_f32_supported(ix, nq);
    if (r.use_tc_f32) r.fs = filter_spec_tc_f32(ix, p.metric, nq);
    r.d_ids = d_ids;
    r.d_dist = d_dist;
    r.d_counts = d_counts;
    // Live mode (one launch over every row behind the bootstrap chunk, thresholds maintained in-kernel): the
    // int8-image filter of f32 / f16 indexes and the int8 tensor-core kernels.  k <= 128: live_refresh keeps the
    // <= k + refresh_every keys that can matter in 16 registers per lane.
    const bool live_kernel = r.use_tc_f32 ? scan_tc_f32_kind(ix, nq) == 8 : (r.use_tc && live_int8_ok(ix, nq));
    r.live_capable = ix.opt.live && live_kernel && k <= 128 && N > r.safe_rows &&
                     !(r.use_tc_f32 && ix.opt.img8_fused == 0);
    // The live launch starts behind a short chunked prefix: a threshold learnt from few rows passes so many pairs
    // that the in-kernel feedback (re-score -> re-select -> re-read, ~10 us = tens of thousands of rows scanned
    // meanwhile) cannot keep up; behind ~64k rows the pass rate is low enough for the lag not to matter.
    const int64_t live_start = ix.opt.live_start_rows > 0 ? ix.opt.live_start_rows : 131072;
    // Measured on B200 (profiles/README.md, round 2): the live launch has a start-up transient and a drain tail that
    // only a long enough scan amortises; below ~0.7M rows the chunked schedule with deferred band pairs is faster.
    // (A guessed start has no prefix to amortise: measured faster than the chunked schedule from 100k rows up -
    // 100k / 250k / 500k rows: 406 k / 306 k / 249 k queries/s against 355 k / 259 k / 240 k - so it lifts this limit.)
    static const bool trace_env = getenv("PKV_TRACE") != nullptr;
    const bool guess_possible = allow_guess && ix.opt.guess && ix.opt.optimistic && !d_bitmap && !trace_env &&
                                ix.sample_rows > 0 && ix.sample_of_rows == N && N <= ix.opt.guess_max_rows &&
                                ix.sample_rows <= r.safe_rows && ix.opt.live_start_rows == 0;
    if (ix.opt.live == 1 && N - live_start < ix.opt.live_min_rows && !guess_possible) r.live_capable = false;
    ScanArgs &a = r.args;
    a.data = ix.d_data;
    a.pitch_bytes = ix.pitch;
    a.dim_pad = ix.dim_pad;
    a.dim = ix.dim;
    a.queries = ws.d_q;
    a.q_mag_f = ws.d_q_mag_f;
    a.q_mag_i = ws.d_q_mag_i;
    a.row_mag_i = ix.d_mag_i;
    a.row_mag_f = ix.d_mag_f;
    a.nq = nq;
    a.metric = p.metric;
    a.topk.cand = ws.d_cand;
    a.topk.cnt = ws.d_cnt;
    a.topk.thr_key = ws.d_thr_key;
    a.topk.thr_f = ws.d_thr_f;
    a.topk.cap = (uint32_t)ws.cap;
    a.topk.bitmap = d_bitmap;
    a.topk.bitmap_stride = p.bitmap_stride_words;
    a.topk.live = 0;
    a.topk.k = k;
    {
        int every = ix.opt.live_refresh > 0 ? ix.opt.live_refresh : 32;
        if (every > k / 2 + 1 && ix.opt.live_refresh <= 0) every = k / 2 + 1;  // a shallow search wants its few candidates to count at once
        if (every < 4) every = 4;
        a.topk.refresh_every = (uint32_t)every;
    }
    a.topk.fs = r.fs;
    a.topk.q_mag_f = ws.d_q_mag_f;
    a.topk.status = ws.d_status;
    r.defer = r.use_tc_f32 && scan_tc_f32_kind(ix, nq) == 8 && ix.opt.img8_defer != 0;
    if (!r.defer && r.use_tc_f32) r.live_capable = false;  // the live launch needs somewhere to park what it cannot take
    a.topk.defer = r.defer ? 1 : 0;

    // Chunk schedule: the first chunk (no threshold yet: every row is a candidate) is sized
    // so it cannot overflow; afterwards a chunk of c rows behind `seen` scanned rows yields
    // about k*c/seen candidates per query, so chunks grow geometrically.
    // (candidate work per chunk grows with the growth factor, chunk count shrinks with its log:
    // ~4x measured best on B200 for k=100)
    double growth = (double)(ws.cap - k) / (3.0 * k);
    if (growth > 4.0 && nq >= 32) growth = 4.0;  // few queries: candidates are cheap, chunks are not
    // int8-image filter: its error band passes ~(growth - 1) * 4k rows per chunk to the re-scorer (3 KB each), so a
    // wide batch does better with fresher thresholds (more, smaller chunks): measured 1.5x best at 256..1024 queries
    // (with the band pairs deferred to the end of the search the volume is ~3x smaller and 4x growth is best again)
    if (r.use_tc_f32 && scan_tc_f32_kind(ix, nq) == 8 && nq > 128 && growth > 1.5 && !r.defer) growth = 1.5;
    if (ix.opt.chunk_growth_x100 > 0) growth = ix.opt.chunk_growth_x100 / 100.0;
    if (growth < 0.25) growth = 0.25;
    // Optimistic mode: once every query has a threshold (after the first chunk) the remaining chunks
    // and their selects are enqueued back to back with no host round trip; the sticky overflow flag
    // is checked once at the end and, if it ever fired, the search is redone with a sync per chunk
    // (which is what splits overflowing ranges).  Bitmap searches stay careful until every query has a
    // threshold: their candidate rate is unknown until the first members are seen.
    const int kind_code =
        r.use_tc_f32 ? scan_tc_f32_kind(ix, nq) : r.use_tc ? 3 : (ix.dtype == PKV_F32 ? 1 : (ix.dtype == PKV_I8 ? 2 : 5));
    bool optimistic = ix.opt.optimistic && !d_bitmap && !trace_env;
    bool allow_live = r.live_capable;
    // Guessed start: instead of LEARNING the thresholds chunk by chunk, guess them - the r-th best of a strided sample
    // of the corpus (a real row's distance that about r * N / sample rows of the corpus beat) - start the candidate
    // lists empty, run ONE live launch over every row, and verify at the end: a query that found k rows under its guess
    // has provably missed nothing; the others (rare, see below) are searched again on the learning schedule.
    // r: with x = k * sample / N sample rows expected among the corpus' true top-k, the guess is too tight for a query
    // (fewer than k corpus rows beat the sample's r-th best) with probability P[Poisson(x) >= r]; r is the smallest rank
    // that keeps this below guess_miss_ppm (1e-4: one query in 10 000, i.e. one batch of 256 in 40 pays a second,
    // small search).  Measured on B200 (profiles/README.md): the tighter the guess the faster - 1M rows r = 4 / 6 / 10 /
    // 16: 215 k / 193 k / 167 k / 147 k queries/s; 10M rows r = 5 / 10: 78 k / 73 k (chunked prefix: 74 k), r = 20
    // overflows the candidate lists (the burst before the thresholds tighten) - hence the cap on the admitted rows.
    int guess_rank = 0;
    if (guess_possible && allow_live && optimistic) {
        int rank;
        if (ix.opt.guess_factor > 0) {   // explicit tightness (experiments): this many times k rows beat the guess
            rank = (int)((double)ix.opt.guess_factor * (double)k * (double)ix.sample_rows / (double)N + 0.5);
            const double admitted = (double)rank * (double)N / (double)ix.sample_rows;
            if (rank < 2 || rank > ix.sample_rows / 4 || admitted > 16000.0) rank = 0;
        } else {
            rank = pkv_guess_rank(k, ix.sample_rows, N, ix.opt.guess_miss_ppm);
        }
        guess_rank = rank;
    }
restart:
    int64_t pos = 0;
    r.min_filled = 0;
    r.last_max_raw = 0;
    r.unsynced = 0;
    r.outputs_written = false;
    r.defer_dirty = false;
    int64_t prev_chunk = 0;
    if (guess_rank > 0) {
        ScanArgs sa = r.args;
        sa.data = ix.d_sample;
        sa.row_mag_i = ix.d_sample_mag_i;
        sa.row_begin = 0;
        sa.row_end = (uint32_t)ix.sample_rows;
        sa.topk.live = 0;
        int n = 0;
        PKV_TRY(launch_scan_simt(ix, sa, s, &n));  // every (query, sample row) pair, exactly
        PKV_TRY(launch_select(ix, ws, nq, k, r.fs, /*clear_tail=*/true, false, nullptr, nullptr, nullptr, s, guess_rank));
        r.launches += n + 1;
        r.scan_launches += n;
        r.min_filled = (uint32_t)k;  // every query has a threshold now
        PKV_TRY(scan_range(r, 0, N, /*sync=*/false, /*live=*/true, /*last=*/true));
        pos = N;
    }
    while (pos < N) {
        const bool filled = r.min_filled >= (uint32_t)k;
        const bool go_live = allow_live && filled && pos >= live_start;
        int64_t chunk;
        if (go_live) {
            chunk = N - pos;
        } else if (!filled) {
            chunk = r.safe_rows;
            // before a live launch a smaller bootstrap chunk pays: its dense scoring and its 4096-key sort are the
            // largest fixed costs of a search
            if (pos == 0 && (allow_live || r.defer) && !d_bitmap && chunk > 2048 && k <= 512) chunk = 2048;
            if (pos == 0 && ix.opt.first_chunk_rows > 0 && ix.opt.first_chunk_rows < chunk)
                chunk = ix.opt.first_chunk_rows;
            // membership bitmap: a threshold-less chunk of c rows pushes (density * c) candidates per query; size the
            // next one from the rate the last one showed (halved: the density may vary) instead of crawling in
            // 4k-row steps through a sparse context - an overflow still splits the range
            if (d_bitmap && pos > 0 && prev_chunk > 0) {
                const double seen = r.last_max_raw > 0 ? (double)r.last_max_raw : 0.5;
                double c = 0.5 * (double)r.safe_rows * (double)prev_chunk / seen;
                if (c > 8.0 * (double)prev_chunk) c = 8.0 * (double)prev_chunk;
                if (c > (double)chunk) chunk = (int64_t)c;
            }
        } else {
            // before a live launch the prefix grows 4x per chunk whatever the batch (few chunks, then one launch)
            chunk = (int64_t)((double)pos * (allow_live ? 4.0 : growth));
            if (chunk < r.safe_rows) chunk = r.safe_rows;
            if (allow_live && pos + chunk > live_start) chunk = live_start - pos > r.safe_rows ? live_start - pos : chunk;
        }
        if (chunk >= 1024 && !go_live) chunk = chunk / 128 * 128;
        if (chunk > N - pos) chunk = N - pos;
        if (chunk < 1) chunk = 1;
        // Every row of the threshold-less first chunk becomes a candidate of every query, so with >= k rows each
        // query has its k-th best afterwards: no host round trip is needed to learn that.  (If that ever failed
        // the thresholds would stay +inf, the next chunk would overflow, and the sticky flag redoes the search.)
        const bool first_fill = optimistic && pos == 0 && !filled && chunk >= k;
        const bool sync = !(optimistic && (first_fill || (filled && pos > 0)));
        PKV_TRY(scan_range(r, pos, pos + chunk, sync, go_live, pos + chunk >= N));
        if (first_fill) r.min_filled = (uint32_t)k;
        prev_chunk = chunk;
        pos += chunk;
    }
    if (r.unsynced > 0) {
        PKV_CUDA(cudaMemcpyAsync(ws.h_status, ws.d_status, sizeof(SearchStatus), cudaMemcpyDeviceToHost, s));
        PKV_CUDA(cudaStreamSynchronize(s));
        if (ix.opt.time_kernels) {
            for (int i = 0; i < r.unsynced; ++i) {
                float ms = 0.f;
                PKV_CUDA(cudaEventElapsedTime(&ms, ws.chunk_ev[2 * i], ws.chunk_ev[2 * i + 1]));
                r.scan_ms += ms;
            }
        }
        if (guess_rank > 0 && !ws.h_status->sticky_overflow && !ws.h_status->defer_overflow &&
            ws.h_status->min_filled < (uint32_t)(N < k ? N : k)) {
            // some query found fewer than k rows under its guessed threshold: the guess was too tight for it.
            // The results of all other queries are final (they are written); the few that failed are searched again
            // with learnt thresholds as a small batch of their own and their rows overwritten.
            guess_rank = 0;
            std::vector<int32_t> cnt((size_t)nq);
            PKV_CUDA(cudaMemcpyAsync(cnt.data(), d_counts, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, s));
            PKV_CUDA(cudaStreamSynchronize(s));
            std::vector<int> failed;
            for (int q = 0; q < nq; ++q)
                if (cnt[(size_t)q] < (int32_t)(N < k ? N : k)) failed.push_back(q);
            if (!r.outputs_written || failed.empty() || failed.size() > 16) {
                r.depth_overflows++;  // (many misses: the sample misrepresents the corpus - learn the thresholds)
                PKV_TRY(launch_reset_state(ws, nq, s));
                goto restart;
            }
            const int nf = (int)failed.size();
            const size_t qbytes = (size_t)ix.dim * (size_t)elem_size(p.query_dtype);
            uint8_t *d_tmp = nullptr;
            const size_t off_ids = (qbytes * (size_t)nf + 255) / 256 * 256, off_dist = off_ids + sizeof(int64_t) * (size_t)nf * k,
                         off_cnt = off_dist + sizeof(float) * (size_t)nf * k;
            PKV_CUDA(cudaMalloc((void **)&d_tmp, off_cnt + sizeof(int32_t) * (size_t)nf));
            for (int i = 0; i < nf; ++i)
                cudaMemcpyAsync(d_tmp + qbytes * (size_t)i, (const uint8_t *)d_qraw + qbytes * (size_t)failed[(size_t)i], qbytes,
                                cudaMemcpyDeviceToDevice, s);
            // (the workspace is free again: every launch of this search has completed)
            const int saved_launches = r.launches, saved_scan = r.scan_launches;
            const double saved_ms = r.scan_ms;
            const SearchStatus st0 = *ws.h_status;
            int rc = search_batch(ix, ws, s, d_tmp, nf, p, nullptr, (int64_t *)(d_tmp + off_ids), (float *)(d_tmp + off_dist),
                                  (int32_t *)(d_tmp + off_cnt), /*allow_guess=*/false);
            if (rc == PKV_OK) {
                for (int i = 0; i < nf; ++i) {
                    const size_t q = (size_t)failed[(size_t)i];
                    cudaMemcpyAsync(d_ids + q * k, d_tmp + off_ids + sizeof(int64_t) * (size_t)i * k, sizeof(int64_t) * k,
                                    cudaMemcpyDeviceToDevice, s);
                    cudaMemcpyAsync(d_dist + q * k, d_tmp + off_dist + sizeof(float) * (size_t)i * k, sizeof(float) * k,
                                    cudaMemcpyDeviceToDevice, s);
                    cudaMemcpyAsync(d_counts + q, d_tmp + off_cnt + sizeof(int32_t) * (size_t)i, sizeof(int32_t),
                                    cudaMemcpyDeviceToDevice, s);
                }
                if (cudaStreamSynchronize(s) != cudaSuccess) rc = fail(PKV_ERR_CUDA, "guess repair: copy failed");
            }
            cudaFree(d_tmp);
            if (rc != PKV_OK) return rc;
            // the inner search has added its own figures to the index counters; this (outer) search adds the rest
            ix.n_fallback += nf;
            ix.n_launches += saved_launches + 3 * nf + 1;
            ix.n_scan_launches += saved_scan;
            ix.n_live_refreshes += st0.live_refreshes;
            ix.n_live_skips += st0.live_skips;
            ix.n_rescored += st0.rescored;
            ix.n_deferred += st0.deferred;
            g_last.scan_ms += saved_ms;
            g_last.kind = kind_code;
            return PKV_OK;
        }
        if (ws.h_status->sticky_overflow) {
            guess_rank = 0;
            // a candidate buffer overflowed somewhere along the unsynced chunks: redo the search on the careful
            // schedule (a sync per chunk, overflowing ranges split, thresholds fixed per launch, nothing deferred)
            r.depth_overflows++;
            optimistic = false;
            allow_live = false;
            r.defer = false;
            a.topk.defer = 0;
            PKV_TRY(launch_reset_state(ws, nq, s));
            goto restart;
        }
    }
    if (r.defer && ws.h_status->defer_overflow) {
        guess_rank = 0;
        // a query parked more pairs than its list holds (status as of the last sync, which followed the deferred pass)
        r.depth_overflows++;
        optimistic = false;
        allow_live = false;
        r.defer = false;
        a.topk.defer = 0;
        PKV_TRY(launch_reset_state(ws, nq, s));
        goto restart;
    }
    if (r.defer_dirty) {
        // a split re-scan behind the last chunk parked pairs again: one more deferred pass and select
        PKV_TRY(launch_rescore_deferred(ix, r.args, ws, s));
        PKV_TRY(launch_select(ix, ws, nq, k, r.fs, false, true, nullptr, nullptr, nullptr, s));
        r.defer_dirty = false;
        r.launches += 2;
        PKV_CUDA(cudaMemcpyAsync(ws.h_status, ws.d_status, sizeof(SearchStatus), cudaMemcpyDeviceToHost, s));
        PKV_CUDA(cudaStreamSynchronize(s));
        if (ws.h_status->sticky_overflow && !ws.h_status->any_overflow) {
            // (cannot happen on the careful schedule: its ranges are split until nothing overflows)
        }
    }
    if (!r.outputs_written) {
        PKV_TRY(launch_finalize(ix, ws, nq, k, d_ids, d_dist, d_counts, s));
        r.launches += 1;
    }
    // h_status holds the status as of the last host sync of this search (every path ends with one)
    ix.n_live_refreshes += ws.h_status->live_refreshes;
    ix.n_live_skips += ws.h_status->live_skips;
    ix.n_rescored += ws.h_status->rescored;
    ix.n_deferred += ws.h_status->deferred;
    ix.n_launches += r.launches;
    ix.n_scan_launches += r.scan_launches;
    ix.n_fallback += r.depth_overflows;
    g_last.scan_ms += r.scan_ms;
    g_last.kind = kind_code;
    return PKV_OK;
}

static int check_search_args(Index *ix, const void *queries, int nq, const pkv_search_params *p, const void *o_ids,
                             const void *o_dist, const void *o_counts) {
    if (!ix) return fail(PKV_ERR_INVALID, "index handle is NULL");
    if (!p) return fail(PKV_ERR_INVALID, "search params are NULL");
    if (nq < 0) return fail(PKV_ERR_INVALID, "nq must be >= 0");
    if (nq > 0 && (!queries || !o_ids || !o_dist || !o_counts))
        return fail(PKV_ERR_INVALID, "queries/outputs must not be NULL");
    if (p->k < 1) return fail(PKV_ERR_INVALID, "k must be a positive integer");
    if (p->k > PKV_MAX_K) return fail(PKV_ERR_INVALID, "k %d exceeds PKV_MAX_K %d", p->k, PKV_MAX_K);
    if (p->metric != PKV_L2 && p->metric != PKV_COSINE && p->metric != PKV_DOT)
        return fail(PKV_ERR_INVALID, "unknown metric %d", p->metric);
    const int qd = p->query_dtype;
    const bool ok = (ix->dtype == PKV_F32 && qd == PKV_F32) || (ix->dtype == PKV_I8 && (qd == PKV_I8 || qd == PKV_F32)) ||
                    (ix->dtype == PKV_F16 && (qd == PKV_F16 || qd == PKV_F32));
    if (!ok) return fail(PKV_ERR_INVALID, "query dtype %d not accepted by an index of dtype %d", qd, ix->dtype);
    if (ix->dtype == PKV_I8 && qd == PKV_F32 && !ix->has_scale)
        return fail(PKV_ERR_NOT_READY, "int8 index has no scale artifact; cannot quantise f32 queries");
    if (p->bitmap && p->bitmap_stride_words < 0) return fail(PKV_ERR_INVALID, "negative bitmap stride");
    if (ix->sealed_rows != ix->rows)
        return fail(PKV_ERR_NOT_READY, "index has %lld appended rows that are not sealed",
                    (long long)(ix->rows - ix->sealed_rows));
    return PKV_OK;
}

static int search_impl(Index &ix, const void *queries, bool host_io, int nq, const pkv_search_params &p,
                       int64_t *out_ids, float *out_dist, int32_t *out_counts, cudaStream_t user_stream) {
    PKV_USE_DEVICE(ix.device);
    std::shared_lock<std::shared_mutex> lock(ix.mu);
    PKV_TRY(check_search_args(&ix, queries, nq, &p, out_ids, out_dist, out_counts));
    if (nq == 0) return PKV_OK;
    const int k = p.k;
    const int qelem = elem_size(p.query_dtype);
    Workspace *ws = nullptr;
    const int sub = nq < SUB_BATCH ? nq : SUB_BATCH;
    PKV_TRY(make_workspace(ix, nq, k, host_io ? (size_t)sub * k : 0, &ws));
    // Host API: the workspace's own stream (the call is synchronous anyway).  Device API: the caller's stream, where
    // NULL means the legacy default stream - the search is then ordered behind whatever produced the queries and
    // the bitmap on stream 0 (torch's default stream), as pkv.h promises.
    cudaStream_t s = host_io ? ws->stream : user_stream;
    struct Guard {
        Index &ix;
        Workspace *ws;
        ~Guard() { release_workspace(ix, ws); }
    } guard{ix, ws};
    const int64_t words_per_bitmap = (ix.sealed_rows + 63) / 64;
    g_last.scan_ms = 0;
    if (ix.opt.time_kernels) PKV_CUDA(cudaEventRecord(ws->ev[2], s));
    for (int q0 = 0; q0 < nq; q0 += SUB_BATCH) {
        const int n = nq - q0 < SUB_BATCH ? nq - q0 : SUB_BATCH;
        const void *d_qraw;
        const uint64_t *d_bitmap = nullptr;
        int64_t *d_ids;
        float *d_dist;
        int32_t *d_counts;
        pkv_search_params pp = p;
        if (host_io) {
            PKV_CUDA(cudaMemcpyAsync(ws->d_qraw, (const uint8_t *)queries + (size_t)q0 * ix.dim * qelem,
                                     (size_t)n * ix.dim * qelem, cudaMemcpyHostToDevice, s));
            d_qraw = ws->d_qraw;
            if (p.bitmap) {
                const size_t words = p.bitmap_stride_words ? (size_t)p.bitmap_stride_words * n : (size_t)words_per_bitmap;
                if (p.bitmap_stride_words && p.bitmap_stride_words < words_per_bitmap)
                    return fail(PKV_ERR_INVALID, "bitmap stride %lld words is shorter than the %lld words the rows need",
                                (long long)p.bitmap_stride_words, (long long)words_per_bitmap);
                if (ws->bitmap_bytes < words * 8) {
                    cudaFree(ws->d_bitmap);
                    ws->d_bitmap = nullptr;
                    ws->bitmap_bytes = 0;
                    PKV_CUDA(cudaMalloc((void **)&ws->d_bitmap, words * 8));
                    ws->bitmap_bytes = words * 8;
                }
                PKV_CUDA(cudaMemcpyAsync(ws->d_bitmap, p.bitmap + (size_t)q0 * p.bitmap_stride_words, words * 8,
                                         cudaMemcpyHostToDevice, s));
                d_bitmap = ws->d_bitmap;
            }
            d_ids = ws->d_out_ids;
            d_dist = ws->d_out_dist;
            d_counts = ws->d_out_counts;
        } else {
            d_qraw = (const uint8_t *)queries + (size_t)q0 * ix.dim * qelem;
            d_bitmap = p.bitmap ? p.bitmap + (size_t)q0 * p.bitmap_stride_words : nullptr;
            d_ids = out_ids + (size_t)q0 * k;
            d_dist = out_dist + (size_t)q0 * k;
            d_counts = out_counts + q0;
        }
        PKV_TRY(search_batch(ix, *ws, s, d_qraw, n, pp, d_bitmap, d_ids, d_dist, d_counts));
        if (host_io) {
            PKV_CUDA(cudaMemcpyAsync(out_ids + (size_t)q0 * k, d_ids, sizeof(int64_t) * (size_t)n * k,
                                     cudaMemcpyDeviceToHost, s));
            PKV_CUDA(cudaMemcpyAsync(out_dist + (size_t)q0 * k, d_dist, sizeof(float) * (size_t)n * k,
                                     cudaMemcpyDeviceToHost, s));
            PKV_CUDA(cudaMemcpyAsync(out_counts + q0, d_counts, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
        }
    }
    if (ix.opt.time_kernels) PKV_CUDA(cudaEventRecord(ws->ev[3], s));
    PKV_CUDA(cudaStreamSynchronize(s));
    if (ix.opt.time_kernels) {
        float ms = 0.f;
        PKV_CUDA(cudaEventElapsedTime(&ms, ws->ev[2], ws->ev[3]));
        g_last.total_ms = ms;
    }
    ix.n_searches += 1;
    ix.n_queries += nq;
    return PKV_OK;
}

// ------------------------------------------------------------ combining
static void run_group(Index &ix, std::vector<PendingSearch *> &group) {
    if (group.size() == 1) {
        PendingSearch &g = *group[0];
        g.status = search_impl(ix, g.queries, true, g.nq, g.params, g.out_ids, g.out_dist, g.out_counts, nullptr);
        if (g.status != PKV_OK) g.error = g_last_error;
        return;
    }
    const pkv_search_params &p = group[0]->params;
    const size_t qbytes = (size_t)ix.dim * elem_size(p.query_dtype);
    int total = 0;
    for (auto *g : group) total += g->nq;
    std::vector<uint8_t> q((size_t)total * qbytes);
    std::vector<int64_t> ids((size_t)total * p.k);
    std::vector<float> dist((size_t)total * p.k);
    std::vector<int32_t> cnt((size_t)total);
    size_t off = 0;
    for (auto *g : group) {
        memcpy(q.data() + off * qbytes, g->queries, (size_t)g->nq * qbytes);
        off += g->nq;
    }
    const int st = search_impl(ix, q.data(), true, total, p, ids.data(), dist.data(), cnt.data(), nullptr);
    off = 0;
    for (auto *g : group) {
        g->status = st;
        if (st != PKV_OK) {
            g->error = g_last_error;
        } else {
            memcpy(g->out_ids, ids.data() + off * p.k, sizeof(int64_t) * (size_t)g->nq * p.k);
            memcpy(g->out_dist, dist.data() + off * p.k, sizeof(float) * (size_t)g->nq * p.k);
            memcpy(g->out_counts, cnt.data() + off, sizeof(int32_t) * (size_t)g->nq);
        }
        off += g->nq;
    }
    ix.n_combined += (int64_t)group.size();
}

static int combined_search(Index &ix, const void *queries, int nq, const pkv_search_params &p, int64_t *out_ids,
                           float *out_dist, int32_t *out_counts) {
    PendingSearch me;
    me.queries = queries;
    me.nq = nq;
    me.params = p;
    me.out_ids = out_ids;
    me.out_dist = out_dist;
    me.out_counts = out_counts;
    Combiner &c = ix.comb;
    std::unique_lock<std::mutex> lk(c.mu);
    c.queue.push_back(&me);
    while (!me.done) {
        if (c.busy) {
            c.cv.wait(lk);
            continue;
        }
        // become the executor: serve the oldest request and everything queued that can share its scan
        c.busy = true;
        std::vector<PendingSearch *> group;
        const pkv_search_params key = c.queue.front()->params;
        int total = 0;
        for (auto it = c.queue.begin(); it != c.queue.end();) {
            PendingSearch *g = *it;
            const bool same = g->params.metric == key.metric && g->params.k == key.k &&
                              g->params.query_dtype == key.query_dtype;
            if (same && total + g->nq <= SUB_BATCH) {
                group.push_back(g);
                total += g->nq;
                it = c.queue.erase(it);
            } else {
                ++it;
            }
        }
        lk.unlock();
        run_group(ix, group);
        lk.lock();
        for (auto *g : group) g->done = true;
        c.busy = false;
        c.cv.notify_all();
    }
    if (me.status != PKV_OK) g_last_error = me.error;
    return me.status;
}

// ------------------------------------------------------------ index storage
static int grow(Index &ix, int64_t need_rows, bool exact = false) {
    if (need_rows <= ix.cap_rows) return PKV_OK;
    int64_t cap = ix.cap_rows > 0 ? ix.cap_rows + ix.cap_rows / 2 : 1024;
    if (cap < need_rows || exact) cap = need_rows;
    uint8_t *nd = nullptr;
    int32_t *nmi = nullptr;
    float *nmf = nullptr;
    int64_t *nids = nullptr;
    __half *nsh = nullptr;
    int8_t *ni8 = nullptr;
    float4 *nim = nullptr;
    cudaError_t e = cudaMalloc((void **)&nd, (size_t)cap * ix.pitch);
    if (e == cudaSuccess && ix.dtype == PKV_F32 && (ix.opt.image_mask & 1))
        e = cudaMalloc((void **)&nsh, (size_t)cap * ix.dim_pad_h * sizeof(__half));
    if (e == cudaSuccess && ix.dtype != PKV_I8 && (ix.opt.image_mask & 2)) {
        e = cudaMalloc((void **)&ni8, (size_t)cap * ix.dim_pad8);
        if (e == cudaSuccess) e = cudaMalloc((void **)&nim, sizeof(float4) * cap);
    }
    if (e == cudaSuccess && ix.dtype == PKV_I8) e = cudaMalloc((void **)&nmi, sizeof(int32_t) * cap);
    if (e == cudaSuccess && ix.dtype != PKV_I8) e = cudaMalloc((void **)&nmf, sizeof(float) * cap);
    if (e == cudaSuccess && ix.d_ids) e = cudaMalloc((void **)&nids, sizeof(int64_t) * cap);
    if (e != cudaSuccess) {
        cudaFree(nd);
        cudaFree(nmi);
        cudaFree(nmf);
        cudaFree(nids);
        cudaFree(nsh);
        cudaFree(ni8);
        cudaFree(nim);
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? PKV_ERR_OOM : PKV_ERR_CUDA,
                    "cannot reserve %lld rows (%lld bytes): %s", (long long)cap, (long long)(cap * ix.pitch),
                    cudaGetErrorString(e));
    }
    if (ix.rows > 0) {
        PKV_CUDA(cudaMemcpy(nd, ix.d_data, (size_t)ix.rows * ix.pitch, cudaMemcpyDeviceToDevice));
        if (nmi) PKV_CUDA(cudaMemcpy(nmi, ix.d_mag_i, sizeof(int32_t) * ix.rows, cudaMemcpyDeviceToDevice));
        if (nmf) PKV_CUDA(cudaMemcpy(nmf, ix.d_mag_f, sizeof(float) * ix.rows, cudaMemcpyDeviceToDevice));
        if (nids) PKV_CUDA(cudaMemcpy(nids, ix.d_ids, sizeof(int64_t) * ix.rows, cudaMemcpyDeviceToDevice));
        if (nsh && ix.d_shadow)
            PKV_CUDA(cudaMemcpy(nsh, ix.d_shadow, (size_t)ix.sealed_rows * ix.dim_pad_h * sizeof(__half),
                                cudaMemcpyDeviceToDevice));
        if (ni8 && ix.d_img8) {
            PKV_CUDA(cudaMemcpy(ni8, ix.d_img8, (size_t)ix.sealed_rows * ix.dim_pad8, cudaMemcpyDeviceToDevice));
            PKV_CUDA(cudaMemcpy(nim, ix.d_img8_meta, sizeof(float4) * ix.sealed_rows, cudaMemcpyDeviceToDevice));
        }
    }
    // an image that did not exist before (option changed between appends) is rebuilt for every row at the next seal
    if ((nsh && !ix.d_shadow) || (ni8 && !ix.d_img8)) ix.image_rows = 0;
    cudaFree(ix.d_data);
    cudaFree(ix.d_mag_i);
    cudaFree(ix.d_mag_f);
    cudaFree(ix.d_ids);
    cudaFree(ix.d_shadow);
    cudaFree(ix.d_img8);
    cudaFree(ix.d_img8_meta);
    ix.d_shadow = nsh;
    ix.d_img8 = ni8;
    ix.d_img8_meta = nim;
    ix.d_data = nd;
    ix.d_mag_i = nmi;
    ix.d_mag_f = nmf;
    ix.d_ids = nids;
    ix.cap_rows = cap;
    return PKV_OK;
}

static int append_impl(Index &ix, const void *rows, const int64_t *row_ids, int64_t n, bool device_src) {
    PKV_USE_DEVICE(ix.device);
    if (n < 0) return fail(PKV_ERR_INVALID, "n must be >= 0");
    if (n == 0) return PKV_OK;
    if (!rows) return fail(PKV_ERR_INVALID, "rows must not be NULL");
    std::unique_lock<std::shared_mutex> lock(ix.mu);
    if (ix.rows + n >= 0xFFFFFFF0ll) return fail(PKV_ERR_UNSUPPORTED, "an index shard holds at most 2^32-16 rows");
    if (row_ids && !ix.d_ids) {
        // first explicit ids: materialise the implicit ones of earlier rows
        int64_t cap = ix.cap_rows > 0 ? ix.cap_rows : 1;
        PKV_CUDA(cudaMalloc((void **)&ix.d_ids, sizeof(int64_t) * cap));
        PKV_TRY(launch_fill_ids(ix.d_ids, 0, ix.rows, ix.row_base, nullptr));
    }
    PKV_TRY(grow(ix, ix.rows + n));
    const cudaMemcpyKind kind = device_src ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    uint8_t *dst = ix.d_data + (size_t)ix.rows * ix.pitch;
    const size_t row_bytes = (size_t)ix.dim * ix.elem;
    if ((int64_t)row_bytes == ix.pitch) {
        PKV_CUDA(cudaMemcpy(dst, rows, (size_t)n * row_bytes, kind));
    } else {
        PKV_CUDA(cudaMemset(dst, 0, (size_t)n * ix.pitch));
        PKV_CUDA(cudaMemcpy2D(dst, ix.pitch, rows, row_bytes, row_bytes, n, kind));
    }
    if (ix.d_ids) {
        if (row_ids)
            PKV_CUDA(cudaMemcpy(ix.d_ids + ix.rows, row_ids, sizeof(int64_t) * n, kind));
        else
            PKV_TRY(launch_fill_ids(ix.d_ids, ix.rows, ix.rows + n, ix.row_base, nullptr));
    }
    ix.rows += n;
    PKV_CUDA(cudaDeviceSynchronize());
    return PKV_OK;
}

}  // namespace pkv

using namespace pkv;

// =============================================================== C ABI
extern "C" {

int pkv_abi_version(void) { return PKV_ABI_VERSION; }

const char *pkv_last_error(void) { return g_last_error.c_str(); }

int pkv_device_count(int *count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (count) *count = 0;
        return fail(PKV_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
    }
    if (count) *count = c;
    if (c == 0) return fail(PKV_ERR_CUDA, "no usable CUDA device: device count is 0");
    return PKV_OK;
}

// ---- codec (host scalars are trivial pure functions; bulk work runs on the GPU)
float pkv_scale_from_absmax(float absmax) {
    if (absmax > 0.0f && absmax <= 3.402823466e+38f) return absmax / 127.0f;
    return 1.0f;
}
void pkv_scale_artifact(float scale, uint8_t out[4]) {
    uint32_t bits;
    memcpy(&bits, &scale, 4);
    for (int i = 0; i < 4; ++i) out[i] = (uint8_t)(bits >> (8 * i));
}
int pkv_artifact_scale(const uint8_t *artifact, size_t len, float *scale) {
    if (!artifact || len != 4) return fail(PKV_ERR_INVALID, "scale artifact must be exactly 4 bytes (got %zu)", len);
    uint32_t bits = (uint32_t)artifact[0] | ((uint32_t)artifact[1] << 8) | ((uint32_t)artifact[2] << 16) |
                    ((uint32_t)artifact[3] << 24);
    float s;
    memcpy(&s, &bits, 4);
    if (!(s > 0.0f) || !(s <= 3.402823466e+38f))
        return fail(PKV_ERR_INVALID, "scale artifact is not a positive finite f32");
    if (scale) *scale = s;
    return PKV_OK;
}

int pkv_blob_absmax_device(int device, const float *d_values, int64_t n, float *absmax, void *stream) {
    PKV_USE_DEVICE(device);
    if (n < 0 || (n > 0 && !d_values) || !absmax) return fail(PKV_ERR_INVALID, "bad arguments");
    float *d_out = nullptr;
    PKV_CUDA(cudaMalloc((void **)&d_out, sizeof(float)));
    int st = launch_absmax(d_values, n, d_out, (cudaStream_t)stream);
    if (st == PKV_OK) {
        cudaError_t e = cudaMemcpyAsync(absmax, d_out, sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) st = fail(PKV_ERR_CUDA, "absmax readback failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d_out);
    return st;
}

int pkv_blob_absmax(int device, const float *values, int64_t n, float *absmax) {
    PKV_USE_DEVICE(device);
    if (n < 0 || (n > 0 && !values) || !absmax) return fail(PKV_ERR_INVALID, "bad arguments");
    float *d = nullptr;
    PKV_CUDA(cudaMalloc((void **)&d, sizeof(float) * (n > 0 ? n : 1)));
    cudaError_t e = cudaMemcpy(d, values, sizeof(float) * n, cudaMemcpyHostToDevice);
    int st = e == cudaSuccess ? pkv_blob_absmax_device(device, d, n, absmax, nullptr)
                              : fail(PKV_ERR_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
    cudaFree(d);
    return st;
}

int pkv_quantize_int8_device(int device, const float *d_values, int64_t n, float scale, int8_t *d_codes,
                             void *stream) {
    PKV_USE_DEVICE(device);
    if (n < 0 || (n > 0 && (!d_values || !d_codes))) return fail(PKV_ERR_INVALID, "bad arguments");
    if (!(scale > 0.0f) || !(scale <= 3.402823466e+38f))
        return fail(PKV_ERR_INVALID, "scale must be a positive finite f32");
    PKV_TRY(launch_quantize(d_values, n, scale, d_codes, (cudaStream_t)stream));
    PKV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PKV_OK;
}

int pkv_quantize_int8(int device, const float *values, int64_t n, float scale, int8_t *codes) {
    PKV_USE_DEVICE(device);
    if (n < 0 || (n > 0 && (!values || !codes))) return fail(PKV_ERR_INVALID, "bad arguments");
    if (n == 0) return PKV_OK;
    float *d = nullptr;
    int8_t *c = nullptr;
    PKV_CUDA(cudaMalloc((void **)&d, sizeof(float) * n));
    cudaError_t e = cudaMalloc((void **)&c, n);
    if (e == cudaSuccess) e = cudaMemcpy(d, values, sizeof(float) * n, cudaMemcpyHostToDevice);
    int st = e == cudaSuccess ? pkv_quantize_int8_device(device, d, n, scale, c, nullptr)
                              : fail(PKV_ERR_CUDA, "quantize staging failed: %s", cudaGetErrorString(e));
    if (st == PKV_OK) {
        e = cudaMemcpy(codes, c, n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) st = fail(PKV_ERR_CUDA, "D2H copy failed: %s", cudaGetErrorString(e));
    }
    cudaFree(d);
    cudaFree(c);
    return st;
}

// ---- index lifecycle
int pkv_index_create(int device, int dim, int dtype, pkv_index **out) {
    if (!out) return fail(PKV_ERR_INVALID, "out must not be NULL");
    *out = nullptr;
    if (dim < 1 || dim > 4096) return fail(PKV_ERR_INVALID, "dim %d out of range (1..4096)", dim);
    if (dtype != PKV_F32 && dtype != PKV_I8 && dtype != PKV_F16) return fail(PKV_ERR_INVALID, "unknown dtype %d", dtype);
    PKV_USE_DEVICE(device);
    cudaDeviceProp prop;
    PKV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(PKV_ERR_UNSUPPORTED, "device %d is sm_%d%d; libpkv is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    Index *ix = new (std::nothrow) Index();
    if (!ix) return fail(PKV_ERR_OOM, "out of host memory");
    ix->device = device;
    ix->dim = dim;
    ix->dtype = dtype;
    ix->elem = elem_size(dtype);
    ix->dim_pad = pad_dim(dim, dtype);
    ix->pitch = (int64_t)ix->dim_pad * ix->elem;
    ix->dim_pad_h = (dim + 63) / 64 * 64;
    ix->dim_pad8 = (dim + 127) / 128 * 128;
    ix->sm_count = prop.multiProcessorCount;
    *out = reinterpret_cast<pkv_index *>(ix);
    return PKV_OK;
}

int pkv_index_destroy(pkv_index *h) {
    if (!h) return PKV_OK;
    Index *ix = reinterpret_cast<Index *>(h);
    DeviceGuard guard;
    if (guard.use(ix->device) != PKV_OK) return PKV_ERR_CUDA;
    for (Workspace *ws : ix->ws_free) delete ws;
    cudaFree(ix->d_data);
    cudaFree(ix->d_ids);
    cudaFree(ix->d_mag_i);
    cudaFree(ix->d_mag_f);
    cudaFree(ix->d_shadow);
    cudaFree(ix->d_img8);
    cudaFree(ix->d_img8_meta);
    cudaFree(ix->d_sample);
    cudaFree(ix->d_sample_mag_i);
    delete ix;
    return PKV_OK;
}

int pkv_index_reserve(pkv_index *h, int64_t rows) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    PKV_USE_DEVICE(ix.device);
    if (rows < 0) return fail(PKV_ERR_INVALID, "rows must be >= 0");
    std::unique_lock<std::shared_mutex> lock(ix.mu);
    if (rows <= ix.cap_rows) return PKV_OK;
    return grow(ix, rows, /*exact=*/true);  // no 1.5x headroom when the caller states the size
}

int pkv_index_append(pkv_index *h, const void *rows, const int64_t *row_ids, int64_t n) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    return append_impl(*reinterpret_cast<Index *>(h), rows, row_ids, n, false);
}
int pkv_index_append_device(pkv_index *h, const void *d_rows, const int64_t *d_row_ids, int64_t n) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    return append_impl(*reinterpret_cast<Index *>(h), d_rows, d_row_ids, n, true);
}

int pkv_index_set_scale(pkv_index *h, const uint8_t *artifact, size_t len) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    if (ix.dtype != PKV_I8) return fail(PKV_ERR_INVALID, "only int8 indexes carry a scale artifact");
    float s = 0.f;
    PKV_TRY(pkv_artifact_scale(artifact, len, &s));
    std::unique_lock<std::shared_mutex> lock(ix.mu);
    ix.scale = s;
    ix.has_scale = true;
    return PKV_OK;
}

int pkv_index_set_row_base(pkv_index *h, int64_t row_base) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    std::unique_lock<std::shared_mutex> lock(ix.mu);
    if (ix.d_ids && ix.rows > 0)
        return fail(PKV_ERR_INVALID, "row_base must be set before rows with explicit ids are appended");
    ix.row_base = row_base;
    return PKV_OK;
}

int pkv_index_seal(pkv_index *h) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    PKV_USE_DEVICE(ix.device);
    std::unique_lock<std::shared_mutex> lock(ix.mu);
    if (ix.sealed_rows < ix.rows) {
        PKV_TRY(launch_row_mags(ix, ix.sealed_rows, ix.rows, nullptr));
        if (ix.dtype == PKV_F32 && ix.d_shadow) {
            // fp16 image of the new rows, scaled by a power of two so that the largest component of
            // the whole index lands in [8192, 16384): exact scaling, no overflow, tiny values keep
            // full relative precision down to 2^-14/16384 of the largest one
            float *d_am = nullptr, am = 0.f;
            PKV_CUDA(cudaMalloc((void **)&d_am, sizeof(float)));
            int st = launch_absmax((const float *)(ix.d_data + (size_t)ix.sealed_rows * ix.pitch),
                                   (ix.rows - ix.sealed_rows) * (int64_t)ix.dim_pad, d_am, nullptr);
            cudaError_t ce = cudaMemcpy(&am, d_am, sizeof(float), cudaMemcpyDeviceToHost);
            cudaFree(d_am);
            PKV_TRY(st);
            PKV_CUDA(ce);
            int64_t from = ix.sealed_rows;
            const bool finite = am <= 3.0e38f;
            if (finite && am > ix.shadow_absmax) ix.shadow_absmax = am;
            if (ix.shadow_scale == 0.f || ix.shadow_absmax * ix.shadow_scale >= 32768.f) {
                float sc = 1.0f;
                if (ix.shadow_absmax > 0.f) sc = ldexpf(1.0f, 13 - ilogbf(ix.shadow_absmax));
                if (!(sc > 0.f) || !(sc <= 3.0e38f)) sc = 1.0f;
                ix.shadow_scale = sc;
                from = 0;  // (re)build the whole image with the new scale
            }
            PKV_TRY(build_shadow(ix, from, ix.rows, nullptr));
        }
        if (ix.d_img8) PKV_TRY(build_img8(ix, ix.image_rows < ix.sealed_rows ? ix.image_rows : ix.sealed_rows, ix.rows, nullptr));
        PKV_TRY(build_sample(ix, nullptr));
        PKV_CUDA(cudaDeviceSynchronize());
        ix.sealed_rows = ix.rows;
        ix.image_rows = ix.rows;
    }
    return PKV_OK;
}

int pkv_index_get_info(const pkv_index *h, pkv_index_info *info) {
    if (!h || !info) return fail(PKV_ERR_INVALID, "NULL argument");
    const Index &ix = *reinterpret_cast<const Index *>(h);
    memset(info, 0, sizeof(*info));
    info->abi_version = PKV_ABI_VERSION;
    info->device = ix.device;
    info->dim = ix.dim;
    info->dtype = ix.dtype;
    info->sealed = ix.sealed_rows == ix.rows ? 1 : 0;
    info->has_scale = ix.has_scale ? 1 : 0;
    info->scale = ix.scale;
    info->rows = ix.rows;
    info->capacity_rows = ix.cap_rows;
    info->device_bytes = ix.cap_rows * (ix.pitch + 4 + (ix.d_ids ? 8 : 0) + (ix.d_shadow ? ix.dim_pad_h * 2 : 0) +
                                        (ix.d_img8 ? ix.dim_pad8 + 16 : 0));
    info->row_base = ix.row_base;
    return PKV_OK;
}

int pkv_search(pkv_index *h, const void *queries, int nq, const pkv_search_params *params, int64_t *out_ids,
               float *out_dist, int32_t *out_counts) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    if (!params) return fail(PKV_ERR_INVALID, "search params are NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    if (ix.opt.combine && nq > 0 && nq <= 64 && !params->bitmap && queries && out_ids && out_dist && out_counts &&
        params->k >= 1 && params->k <= PKV_MAX_K)
        return combined_search(ix, queries, nq, *params, out_ids, out_dist, out_counts);
    return search_impl(ix, queries, true, nq, *params, out_ids, out_dist, out_counts, nullptr);
}

int pkv_search_device(pkv_index *h, const void *d_queries, int nq, const pkv_search_params *params, int64_t *d_out_ids,
                      float *d_out_dist, int32_t *d_out_counts, void *stream) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    if (!params) return fail(PKV_ERR_INVALID, "search params are NULL");
    return search_impl(*reinterpret_cast<Index *>(h), d_queries, false, nq, *params, d_out_ids, d_out_dist,
                       d_out_counts, (cudaStream_t)stream);
}

int pkv_distances_device(pkv_index *h, const void *d_queries, int nq, int metric, int query_dtype, float *d_out,
                         void *stream) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    PKV_USE_DEVICE(ix.device);
    std::shared_lock<std::shared_mutex> lock(ix.mu);
    pkv_search_params p;
    memset(&p, 0, sizeof(p));
    p.metric = metric;
    p.k = 1;
    p.query_dtype = query_dtype;
    int dummy = 0;
    PKV_TRY(check_search_args(&ix, d_queries, nq, &p, d_out, d_out, &dummy));
    if (nq == 0 || ix.sealed_rows == 0) return PKV_OK;
    if (nq > SUB_BATCH) return fail(PKV_ERR_INVALID, "pkv_distances_device takes at most %d queries per call", SUB_BATCH);
    Workspace *ws = nullptr;
    PKV_TRY(make_workspace(ix, nq, 1, 0, &ws));
    struct Guard {
        Index &ix;
        Workspace *ws;
        ~Guard() { release_workspace(ix, ws); }
    } guard{ix, ws};
    cudaStream_t s = (cudaStream_t)stream;  // NULL: the legacy default stream (ordered behind the caller's stream-0 work)
    PKV_TRY(launch_prep_queries(ix, *ws, d_queries, nq, query_dtype, s));
    ScanArgs a{};
    a.data = ix.d_data;
    a.pitch_bytes = ix.pitch;
    a.row_begin = 0;
    a.row_end = (uint32_t)ix.sealed_rows;
    a.dim_pad = ix.dim_pad;
    a.dim = ix.dim;
    a.queries = ws->d_q;
    a.q_mag_f = ws->d_q_mag_f;
    a.q_mag_i = ws->d_q_mag_i;
    a.row_mag_i = ix.d_mag_i;
    a.row_mag_f = ix.d_mag_f;
    a.nq = nq;
    a.metric = metric;
    a.dense_out = d_out;
    a.dense_stride = ix.sealed_rows;
    int n = 0;
    PKV_TRY(launch_scan_simt(ix, a, s, &n));
    ix.n_launches += n + 1;
    PKV_CUDA(cudaStreamSynchronize(s));
    return PKV_OK;
}

int pkv_rank_groups_device(pkv_index *h, const void *d_queries, int nq, const pkv_rank_params *p, int64_t *d_out_groups,
                           double *d_out_agg, int32_t *d_out_count, void *stream) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    if (!p) return fail(PKV_ERR_INVALID, "rank params are NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    PKV_USE_DEVICE(ix.device);
    if (nq < 1) return fail(PKV_ERR_INVALID, "nq must be >= 1");
    if (p->limit < 1 || p->offset < 0 || p->offset + p->limit > 2048)
        return fail(PKV_ERR_INVALID, "need limit >= 1, offset >= 0 and offset + limit <= 2048");
    if (p->aggregation < PKV_AGG_MIN || p->aggregation > PKV_AGG_AVG) return fail(PKV_ERR_INVALID, "unknown aggregation");
    if (p->n_groups < 0 || p->n_groups >= 0xFFFFFFFFll) return fail(PKV_ERR_INVALID, "n_groups out of range");
    if (!p->d_group_of_row || !d_out_groups || !d_out_agg || !d_out_count) return fail(PKV_ERR_INVALID, "NULL buffer");
    const int64_t rows = ix.sealed_rows;
    if ((double)rows * nq > 4.0e9)
        return fail(PKV_ERR_UNSUPPORTED, "%d query vectors x %lld rows exceeds the dense scoring buffer", nq, (long long)rows);
    cudaStream_t s = (cudaStream_t)stream;
    float *d_dist = nullptr;
    PKV_CUDA(cudaMallocAsync((void **)&d_dist, sizeof(float) * (size_t)(rows > 0 ? rows : 1) * nq, s));
    int st = rows > 0 ? pkv_distances_device(h, d_queries, nq, p->metric, p->query_dtype, d_dist, stream) : PKV_OK;
    if (st == PKV_OK)
        st = rank_groups(d_dist, rows, nq, p->d_group_of_row, p->d_weights, p->n_groups, p->aggregation, p->offset,
                         p->limit, d_out_groups, d_out_agg, d_out_count, s);
    cudaFreeAsync(d_dist, s);
    if (st != PKV_OK) return st;
    PKV_CUDA(cudaStreamSynchronize(s));
    return PKV_OK;
}

// similar_to (item_similarity.rs:432-581): the target item's own stored vectors are the queries; AGG runs over every
// (target vector, candidate vector) pair the xmodal rules admit.
int pkv_similar_to_device(pkv_index *h, const pkv_similar_params *p, int64_t *d_out_groups, double *d_out_agg,
                          int32_t *d_out_count, void *stream) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    if (!p) return fail(PKV_ERR_INVALID, "similar_to params are NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    PKV_USE_DEVICE(ix.device);
    if (p->n_targets < 1 || !p->d_target_rows) return fail(PKV_ERR_INVALID, "the target item has no stored vectors");
    if (p->limit < 1 || p->offset < 0 || p->offset + p->limit > 2048)
        return fail(PKV_ERR_INVALID, "need limit >= 1, offset >= 0 and offset + limit <= 2048");
    if (p->aggregation < PKV_AGG_MIN || p->aggregation > PKV_AGG_AVG) return fail(PKV_ERR_INVALID, "unknown aggregation");
    if (p->metric != PKV_L2 && p->metric != PKV_COSINE) return fail(PKV_ERR_INVALID, "distance function must be L2 or COSINE");
    if (p->n_groups < 0 || p->n_groups >= 0xFFFFFFFFll) return fail(PKV_ERR_INVALID, "n_groups out of range");
    if (!p->d_group_of_row || !d_out_groups || !d_out_agg || !d_out_count) return fail(PKV_ERR_INVALID, "NULL buffer");
    if (p->clip_xmodal && !p->d_modality)
        return fail(PKV_ERR_INVALID, "clip_xmodal needs the per-row modality of the space (image setter / text sibling)");
    const int64_t rows = ix.sealed_rows;
    const int T = p->n_targets;
    if ((double)rows * T > 4.0e9)
        return fail(PKV_ERR_UNSUPPORTED, "%d target vectors x %lld rows exceeds the dense scoring buffer", T, (long long)rows);
    cudaStream_t s = (cudaStream_t)stream;
    void *d_q = nullptr;
    float *d_dist = nullptr, *d_qw = nullptr;
    uint8_t *d_qm = nullptr;
    PKV_CUDA(cudaMallocAsync(&d_q, (size_t)T * ix.dim * ix.elem, s));
    PKV_CUDA(cudaMallocAsync((void **)&d_dist, sizeof(float) * (size_t)(rows > 0 ? rows : 1) * T, s));
    PKV_CUDA(cudaMallocAsync((void **)&d_qw, sizeof(float) * T, s));
    PKV_CUDA(cudaMallocAsync((void **)&d_qm, (size_t)T, s));
    int st;
    {
        std::shared_lock<std::shared_mutex> lock(ix.mu);
        st = launch_gather_rows(ix, p->d_target_rows, T, d_q, s);
    }
    if (st == PKV_OK) st = launch_gather_attrs(p->d_target_rows, T, rows, p->d_modality, p->d_weights, d_qm, d_qw, s);
    if (st == PKV_OK && rows > 0) st = pkv_distances_device(h, d_q, T, p->metric, ix.dtype, d_dist, stream);
    if (st == PKV_OK) {
        PairRules rules;
        rules.row_modality = p->d_modality;
        rules.q_modality = d_qm;
        rules.q_weights = p->d_weights ? d_qw : nullptr;
        rules.clip_xmodal = p->clip_xmodal ? 1 : 0;
        rules.skip_i2i = (p->clip_xmodal && !p->xmodal_i2i) ? 1 : 0;
        rules.skip_t2t = (p->clip_xmodal && !p->xmodal_t2t) ? 1 : 0;
        st = rank_groups(d_dist, rows, T, p->d_group_of_row, p->d_weights, p->n_groups, p->aggregation, p->offset, p->limit,
                         d_out_groups, d_out_agg, d_out_count, s, &rules);
    }
    cudaFreeAsync(d_q, s);
    cudaFreeAsync(d_dist, s);
    cudaFreeAsync(d_qw, s);
    cudaFreeAsync(d_qm, s);
    if (st != PKV_OK) return st;
    PKV_CUDA(cudaStreamSynchronize(s));
    return PKV_OK;
}

int pkv_index_get_rows_device(pkv_index *h, const int64_t *d_rows, int n, void *d_out, void *stream) {
    if (!h) return fail(PKV_ERR_INVALID, "index handle is NULL");
    Index &ix = *reinterpret_cast<Index *>(h);
    PKV_USE_DEVICE(ix.device);
    if (n < 0 || (n > 0 && (!d_rows || !d_out))) return fail(PKV_ERR_INVALID, "bad arguments");
    std::shared_lock<std::shared_mutex> lock(ix.mu);
    PKV_TRY(launch_gather_rows(ix, d_rows, n, d_out, (cudaStream_t)stream));
    PKV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PKV_OK;
}

int pkv_merge_topk_device(int device, const int64_t *d_ids, const float *d_dist, int parts, int nq, int k,
                          int64_t *d_out_ids, float *d_out_dist, int32_t *d_out_counts, void *stream) {
    PKV_USE_DEVICE(device);
    if (parts < 1 || parts > 256 || nq < 0 || k < 1) return fail(PKV_ERR_INVALID, "bad merge shape");
    if (nq > 0 && (!d_ids || !d_dist || !d_out_ids || !d_out_dist || !d_out_counts))
        return fail(PKV_ERR_INVALID, "NULL buffer");
    PKV_TRY(launch_merge(d_ids, d_dist, parts, nq, k, d_out_ids, d_out_dist, d_out_counts, (cudaStream_t)stream));
    PKV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PKV_OK;
}

int pkv_pack_topk_device(int device, const int64_t *d_ids, const float *d_dist, int64_t n, void *d_packed, void *stream) {
    PKV_USE_DEVICE(device);
    if (n < 0) return fail(PKV_ERR_INVALID, "n must be >= 0");
    if (n > 0 && (!d_ids || !d_dist || !d_packed)) return fail(PKV_ERR_INVALID, "NULL buffer");
    return launch_pack_topk(d_ids, d_dist, n, d_packed, (cudaStream_t)stream);
}

int pkv_merge_packed_device(int device, const void *d_packed, int parts, int nq, int k, int64_t *d_out_ids,
                            float *d_out_dist, int32_t *d_out_counts, void *stream) {
    PKV_USE_DEVICE(device);
    if (parts < 1 || parts > 256 || nq < 0 || k < 1) return fail(PKV_ERR_INVALID, "bad merge shape");
    if (nq > 0 && (!d_packed || !d_out_ids || !d_out_dist || !d_out_counts)) return fail(PKV_ERR_INVALID, "NULL buffer");
    return launch_merge_packed(d_packed, parts, nq, k, d_out_ids, d_out_dist, d_out_counts, (cudaStream_t)stream);
}

int pkv_aggregate_device(int device, const float *d_dist, const int64_t *d_item_of_row, const float *d_weights,
                         int64_t n, int64_t n_items, int agg, double *d_out, void *stream) {
    PKV_USE_DEVICE(device);
    if (n < 0 || n_items < 0 || agg < 0 || agg > 2) return fail(PKV_ERR_INVALID, "bad aggregate arguments");
    if ((n > 0 && (!d_dist || !d_item_of_row)) || (n_items > 0 && !d_out)) return fail(PKV_ERR_INVALID, "NULL buffer");
    PKV_TRY(launch_aggregate(d_dist, d_item_of_row, d_weights, n, n_items, agg, d_out, (cudaStream_t)stream));
    PKV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PKV_OK;
}

int pkv_index_counters(pkv_index *h, pkv_counters *out) {
    if (!h || !out) return fail(PKV_ERR_INVALID, "NULL argument");
    Index &ix = *reinterpret_cast<Index *>(h);
    memset(out, 0, sizeof(*out));
    out->searches = ix.n_searches;
    out->queries = ix.n_queries;
    out->kernel_launches = ix.n_launches;
    out->scan_launches = ix.n_scan_launches;
    out->fallback_queries = ix.n_fallback;
    out->combined_searches = ix.n_combined;
    out->live_refreshes = ix.n_live_refreshes;
    out->live_refresh_skips = ix.n_live_skips;
    out->rescored_pairs = ix.n_rescored;
    out->deferred_pairs = ix.n_deferred;
    out->last_scan_ms = g_last.scan_ms;
    out->last_total_ms = g_last.total_ms;
    out->last_scan_kind = g_last.kind;
    return PKV_OK;
}

int pkv_guess_rank(int k, int64_t sample_rows, int64_t rows, int miss_ppm) {
    if (k < 1 || sample_rows < 1 || rows < 1) return 0;
    const double x = (double)k * (double)sample_rows / (double)rows;
    const double eps = (double)miss_ppm * 1e-6;
    double term = exp(-x), cdf = 0.0;  // term = P[Poisson(x) = rank], cdf = P[Poisson(x) < rank]
    int rank;
    for (rank = 0; rank < 4096; ++rank) {
        if (rank >= 2 && 1.0 - cdf <= eps) break;
        cdf += term;
        term *= x / (double)(rank + 1);
    }
    // the lists must survive the burst of rows a guess admits before the in-kernel feedback tightens it
    const double admitted = (double)rank * (double)rows / (double)sample_rows;
    if (rank > sample_rows / 4 || admitted > 16000.0) return 0;
    return rank;
}

int pkv_index_set_option(pkv_index *h, const char *name, int64_t value) {
    if (!h || !name) return fail(PKV_ERR_INVALID, "NULL argument");
    Index &ix = *reinterpret_cast<Index *>(h);
    std::unique_lock<std::shared_mutex> lock(ix.mu);
    if (!strcmp(name, "force_simt")) ix.opt.force_simt = (int)value;
    else if (!strcmp(name, "candidate_capacity")) ix.opt.candidate_capacity = value;
    else if (!strcmp(name, "first_chunk_rows")) ix.opt.first_chunk_rows = value;
    else if (!strcmp(name, "chunk_growth_x100")) ix.opt.chunk_growth_x100 = value;
    else if (!strcmp(name, "time_kernels")) ix.opt.time_kernels = (int)value;
    else if (!strcmp(name, "tc_min_queries")) ix.opt.tc_min_queries = (int)value;
    else if (!strcmp(name, "tc_prefetch_tiles")) ix.opt.tc_prefetch_tiles = (int)value;
    else if (!strcmp(name, "tc_cta2")) ix.opt.tc_cta2 = (int)value;
    else if (!strcmp(name, "tc_ts")) ix.opt.tc_ts = (int)value;
    else if (!strcmp(name, "ts_groups")) ix.opt.ts_groups = (int)value;
    else if (!strcmp(name, "ts_stages")) ix.opt.ts_stages = (int)value;
    else if (!strcmp(name, "ts_acc_buffers")) ix.opt.ts_acc_buffers = (int)value;
    else if (!strcmp(name, "ts_chunks")) ix.opt.ts_chunks = (int)(value < 0 ? 0 : (value > 4 ? 4 : value));
    else if (!strcmp(name, "use_shadow")) ix.opt.use_shadow = (int)value;
    else if (!strcmp(name, "image_mask")) ix.opt.image_mask = (int)(value & 3);
    else if (!strcmp(name, "img8_max_queries")) ix.opt.img8_max_queries = (int)value;
    else if (!strcmp(name, "img8_peak_sigma_x10")) ix.opt.img8_peak_sigma_x10 = (int)value;
    else if (!strcmp(name, "simt_bootstrap")) ix.opt.simt_bootstrap = (int)value;
    else if (!strcmp(name, "optimistic")) ix.opt.optimistic = (int)value;
    else if (!strcmp(name, "combine")) ix.opt.combine = (int)value;
    else if (!strcmp(name, "tc_min_queries_f32")) ix.opt.tc_min_queries_f32 = (int)value;
    else if (!strcmp(name, "tc_min_queries_img")) ix.opt.tc_min_queries_img = (int)value;
    else if (!strcmp(name, "live")) ix.opt.live = (int)value;
    else if (!strcmp(name, "live_refresh")) ix.opt.live_refresh = (int)value;
    else if (!strcmp(name, "img8_fused")) ix.opt.img8_fused = (int)value;
    else if (!strcmp(name, "img8_defer")) ix.opt.img8_defer = (int)value;
    else if (!strcmp(name, "img8_epi")) ix.opt.img8_epi = (int)value;
    else if (!strcmp(name, "live_start_rows")) ix.opt.live_start_rows = value;
    else if (!strcmp(name, "live_min_rows")) ix.opt.live_min_rows = value;
    else if (!strcmp(name, "guess")) ix.opt.guess = (int)value;
    else if (!strcmp(name, "guess_max_rows")) ix.opt.guess_max_rows = value;
    else if (!strcmp(name, "guess_factor")) ix.opt.guess_factor = (int)value;
    else if (!strcmp(name, "guess_miss_ppm")) ix.opt.guess_miss_ppm = (int)value;
    else return fail(PKV_ERR_INVALID, "unknown option '%s'", name);
    return PKV_OK;
}

}  // extern "C"
