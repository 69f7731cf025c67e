// pkv_scan_tc.cu — tensor-core scan for int8 rows (tcgen05.mma kind::i8, s32 accumulators in TMEM).
//
// Replaces vec_distance_{cosine,L2}(vec_int8(quant), vec_int8(?)) evaluated row by row
// (pql/builder/filters/image_embeddings.rs:351-362, text_embeddings.rs:407-418) with a dense
// int8 contraction rows x queries; the integer dot products are exact, the final f32 key is
// computed with the reference's own operation sequence (pkv_device.cuh: i8_key) for the few
// pairs that survive the threshold, so top-k ids and distances are bit-exact.
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: streams 128-row x 128-byte K-chunks of the corpus into a ring of
//               SWIZZLE_128B stages (cp.async.bulk.tensor + mbarrier complete_tx)
//   warp 1      owns TMEM; one elected lane issues tcgen05.mma (M=128 rows, N=128 queries,
//               K=32 per instruction) against the query tile that stays resident in shared memory,
//               and tcgen05.commit's stage release / accumulator-ready barriers
//   warps 2..17 epilogue: tcgen05.ld the 128x128 s32 accumulator (double-buffered in TMEM, so the
//               next tile's MMAs overlap), integer pre-filter against a per-(warp,query) bound,
//               exact filter, exact key, candidate push
//
// Algorithmic bytes per row per query-tile pass: dim_pad (+4 for the row norm);
// algorithmic ops: 2 * 128 * dim_pad per row.
#include "pkv_tc.cuh"

namespace pkv {

namespace {

constexpr int TILE_M = 128;         // corpus rows per MMA
constexpr int TILE_N = 128;         // queries resident per CTA
constexpr int CHUNK_BYTES = 128;    // K bytes per stage row (one swizzle atom)
constexpr int STAGE_BYTES = TILE_M * CHUNK_BYTES;  // 16 KiB
constexpr int QCHUNK_BYTES = TILE_N * CHUNK_BYTES;
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARPS = 16;       // four warps per TMEM lane quarter, 32 columns each
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int SV_WARPS = 2;         // service warps (live mode): thresholds into shared memory, in-kernel threshold selections
constexpr int SV_WARP0 = 2 + EPI_WARPS;
constexpr int TC_THREADS = 64 + EPI_THREADS + SV_WARPS * 32;
constexpr int TMEM_COLS = 2 * TILE_N;  // double-buffered accumulator
constexpr int COLS_PER_WARP = TILE_N / 4;
constexpr int HOLD_CAP = 64;        // staged pre-filter survivors per epilogue warp
constexpr int HOLD_FLUSH = 32;      // flush (lane-parallel) once this many are parked

struct TcShared {  // control block behind the data stages
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t q_full;
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    uint32_t hold_cnt[EPI_WARPS];
    uint32_t epi_done;                 // epilogue warps that have finished their tiles
    uint32_t refresh_req[TILE_N];      // live mode: a push of this query hit its refresh trigger
    alignas(16) float thr[TILE_N];  // per query: filter threshold thr_f (metric specific)
    float tq[TILE_N];        // per query: finite, clamped figure the integer bound is derived from
    int q_mag[TILE_N];       // per query: integer squared norm
    alignas(16) int bound[EPI_WARPS][COLS_PER_WARP];  // per epilogue warp: integer pre-filter bound, current tile
    uint32_t hold_row[EPI_WARPS][HOLD_CAP];  // pre-filter survivors waiting for the lane-parallel flush
    int hold_dot[EPI_WARPS][HOLD_CAP];
    uint32_t hold_col[EPI_WARPS][HOLD_CAP];
};

// Exact filter, exact key, candidate push for one pre-filter survivor.
// Live mode: a push that hits the query's refresh trigger leaves a request for the service warps.
template <int METRIC>
__device__ __noinline__ void consider(const ScanArgs &a, int q0, int col, int d, uint32_t row, TcShared *sh) {
    const int q = q0 + col;
    if (q >= a.nq || row >= a.row_end) return;
    const int am = __ldg(a.row_mag_i + row);
    const int bm = sh->q_mag[col];
    if (!exact_filter<METRIC>(d, am, bm, *(volatile const float *)&sh->thr[col])) return;
    if (!topk_member(a.topk, q, row)) return;
    const int8_t *rowp = (const int8_t *)a.data + (size_t)row * (size_t)a.pitch_bytes;
    const int8_t *qp = (const int8_t *)a.queries + (size_t)q * a.dim_pad;
    const float key = i8_key(METRIC, d, am, bm, a.dim, rowp, qp);
    if (a.topk.live) {
        if (topk_push_live(a.topk, q, row, key)) *(volatile uint32_t *)&sh->refresh_req[col] = 1u;
    } else {
        topk_push(a.topk, q, row, key);
    }
}

// A lane found a pre-filter survivor: park it in its warp's shared-memory list (cheap) so that the
// expensive part runs later with 32 survivors per warp in flight instead of one.
template <int METRIC>
__device__ __noinline__ void hold(const ScanArgs &a, int q0, int col, int d, uint32_t row, TcShared *sh, int ew) {
    const uint32_t slot = atomicAdd(&sh->hold_cnt[ew], 1u);
    if (slot < HOLD_CAP) {
        sh->hold_row[ew][slot] = row;
        sh->hold_dot[ew][slot] = d;
        sh->hold_col[ew][slot] = (uint32_t)col;
    } else {
        consider<METRIC>(a, q0, col, d, row, sh);  // list full (unthresholded first chunk): do it now
    }
}

// Warp-private flush: no CTA-wide barrier, only __syncwarp.
template <int METRIC>
__device__ __forceinline__ void flush_held(const ScanArgs &a, int q0, TcShared *sh, int ew, int lane, uint32_t min_cnt) {
    __syncwarp();
    const uint32_t cnt = sh->hold_cnt[ew];
    if (cnt < min_cnt) return;
    const uint32_t n = cnt < HOLD_CAP ? cnt : HOLD_CAP;
    for (uint32_t e = lane; e < n; e += 32)
        consider<METRIC>(a, q0, (int)sh->hold_col[ew][e], sh->hold_dot[ew][e], sh->hold_row[ew][e], sh);
    __syncwarp();
    if (lane == 0) sh->hold_cnt[ew] = 0;
    __syncwarp();
}

template <int METRIC>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_i8_tc_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                  const ScanArgs a, const int q0, const int kchunks, const int stages, const int prefetch_tiles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // SWIZZLE_128B needs 1024-B alignment
    uint8_t *s_q = smem;                                    // [kchunks][128 queries][128 B]
    uint8_t *s_a = smem + (size_t)kchunks * QCHUNK_BYTES;   // [stages][128 rows][128 B]
    TcShared *sh = reinterpret_cast<TcShared *>(s_a + (size_t)stages * STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + TILE_M - 1) / TILE_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&sh->full[s], 1);
            tc::mbar_init(&sh->empty[s], 1);
        }
        tc::mbar_init(&sh->q_full, 1);
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&sh->tmem_full[b], 1);
            tc::mbar_init(&sh->tmem_empty[b], EPI_WARPS);
        }
        for (int w = 0; w < EPI_WARPS; ++w) sh->hold_cnt[w] = 0;
        sh->epi_done = 0;
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_rows);
        tc::prefetch_tmap(&tmap_q);
    }
    if (warp == 1) {
        tc::tmem_alloc(&sh->tmem_base, TMEM_COLS);
        tc::tmem_relinquish();
    }
    if (warp >= 2) {
        const int col = threadIdx.x - 64;
        if (col < TILE_N) {
            sh->refresh_req[col] = 0;
            const int q = q0 + col;
            const float thr = q < a.nq ? __ldg(a.topk.thr_f + q) : -__int_as_float(0x7f800000);
            const int bm = q < a.nq ? __ldg(a.q_mag_i + q) : 0;
            sh->thr[col] = thr;
            sh->q_mag[col] = bm;
            sh->tq[col] = prefilter_query_figure<METRIC>(thr, bm);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            tc::mbar_expect_tx(&sh->q_full, (uint32_t)kchunks * QCHUNK_BYTES);
            for (int kc = 0; kc < kchunks; ++kc)
                tc::tma_load_2d(s_q + (size_t)kc * QCHUNK_BYTES, &tmap_q, &sh->q_full, kc * CHUNK_BYTES, q0);
            uint32_t s = 0, ph = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row0 = (int)(a.row_begin + tile * TILE_M);
                // pull a tile further ahead into L2: more HBM bytes in flight than the smem ring holds
                const uint32_t ptile = tile + (uint32_t)prefetch_tiles * gridDim.x;
                if (prefetch_tiles > 0 && ptile < ntiles)
                    for (int kc = 0; kc < kchunks; ++kc)
                        tc::tma_prefetch_2d(&tmap_rows, kc * CHUNK_BYTES, (int)(a.row_begin + ptile * TILE_M));
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    tc::mbar_expect_tx(&sh->full[s], STAGE_BYTES);
                    tc::tma_load_2d(s_a + (size_t)s * STAGE_BYTES, &tmap_rows, &sh->full[s], kc * CHUNK_BYTES, row0);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // the whole warp runs the loop (warp-uniform addresses and descriptors stay in uniform
        // registers); one elected lane issues
        {
            constexpr uint32_t idesc = tc::make_idesc(/*S32*/ 2, /*INT8*/ 1, TILE_M, TILE_N);
            const bool issuer = tc::elect_one();
            tc::mbar_wait(&sh->q_full, 0);
            tc::fence_after_sync();
            uint32_t s = 0, ph = 0, t = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1, bph = (t >> 1) & 1;
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * TILE_N;
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
                    const uint64_t a_desc = tc::smem_desc_sw128(tc::smem_u32(s_a) + s * STAGE_BYTES);
                    const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(s_q) + (uint32_t)kc * QCHUNK_BYTES);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < CHUNK_BYTES / 32; ++k)
                            tc::mma_i8(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                        tc::mma_commit(&sh->empty[s]);  // stage free once these MMAs have read it
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
                if (issuer) tc::mma_commit(&sh->tmem_full[buf]);  // accumulator complete
                __syncwarp();
            }
        }
    } else if (warp >= SV_WARP0) {
        // ===================== service warps (live mode) =====================
        if (a.topk.live) {
            const int sw = warp - SV_WARP0;
            for (;;) {
                // one lane reads the flag: the whole warp must take the same exit
            const bool last = __shfl_sync(0xffffffffu, (uint32_t)(*(volatile uint32_t *)&sh->epi_done >= (uint32_t)EPI_WARPS), 0) != 0;
                for (int c = sw * 32 + lane; c < TILE_N; c += SV_WARPS * 32) {
                    const int q = q0 + c;
                    int trig_q = -1;
                    if (q < a.nq) {
                        const float thr = ld_live_f32(a.topk.thr_f + q);
                        *(volatile float *)&sh->thr[c] = thr;
                        *(volatile float *)&sh->tq[c] = prefilter_query_figure<METRIC>(thr, sh->q_mag[c]);
                        if (*(volatile uint32_t *)&sh->refresh_req[c] && atomicExch(&sh->refresh_req[c], 0u)) trig_q = q;
                    }
                    live_refresh_pending<16>(a.topk, trig_q, lane);
                }
                if (last) break;
                __nanosleep(256);
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;           // 0..7
        const int quarter = warp & 3;      // TMEM lane quarter this warp may access
        const int col0 = (ew >> 2) * COLS_PER_WARP;  // which 32 columns
        uint32_t t = 0;
        uint32_t row = a.row_begin + blockIdx.x * TILE_M + quarter * 32 + lane;
        int am = (blockIdx.x < ntiles && row < a.row_end) ? __ldg(a.row_mag_i + row) : -1;
        // live mode: the service warps keep sh->thr / sh->tq fresh (a stale value is only looser)
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1, bph = (t >> 1) & 1;
            const uint32_t cur_row = row;
            const bool row_ok = am >= 0;
            // rows past the end must not loosen (min) the warp's bound
            const int am_min = __reduce_min_sync(0xffffffffu, row_ok ? am : 2147483647);
            const int am_max = __reduce_max_sync(0xffffffffu, row_ok ? am : 0);
            // prefetch the next tile's row norm behind this tile's work
            row = a.row_begin + (tile + gridDim.x) * TILE_M + quarter * 32 + lane;
            am = (tile + gridDim.x < ntiles && row < a.row_end) ? __ldg(a.row_mag_i + row) : -1;
            const float s_min = sqrtf((float)am_min), s_max = sqrtf((float)am_max), am_min_f = (float)am_min;
#pragma unroll
            for (int i = 0; i < COLS_PER_WARP / 32; ++i) {
                const int c = i * 32 + lane;
                sh->bound[ew][c] = prefilter_bound<METRIC>(sh->tq[col0 + c], s_min, s_max, am_min_f);
            }
            __syncwarp();
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * TILE_N + col0;
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                uint32_t v[32];
                tc::tmem_ld_32x32(taddr + c * 32, v);
                tc::tmem_ld_wait();
                // sign bit of (bound - 1 - d) is set iff d >= bound: OR them all, branch once
                int any = 0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int4 b = *reinterpret_cast<const int4 *>(&sh->bound[ew][c * 32 + j]);
                    any |= (b.x - (int)v[j] - 1) | (b.y - (int)v[j + 1] - 1) | (b.z - (int)v[j + 2] - 1) | (b.w - (int)v[j + 3] - 1);
                }
                if (any < 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int d = (int)v[j];
                        if (d >= sh->bound[ew][c * 32 + j]) hold<METRIC>(a, q0, col0 + c * 32 + j, d, cur_row, sh, ew);
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
            flush_held<METRIC>(a, q0, sh, ew, lane, HOLD_FLUSH);
        }
        flush_held<METRIC>(a, q0, sh, ew, lane, 1);
        __syncwarp();
        if (lane == 0) atomicAdd(&sh->epi_done, 1u);
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved at run time so that libpkv.so carries no link-time dependency on libcuda.so
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

}  // namespace

// 2-D map over a row-major [rows][inner_bytes] byte matrix, box = 128 B x box_rows, SWIZZLE_128B
int make_tmap_bytes(CUtensorMap *m, const void *base, uint64_t inner_bytes, uint64_t rows, uint64_t pitch_bytes,
                    uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(PKV_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable in this driver");
    cuuint64_t dims[2] = {inner_bytes, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {CHUNK_BYTES, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(PKV_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return PKV_OK;
}

namespace {

template <int METRIC>
int launch_metric(const Index &ix, const ScanArgs &a, const CUtensorMap &mrows, const CUtensorMap &mq, int q0,
                  int kchunks, int stages, size_t smem, cudaStream_t s) {
    auto kernel = scan_i8_tc_kernel<METRIC>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + TILE_M - 1) / TILE_M;
    const unsigned grid = ntiles < (uint32_t)ix.sm_count ? ntiles : (unsigned)ix.sm_count;
    kernel<<<grid, TC_THREADS, smem, s>>>(mrows, mq, a, q0, kchunks, stages, ix.opt.tc_prefetch_tiles);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

}  // namespace

int launch_scan_tc2_tile(const Index &ix, const ScanArgs &a, const CUtensorMap &mrows, const CUtensorMap &mq, int q0,
                         cudaStream_t s);
// pkv_scan_ts.cu
bool scan_ts_supported(const Index &ix);
int scan_ts_queries_per_launch(const Index &ix);
int launch_scan_ts(const Index &ix, const ScanArgs &a, int q0, cudaStream_t s);

bool scan_tc_supported(const Index &ix, int nq) {
    if (ix.dtype != PKV_I8 || ix.opt.force_simt) return false;
    if (ix.dim_pad > 1024) return false;  // the resident query tile must leave room for >= 4 stages
    return nq >= ix.opt.tc_min_queries;
}

int launch_scan_tc(const Index &ix, const ScanArgs &a, cudaStream_t s, int *launches) {
    if (a.row_end <= a.row_begin || a.nq <= 0) return PKV_OK;
    const int kchunks = ix.dim_pad / CHUNK_BYTES;
    const size_t ctrl = sizeof(TcShared);
    int stages = (int)((227 * 1024 - 1024 - ctrl - (size_t)kchunks * QCHUNK_BYTES) / STAGE_BYTES);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return fail(PKV_ERR_UNSUPPORTED, "dim %d leaves no room for the row stages", ix.dim);
    const size_t smem = 1024 + (size_t)kchunks * QCHUNK_BYTES + (size_t)stages * STAGE_BYTES + ctrl;
    CUtensorMap mrows, mq;
    PKV_TRY(make_tmap_bytes(&mrows, ix.d_data, (uint64_t)ix.dim_pad, (uint64_t)ix.sealed_rows, (uint64_t)ix.pitch, TILE_M));
    PKV_TRY(make_tmap_bytes(&mq, a.queries, (uint64_t)ix.dim_pad, (uint64_t)a.nq, (uint64_t)ix.dim_pad, TILE_N));
    for (int q0 = 0; q0 < a.nq;) {
        *launches += 1;
        const int remaining = a.nq - q0;
        if (remaining > TILE_N && scan_ts_supported(ix)) {
            // queries resident in TMEM, up to ts_groups x 256 of them per corpus pass
            PKV_TRY(launch_scan_ts(ix, a, q0, s));
            q0 += scan_ts_queries_per_launch(ix);
            continue;
        }
        if (ix.opt.tc_cta2 && remaining > TILE_N && (ix.sm_count % 2) == 0) {
            // 256 queries per corpus pass on a CTA pair (cta_group::2)
            PKV_TRY(launch_scan_tc2_tile(ix, a, mrows, mq, q0, s));
            q0 += 2 * TILE_N;
            continue;
        }
        switch (a.metric) {
            case PKV_COSINE: PKV_TRY(launch_metric<PKV_COSINE>(ix, a, mrows, mq, q0, kchunks, stages, smem, s)); break;
            case PKV_L2: PKV_TRY(launch_metric<PKV_L2>(ix, a, mrows, mq, q0, kchunks, stages, smem, s)); break;
            default: PKV_TRY(launch_metric<PKV_DOT>(ix, a, mrows, mq, q0, kchunks, stages, smem, s)); break;
        }
        q0 += TILE_N;
    }
    return PKV_OK;
}

}  // namespace pkv
