// pkv_scan_ts.cu — int8 tensor-core scan with the QUERIES RESIDENT IN TENSOR MEMORY.
//
// Same arithmetic and exactness as pkv_scan_tc.cu / pkv_scan_tc2.cu (integer dot products on
// tcgen05 kind::i8, exact key for the survivors, bit-exact ids and distances); what changes is where
// the operands live.  With both UMMA operands in shared memory a 256x256 tile moves 96 B/clk of
// operand reads + 32 B/clk of TMA writes through a 128 B/clk shared memory: the tensor pipe starves.
// The query tile never changes during a pass, so here it is the A operand and sits in TMEM
// (tcgen05.mma ... [d_tmem], [a_tmem], b_desc): shared memory only carries the streamed corpus rows
// (64 B/clk of operand reads + 32 B/clk of TMA writes), and the 96 KB the query tile used to occupy
// becomes pipeline depth (~208 KB of row stages per CTA).  A stage holds CPS K-chunks (8 KB boxes)
// behind ONE barrier: the issuing warp's loop iteration (barrier round trip + issue) costs ~400
// clocks, an N=128 UMMA runs 64, so a stage must carry >= 8 of them to keep the tensor pipe fed.
//
//   cluster of 2 CTAs (cta_group::2), M = 256 queries (128 per CTA = the 128 TMEM lanes),
//   N = 128 corpus rows per tile (64 staged by each CTA), K = 32 per instruction.
//   TMEM columns: [0, dim_pad/4) the CTA's 128 queries (4 int8 per 32-bit column, lane = query),
//                 [256, 384) and [384, 512) the double-buffered 128 x 128 s32 accumulator
//                 (lane = query, column = row of the tile).
//   The accumulator is transposed with respect to pkv_scan_tc2.cu: an epilogue thread owns ONE
//   query and sees 32 rows per tcgen05.ld, so its pre-filter bound lives in a register.
//
// One launch serves up to `groups` x 256 queries: pair p scans for query group p % groups, and the
// pairs of one tile sequence read the same row tiles at about the same time, so every tile comes
// from HBM once and from L2 for the other groups.
//
// Algorithmic bytes per row per launch: dim_pad (+4 for the row norm); ops: 2 * queries * dim_pad.
#include "pkv_tc.cuh"

namespace pkv {

namespace {

constexpr int QM_CTA = 128;     // queries per CTA (TMEM lanes)
constexpr int QM = 256;         // queries per pair (MMA M)
constexpr int CHUNK_BYTES = 128;
constexpr int MAX_STAGES = 26;
constexpr int EPI_WARPS = 16;   // lane quarter = warp & 3, 32 accumulator columns each
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int SV_WARPS = 2;     // service warps (live mode): publish the CTA's thresholds in shared memory and run the
                                // in-kernel threshold selections the epilogue warps ask for
constexpr int SV_WARP0 = 2 + EPI_WARPS;
constexpr int TS_THREADS = 64 + EPI_THREADS + SV_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int HOLD_CAP = 64;
constexpr int HOLD_FLUSH = 32;

struct TsShared {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[3];
    uint64_t tmem_empty[3];
    uint32_t tmem_base;
    uint32_t pad;
    uint32_t hold_cnt[EPI_WARPS];
    uint32_t epi_done;                 // epilogue warps that have finished their tiles
    uint32_t refresh_req[QM_CTA];      // live mode: a push of this query hit its refresh trigger
    alignas(16) float thr[QM_CTA];
    alignas(16) int q_mag[QM_CTA];
    uint32_t hold_row[EPI_WARPS][HOLD_CAP];
    int hold_dot[EPI_WARPS][HOLD_CAP];
    uint32_t hold_col[EPI_WARPS][HOLD_CAP];
};

// Exact filter, exact key, candidate push for one pre-filter survivor (col = query within the CTA).
// Live mode: a push that hits the query's refresh trigger leaves a request for the service warps.
template <int METRIC>
__device__ __noinline__ void consider_ts(const ScanArgs &a, int qbase, int col, int d, uint32_t row, TsShared *sh) {
    const int q = qbase + col;
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

template <int METRIC>
__device__ __noinline__ void hold_ts(const ScanArgs &a, int qbase, int col, int d, uint32_t row, TsShared *sh, int ew) {
    const uint32_t slot = atomicAdd(&sh->hold_cnt[ew], 1u);
    if (slot < HOLD_CAP) {
        sh->hold_row[ew][slot] = row;
        sh->hold_dot[ew][slot] = d;
        sh->hold_col[ew][slot] = (uint32_t)col;
    } else {
        consider_ts<METRIC>(a, qbase, col, d, row, sh);
    }
}

template <int METRIC>
__device__ __forceinline__ void flush_ts(const ScanArgs &a, int qbase, TsShared *sh, int ew, int lane, uint32_t min_cnt) {
    __syncwarp();
    const uint32_t cnt = sh->hold_cnt[ew];
    if (cnt < min_cnt) return;
    const uint32_t n = cnt < HOLD_CAP ? cnt : HOLD_CAP;
    for (uint32_t e = lane; e < n; e += 32)
        consider_ts<METRIC>(a, qbase, (int)sh->hold_col[ew][e], sh->hold_dot[ew][e], sh->hold_row[ew][e], sh);
    __syncwarp();
    if (lane == 0) sh->hold_cnt[ew] = 0;
    __syncwarp();
}

// TN = corpus rows per tile (MMA N), NBUF = accumulator buffers in TMEM behind the query columns.  The round trip
// "MMAs of tile t done -> commit -> epilogue warps wake -> tcgen05.ld -> (remote) arrive -> issuing warp wakes ->
// first MMA of the tile that reuses the buffer" costs ~1500-2000 clocks, about one 128-row tile of tensor work, so
// with two buffers the pipe idled ~25-40 % of the time; three buffers hide it (96-row tiles when D > 512 leaves
// only 320 TMEM columns).
template <int METRIC, int CPS, int TN, int NBUF>
__global__ void __launch_bounds__(TS_THREADS, 1)
scan_i8_ts_kernel(const __grid_constant__ CUtensorMap tmap_rows, const ScanArgs a, const int q0, const int groups,
                  const int kchunks, const int stages) {
    constexpr int TILE_N = TN;
    constexpr int ROWS_CTA = TN / 2;                   // rows staged by each CTA per tile
    constexpr int BOX_BYTES = ROWS_CTA * CHUNK_BYTES;  // one TMA box
    constexpr int EPI_USED = (TN / 32) * 4;            // epilogue warps with work: 32 accumulator columns each
    const uint32_t acc_col0 = (uint32_t)(a.dim_pad / 4 + 31) / 32 * 32;  // accumulators behind the query columns
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    constexpr int STAGE_BYTES = CPS * BOX_BYTES;
    uint8_t *s_b = smem;  // [stages][CPS chunks][64 rows][128 B]
    TsShared *sh = reinterpret_cast<TsShared *>(s_b + (size_t)stages * STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t grp = pair % (uint32_t)groups, seq = pair / (uint32_t)groups, nseq = npairs / (uint32_t)groups;
    const int qbase = q0 + (int)grp * QM + (int)rank * QM_CTA;  // first query of this CTA
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + TILE_N - 1) / TILE_N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&sh->full[s], 1);
            tc::mbar_init(&sh->empty[s], 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            tc::mbar_init(&sh->tmem_full[b], 1);
            tc::mbar_init(&sh->tmem_empty[b], 2 * EPI_USED);
        }
        for (int w = 0; w < EPI_WARPS; ++w) sh->hold_cnt[w] = 0;
        sh->epi_done = 0;
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_rows);
    }
    if (warp == 1) {
        tc::tmem_alloc_cta2(&sh->tmem_base, TMEM_COLS);
        tc::tmem_relinquish_cta2();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    // per-thread query figures (epilogue thread = one query)
    float tq = 0.f;
    if (threadIdx.x < QM_CTA) sh->refresh_req[threadIdx.x] = 0;
    if (warp >= 2 && warp < SV_WARP0) {
        const int ew = warp - 2, quarter = warp & 3;
        const int col = quarter * 32 + lane;
        const int q = qbase + col;
        const float thr = q < a.nq ? __ldg(a.topk.thr_f + q) : -__int_as_float(0x7f800000);
        const int bm = q < a.nq ? __ldg(a.q_mag_i + q) : 0;
        tq = prefilter_query_figure<METRIC>(thr, bm);
        if ((ew >> 2) == 0) {
            sh->thr[col] = thr;
            sh->q_mag[col] = bm;
        }
        // this CTA's 128 queries -> TMEM columns [0, dim_pad/4): lane = query, 4 codes per column.
        // The four warps of a lane quarter split the K chunks.
        const int qrow = q < a.nq ? q : (a.nq - 1);  // padding lanes replay a real query; their bound rejects everything
        const uint8_t *qp = (const uint8_t *)a.queries + (size_t)qrow * a.dim_pad;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // all global loads first (<= 8 chunks of 32 codes per warp for D <= 1024), then the TMEM stores: the
        // prologue pays one memory latency instead of one per chunk
        const int n8 = a.dim_pad / 32;
        uint4 qlo[8], qhi[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int c8 = (ew >> 2) + it * (EPI_WARPS / 4);
            if (c8 < n8) {
                qlo[it] = __ldg(reinterpret_cast<const uint4 *>(qp + c8 * 32));
                qhi[it] = __ldg(reinterpret_cast<const uint4 *>(qp + c8 * 32 + 16));
            }
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int c8 = (ew >> 2) + it * (EPI_WARPS / 4);
            if (c8 < n8) {
                const uint32_t v[8] = {qlo[it].x, qlo[it].y, qlo[it].z, qlo[it].w, qhi[it].x, qhi[it].y, qhi[it].z, qhi[it].w};
                tc::tmem_st_32x8(lane_addr + (uint32_t)c8 * 8, v);
            }
        }
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();  // both CTAs: barriers initialised, TMEM allocated, queries resident
    tc::fence_after_sync();

    if (warp == 0) {
        // ===================== TMA producer (both CTAs: own 64 rows of every tile) =====================
        // whole warp, warp-uniform addresses; one elected lane issues
        if (seq < nseq) {
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0;
            const uint32_t full0 = tc::mapa(tc::smem_u32(&sh->full[0]), 0);
            for (uint32_t tile = seq; tile < ntiles; tile += nseq) {
                const int row0 = (int)(a.row_begin + tile * TILE_N + rank * ROWS_CTA);
                for (int kc = 0; kc < kchunks; kc += CPS) {
                    const int n = kchunks - kc < CPS ? kchunks - kc : CPS;  // chunks in this stage
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    if (issuer) {
                        if (rank == 0) tc::mbar_expect_tx(&sh->full[s], (uint32_t)(2 * n * BOX_BYTES));  // both CTAs' bytes
#pragma unroll
                        for (int j = 0; j < CPS; ++j)
                            if (j < n)
                                tc::tma_load_2d_cta2(s_b + (size_t)s * STAGE_BYTES + j * BOX_BYTES, &tmap_rows, full0 + s * 8u,
                                                     (kc + j) * CHUNK_BYTES, row0);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        // The whole warp runs the loop so that every address and descriptor is warp-uniform (uniform
        // registers, no per-instruction vector->uniform moves); one elected lane issues.
        if (rank == 0 && seq < nseq) {
            constexpr uint32_t idesc = tc::make_idesc(/*S32*/ 2, /*INT8*/ 1, QM, TILE_N);
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0, buf = 0, bph = 0;
            for (uint32_t tile = seq; tile < ntiles; tile += nseq) {
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + acc_col0 + buf * TILE_N;
                for (int kc = 0; kc < kchunks; kc += CPS) {
                    const int n = kchunks - kc < CPS ? kchunks - kc : CPS;
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
                    const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(s_b) + s * STAGE_BYTES);
                    const uint32_t a_tmem = tmem_base + (uint32_t)kc * (CHUNK_BYTES / 4);
                    if (issuer) {
#pragma unroll
                        for (int j = 0; j < CPS; ++j) {
                            if (j < n) {
#pragma unroll
                                for (int k = 0; k < CHUNK_BYTES / 32; ++k)
                                    tc::mma_i8_ts_cta2(d_tmem, a_tmem + j * (CHUNK_BYTES / 4) + k * 8,
                                                   b_desc + (uint64_t)(j * (BOX_BYTES / 16) + k * 2), idesc, (kc | j | k) != 0);
                            }
                        }
                        tc::mma_commit_cta2(&sh->empty[s]);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
                if (issuer) tc::mma_commit_cta2(&sh->tmem_full[buf]);
                __syncwarp();
                if (++buf == NBUF) { buf = 0; bph ^= 1; }
            }
        }
    } else if (seq < nseq && warp - 2 < EPI_USED) {
        // ===================== epilogue (both CTAs: own 128 queries x the tile's rows) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col0 = (ew >> 2) * 32;           // accumulator columns = rows of the tile
        const int qcol = quarter * 32 + lane;      // this thread's query within the CTA
        uint32_t buf = 0, bph = 0;
        // lane j prefetches the norm of row col0 + j (the warp's 32 rows of the tile)
        uint32_t nrow = a.row_begin + seq * TILE_N + col0 + lane;
        int am = (seq < ntiles && nrow < a.row_end) ? __ldg(a.row_mag_i + nrow) : -1;
        const uint32_t empty0 = tc::mapa(tc::smem_u32(&sh->tmem_empty[0]), 0);
        // live mode: this thread's query threshold is re-read once per tile from shared memory, where the service warps
        // publish what the whole grid tightens in global memory (a global load here would stall the accumulator hand-off)
        const bool live = a.topk.live != 0 && (qbase + qcol) < a.nq;
        const int bm_q = live ? sh->q_mag[qcol] : 0;
        for (uint32_t tile = seq; tile < ntiles; tile += nseq) {
            const uint32_t row_first = a.row_begin + tile * TILE_N + col0;
            const bool row_ok = am >= 0;
            const int am_min = __reduce_min_sync(0xffffffffu, row_ok ? am : 2147483647);
            const int am_max = __reduce_max_sync(0xffffffffu, row_ok ? am : 0);
            nrow = a.row_begin + (tile + nseq) * TILE_N + col0 + lane;
            am = (tile + nseq < ntiles && nrow < a.row_end) ? __ldg(a.row_mag_i + nrow) : -1;
            {
                // the norms of the tile after next are pulled into L2 now (see pkv_scan_img8.cu: one slow DRAM access
                // among the 32 epilogue warps of an accumulator stalls the whole hand-off)
                const uint32_t prow = a.row_begin + (tile + 3 * nseq) * TILE_N + col0 + lane;
                if (tile + 3 * nseq < ntiles && prow < a.row_end && (lane & 7) == 0)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.row_mag_i + prow));
            }
            if (live) tq = prefilter_query_figure<METRIC>(*(volatile const float *)&sh->thr[qcol], bm_q);
            const int bound = prefilter_bound<METRIC>(tq, sqrtf((float)am_min), sqrtf((float)am_max), (float)am_min);
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col0 + buf * TILE_N + col0, v);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(empty0 + buf * 8u);  // accumulator is in registers
            if (++buf == NBUF) { buf = 0; bph ^= 1; }
            // sign bit of (bound - 1 - d) is set iff d >= bound: OR them all, branch once
            int any = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) any |= bound - (int)v[j] - 1;
            if (any < 0) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int d = (int)v[j];
                    if (d >= bound) hold_ts<METRIC>(a, qbase, qcol, d, row_first + j, sh, ew);
                }
            }
            flush_ts<METRIC>(a, qbase, sh, ew, lane, HOLD_FLUSH);
        }
        flush_ts<METRIC>(a, qbase, sh, ew, lane, 1);
        __syncwarp();
        if (lane == 0) atomicAdd(&sh->epi_done, 1u);
    } else if (warp >= SV_WARP0 && a.topk.live) {
        // ===================== service warps (live mode) =====================
        const int sw = warp - SV_WARP0;
        const uint32_t expected = seq < nseq ? (uint32_t)EPI_USED : 0u;
        for (;;) {
            // one lane reads the flag: the whole warp must take the same exit
            const bool last = __shfl_sync(0xffffffffu, (uint32_t)(*(volatile uint32_t *)&sh->epi_done >= expected), 0) != 0;
            int trig_q = -1;
            for (int c = sw * 32 + lane; c < QM_CTA; c += SV_WARPS * 32) {
                const int q = qbase + c;
                if (q < a.nq) {
                    *(volatile float *)&sh->thr[c] = ld_live_f32(a.topk.thr_f + q);
                    if (*(volatile uint32_t *)&sh->refresh_req[c] && atomicExch(&sh->refresh_req[c], 0u)) trig_q = q;
                }
                live_refresh_pending<16>(a.topk, trig_q, lane);
                trig_q = -1;
            }
            if (last) break;
            __nanosleep(256);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();  // the peer may still be reading this CTA's row stages / signalling its barriers
    if (warp == 1) tc::tmem_dealloc_cta2(tmem_base, TMEM_COLS);
}

template <int METRIC, int CPS, int TN, int NBUF>
int launch_ts(const Index &ix, const ScanArgs &a, int q0, int groups, int kchunks, cudaStream_t s) {
    constexpr int ROWS_CTA = TN / 2;
    constexpr int STAGE_BYTES = CPS * ROWS_CTA * CHUNK_BYTES;
    CUtensorMap mrows;
    PKV_TRY(make_tmap_bytes(&mrows, ix.d_data, (uint64_t)ix.dim_pad, (uint64_t)ix.sealed_rows, (uint64_t)ix.pitch, ROWS_CTA));
    const size_t ctrl = sizeof(TsShared);
    int stages = (int)((227 * 1024 - 1024 - ctrl) / STAGE_BYTES);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (ix.opt.ts_stages > 1 && ix.opt.ts_stages < stages) stages = ix.opt.ts_stages;
    const size_t smem = 1024 + (size_t)stages * STAGE_BYTES + ctrl;
    auto kernel = scan_i8_ts_kernel<METRIC, CPS, TN, NBUF>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + TN - 1) / TN;
    uint32_t pairs = (uint32_t)ix.sm_count / 2;
    pairs = pairs / groups * groups;
    const uint32_t want = ntiles * (uint32_t)groups;  // one tile sequence per tile at most
    if (pairs > want) pairs = want;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TS_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PKV_CUDA(cudaLaunchKernelEx(&cfg, kernel, mrows, a, q0, groups, kchunks, stages));
    return PKV_OK;
}

// Tile shape by what TMEM has left behind the query columns (dim_pad/4, rounded up to 32):
// >= 384 columns: three 128-row accumulators; >= 288: three 96-row ones; else two 128-row ones.
template <int METRIC, int CPS>
int launch_ts_shape(const Index &ix, const ScanArgs &a, int q0, int groups, int kchunks, cudaStream_t s) {
    const int free_cols = TMEM_COLS - (ix.dim_pad / 4 + 31) / 32 * 32;
    const int nbuf_opt = ix.opt.ts_acc_buffers;
    if (nbuf_opt != 2 && free_cols >= 384) return launch_ts<METRIC, CPS, 128, 3>(ix, a, q0, groups, kchunks, s);
    if (nbuf_opt == 3 && free_cols >= 288) return  // measured 1-3 % slower than two 128-row buffers at D=768
        launch_ts<METRIC, CPS, 96, 3>(ix, a, q0, groups, kchunks, s);
    return launch_ts<METRIC, CPS, 128, 2>(ix, a, q0, groups, kchunks, s);
}

}  // namespace

int scan_ts_queries_per_launch(const Index &ix) {
    int g = ix.opt.ts_groups;
    if (g < 1) g = 1;
    if (g > 4) g = 4;
    return g * QM;
}

bool scan_ts_supported(const Index &ix) {
    return ix.dtype == PKV_I8 && ix.opt.tc_ts && ix.dim_pad <= 1024 && (ix.sm_count % 2) == 0;
}

// One launch over [row_begin,row_end) for queries [q0, min(nq, q0 + groups*256)).
int launch_scan_ts(const Index &ix, const ScanArgs &a, int q0, cudaStream_t s) {
    const int kchunks = ix.dim_pad / CHUNK_BYTES;
    int groups = (a.nq - q0 + QM - 1) / QM;
    const int gmax = scan_ts_queries_per_launch(ix) / QM;
    if (groups > gmax) groups = gmax;
    if (groups < 1) groups = 1;
    int cps = ix.opt.ts_chunks > 0 ? ix.opt.ts_chunks : (kchunks % 3 == 0 ? 3 : 2);
    if (cps > kchunks) cps = kchunks;
#define PKV_TS_CASE(M, C) \
    if (a.metric == M && cps == C) return launch_ts_shape<M, C>(ix, a, q0, groups, kchunks, s)
    PKV_TS_CASE(PKV_COSINE, 1); PKV_TS_CASE(PKV_COSINE, 2); PKV_TS_CASE(PKV_COSINE, 3); PKV_TS_CASE(PKV_COSINE, 4);
    PKV_TS_CASE(PKV_L2, 1); PKV_TS_CASE(PKV_L2, 2); PKV_TS_CASE(PKV_L2, 3); PKV_TS_CASE(PKV_L2, 4);
    PKV_TS_CASE(PKV_DOT, 1); PKV_TS_CASE(PKV_DOT, 2); PKV_TS_CASE(PKV_DOT, 3); PKV_TS_CASE(PKV_DOT, 4);
#undef PKV_TS_CASE
    return fail(PKV_ERR_INVALID, "unsupported metric %d / ts_chunks %d", a.metric, cps);
}

}  // namespace pkv
