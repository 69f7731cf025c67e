// pkv_scan_tc2.cu — 2-CTA (cta_group::2) tensor-core scan for int8 rows.
//
// Same algorithm and exactness as pkv_scan_tc.cu; the difference is the tile shape.  Two CTAs on
// the two SMs of a TPC form a cluster and issue ONE tcgen05.mma.cta_group::2 of M=256 rows x
// N=256 queries: each CTA stages its own 128 corpus rows (A) and keeps 128 of the 256 queries
// resident (B is split across the pair, each tensor core reads the peer's half), so a pass over
// the corpus serves 256 queries with the same shared-memory footprint and the same HBM/L2 bytes
// per SM as the 128-query 1-CTA kernel — half the corpus passes for a large batch.
//
//   CTA rank 0 (leader): arms the stage barriers with the pair's byte count, issues the MMAs and
//                        the multicast commits; both CTAs run a TMA producer and 8 epilogue warps.
//   full[s]      lives in the leader, completed by BOTH CTAs' TMA loads (cta_group::2 form)
//   empty[s], tmem_full[b]  exist in both CTAs, arrived by tcgen05.commit ... multicast::cluster
//   tmem_empty[b] lives in the leader, 32 arrivals (16 epilogue warps x 2 CTAs, remote via mapa)
#include "pkv_tc.cuh"

namespace pkv {

namespace {

constexpr int ROWS_PER_CTA = 128;
constexpr int TILE_ROWS = 256;
constexpr int QN = 256;      // queries per pass of the pair (MMA N)
constexpr int QN_CTA = 128;  // resident in each CTA
constexpr int CHUNK_BYTES = 128;
constexpr int STAGE_BYTES = ROWS_PER_CTA * CHUNK_BYTES;
constexpr int QCHUNK_BYTES = QN_CTA * CHUNK_BYTES;
constexpr int MAX_STAGES = 8;
constexpr int EPI_WARPS = 16;  // four per TMEM lane quarter, 64 columns each: enough warps per scheduler to hide TMEM-load latency
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int TC_THREADS = 64 + EPI_THREADS;
constexpr int TMEM_COLS = 2 * QN;  // double-buffered 128 x 256 accumulator: the whole TMEM
constexpr int COLS_PER_WARP = QN / 4;
constexpr int HOLD_CAP = 64;     // staged pre-filter survivors per epilogue warp
constexpr int HOLD_FLUSH = 32;   // flush (lane-parallel) once this many are parked

struct Tc2Shared {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t q_full;
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    uint32_t hold_cnt[EPI_WARPS];
    alignas(16) float thr[QN];
    alignas(16) float tq[QN];
    alignas(16) int q_mag[QN];
    alignas(16) int bound[EPI_WARPS][COLS_PER_WARP];
    uint32_t hold_row[EPI_WARPS][HOLD_CAP];
    int hold_dot[EPI_WARPS][HOLD_CAP];
    uint32_t hold_col[EPI_WARPS][HOLD_CAP];
};

template <int METRIC>
__device__ __noinline__ void consider2(const ScanArgs &a, int q0, int col, int d, uint32_t row, const Tc2Shared *sh) {
    const int q = q0 + col;
    if (q >= a.nq || row >= a.row_end) return;
    const int am = __ldg(a.row_mag_i + row);
    const int bm = sh->q_mag[col];
    if (!exact_filter<METRIC>(d, am, bm, sh->thr[col])) return;
    if (!topk_member(a.topk, q, row)) return;
    const int8_t *rowp = (const int8_t *)a.data + (size_t)row * (size_t)a.pitch_bytes;
    const int8_t *qp = (const int8_t *)a.queries + (size_t)q * a.dim_pad;
    topk_push(a.topk, q, row, i8_key(METRIC, d, am, bm, a.dim, rowp, qp));
}

// A lane found a pre-filter survivor: park it in its warp's shared-memory list (cheap) so that the
// expensive part runs later with 32 survivors per warp in flight instead of one.
template <int METRIC>
__device__ __noinline__ void hold2(const ScanArgs &a, int q0, int col, int d, uint32_t row, Tc2Shared *sh, int ew) {
    const uint32_t slot = atomicAdd(&sh->hold_cnt[ew], 1u);
    if (slot < HOLD_CAP) {
        sh->hold_row[ew][slot] = row;
        sh->hold_dot[ew][slot] = d;
        sh->hold_col[ew][slot] = (uint32_t)col;
    } else {
        consider2<METRIC>(a, q0, col, d, row, sh);  // list full (unthresholded first chunk): do it now
    }
}

// Warp-private flush: no CTA-wide barrier, only __syncwarp.
template <int METRIC>
__device__ __forceinline__ void flush_held2(const ScanArgs &a, int q0, Tc2Shared *sh, int ew, int lane, uint32_t min_cnt) {
    __syncwarp();
    const uint32_t cnt = sh->hold_cnt[ew];
    if (cnt < min_cnt) return;
    const uint32_t n = cnt < HOLD_CAP ? cnt : HOLD_CAP;
    for (uint32_t e = lane; e < n; e += 32)
        consider2<METRIC>(a, q0, (int)sh->hold_col[ew][e], sh->hold_dot[ew][e], sh->hold_row[ew][e], sh);
    __syncwarp();
    if (lane == 0) sh->hold_cnt[ew] = 0;
    __syncwarp();
}

template <int METRIC>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_i8_tc2_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                   const ScanArgs a, const int q0, const int kchunks, const int stages, const int prefetch_tiles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t *s_q = smem;                                   // [kchunks][128 queries][128 B]
    uint8_t *s_a = smem + (size_t)kchunks * QCHUNK_BYTES;  // [stages][128 rows][128 B]
    Tc2Shared *sh = reinterpret_cast<Tc2Shared *>(s_a + (size_t)stages * STAGE_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + TILE_ROWS - 1) / TILE_ROWS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&sh->full[s], 1);
            tc::mbar_init(&sh->empty[s], 1);
        }
        tc::mbar_init(&sh->q_full, 1);
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&sh->tmem_full[b], 1);
            tc::mbar_init(&sh->tmem_empty[b], 2 * EPI_WARPS);
        }
        for (int w = 0; w < EPI_WARPS; ++w) sh->hold_cnt[w] = 0;
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_rows);
        tc::prefetch_tmap(&tmap_q);
        // this CTA's half of the query tile, resident for the whole kernel
        tc::mbar_expect_tx(&sh->q_full, (uint32_t)kchunks * QCHUNK_BYTES);
        for (int kc = 0; kc < kchunks; ++kc)
            tc::tma_load_2d(s_q + (size_t)kc * QCHUNK_BYTES, &tmap_q, &sh->q_full, kc * CHUNK_BYTES,
                            q0 + (int)rank * QN_CTA);
    }
    if (warp == 1) {
        tc::tmem_alloc_cta2(&sh->tmem_base, TMEM_COLS);
        tc::tmem_relinquish_cta2();
    }
    if (warp >= 2) {
        for (int col = threadIdx.x - 64; col < QN; col += EPI_THREADS) {
            const int q = q0 + col;
            const float thr = q < a.nq ? __ldg(a.topk.thr_f + q) : -__int_as_float(0x7f800000);
            const int bm = q < a.nq ? __ldg(a.q_mag_i + q) : 0;
            sh->thr[col] = thr;
            sh->q_mag[col] = bm;
            sh->tq[col] = prefilter_query_figure<METRIC>(thr, bm);
        }
    }
    __syncthreads();               // barrier inits visible CTA-wide before anyone waits on q_full
    tc::mbar_wait(&sh->q_full, 0);  // own query half landed
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();            // both CTAs: barriers initialised, TMEM allocated, queries resident
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t tile = pair; tile < ntiles; tile += npairs) {
                const int row0 = (int)(a.row_begin + tile * TILE_ROWS + rank * ROWS_PER_CTA);
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    if (rank == 0) tc::mbar_expect_tx(&sh->full[s], 2 * STAGE_BYTES);  // both CTAs' bytes
                    tc::tma_load_2d_cta2(s_a + (size_t)s * STAGE_BYTES, &tmap_rows,
                                         tc::mapa(tc::smem_u32(&sh->full[s]), 0), kc * CHUNK_BYTES, row0);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        // whole warp, warp-uniform descriptors, one elected lane issues (see pkv_scan_tc.cu)
        if (rank == 0) {
            constexpr uint32_t idesc = tc::make_idesc(/*S32*/ 2, /*INT8*/ 1, TILE_ROWS, QN);
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0, t = 0;
            for (uint32_t tile = pair; tile < ntiles; tile += npairs, ++t) {
                const uint32_t buf = t & 1, bph = (t >> 1) & 1;
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * QN;
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
                    const uint64_t a_desc = tc::smem_desc_sw128(tc::smem_u32(s_a) + s * STAGE_BYTES);
                    const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(s_q) + (uint32_t)kc * QCHUNK_BYTES);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < CHUNK_BYTES / 32; ++k)
                            tc::mma_i8_cta2(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                            (kc | k) != 0);
                        tc::mma_commit_cta2(&sh->empty[s]);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
                if (issuer) tc::mma_commit_cta2(&sh->tmem_full[buf]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue (both CTAs, own 128 rows x 256 queries) =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col0 = (ew >> 2) * COLS_PER_WARP;
        const uint32_t row_off = rank * ROWS_PER_CTA + quarter * 32 + lane;
        uint32_t t = 0;
        uint32_t row = a.row_begin + pair * TILE_ROWS + row_off;
        int am = (pair < ntiles && row < a.row_end) ? __ldg(a.row_mag_i + row) : -1;
        const uint32_t empty0 = tc::mapa(tc::smem_u32(&sh->tmem_empty[0]), 0);
        const uint32_t empty1 = tc::mapa(tc::smem_u32(&sh->tmem_empty[1]), 0);
        for (uint32_t tile = pair; tile < ntiles; tile += npairs, ++t) {
            const uint32_t buf = t & 1, bph = (t >> 1) & 1;
            const uint32_t cur_row = row;
            const bool row_ok = am >= 0;
            const int am_min = __reduce_min_sync(0xffffffffu, row_ok ? am : 2147483647);
            const int am_max = __reduce_max_sync(0xffffffffu, row_ok ? am : 0);
            row = a.row_begin + (tile + npairs) * TILE_ROWS + row_off;
            am = (tile + npairs < ntiles && row < a.row_end) ? __ldg(a.row_mag_i + row) : -1;
            const float s_min = sqrtf((float)am_min), s_max = sqrtf((float)am_max), am_min_f = (float)am_min;
#pragma unroll
            for (int i = 0; i < COLS_PER_WARP / 32; ++i) {
                const int c = i * 32 + lane;
                sh->bound[ew][c] = prefilter_bound<METRIC>(sh->tq[col0 + c], s_min, s_max, am_min_f);
            }
            __syncwarp();
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * QN + col0;
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                uint32_t v[32];
                tc::tmem_ld_32x32(taddr + c * 32, v);
                tc::tmem_ld_wait();
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
                        if (d >= sh->bound[ew][c * 32 + j]) hold2<METRIC>(a, q0, col0 + c * 32 + j, d, cur_row, sh, ew);
                    }
                }
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(buf ? empty1 : empty0);
            flush_held2<METRIC>(a, q0, sh, ew, lane, HOLD_FLUSH);
        }
        flush_held2<METRIC>(a, q0, sh, ew, lane, 1);
    }

    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();  // the peer may still be reading this CTA's query half / signalling its barriers
    if (warp == 1) tc::tmem_dealloc_cta2(tmem_base, TMEM_COLS);
}

template <int METRIC>
int launch2(const Index &ix, const ScanArgs &a, const CUtensorMap &mrows, const CUtensorMap &mq, int q0, int kchunks,
            int stages, size_t smem, cudaStream_t s) {
    auto kernel = scan_i8_tc2_kernel<METRIC>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + TILE_ROWS - 1) / TILE_ROWS;
    const uint32_t max_pairs = (uint32_t)ix.sm_count / 2;
    const unsigned pairs = ntiles < max_pairs ? ntiles : max_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PKV_CUDA(cudaLaunchKernelEx(&cfg, kernel, mrows, mq, a, q0, kchunks, stages, ix.opt.tc_prefetch_tiles));
    return PKV_OK;
}

}  // namespace

// One pass of the 2-CTA kernel over [row_begin,row_end) for queries [q0, q0+256).
int launch_scan_tc2_tile(const Index &ix, const ScanArgs &a, const CUtensorMap &mrows, const CUtensorMap &mq, int q0,
                         cudaStream_t s) {
    const int kchunks = ix.dim_pad / CHUNK_BYTES;
    const size_t ctrl = sizeof(Tc2Shared);
    int stages = (int)((227 * 1024 - 1024 - ctrl - (size_t)kchunks * QCHUNK_BYTES) / STAGE_BYTES);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) return fail(PKV_ERR_UNSUPPORTED, "dim %d leaves no room for the row stages", ix.dim);
    const size_t smem = 1024 + (size_t)kchunks * QCHUNK_BYTES + (size_t)stages * STAGE_BYTES + ctrl;
    switch (a.metric) {
        case PKV_COSINE: return launch2<PKV_COSINE>(ix, a, mrows, mq, q0, kchunks, stages, smem, s);
        case PKV_L2: return launch2<PKV_L2>(ix, a, mrows, mq, q0, kchunks, stages, smem, s);
        default: return launch2<PKV_DOT>(ix, a, mrows, mq, q0, kchunks, stages, smem, s);
    }
}

}  // namespace pkv
