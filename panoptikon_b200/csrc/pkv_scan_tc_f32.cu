// pkv_scan_tc_f32.cu — tensor-core scan for floating-point rows (f32 and f16 indexes): a reduced-
// precision tcgen05 contraction as a conservative FILTER, exact f32 re-scoring of the survivors.
//
// Why a filter: at 256 queries per pass the f32 scan needs 2*256 flop per 4 corpus bytes, ~20x what
// the FFMA pipe delivers at HBM speed, so the contraction has to run on tensor cores; but no
// tensor-core input format meets the 1e-5 relative tolerance on the score.  So the tensor-core dot
// product only decides which (row, query) pairs MAY belong to the top-k — with a rigorous error
// bound |dot~ - dot| <= eps * |a| * |q| folded into the threshold — and those pairs (a few thousand
// per query per pass) are re-scored by rescore_kernel from the exact stored rows with the same
// FFMA summation order as the CUDA-core scan, so scores and ids do not depend on which path ran.
//
// Two operand formats (template KIND):
//   KIND_TF32  rows read as stored (f32), kind::tf32, eps = 2.2e-3 (>= 2^-9: 10-bit mantissas)
//   KIND_F16   rows read from an fp16 image: the index's power-of-two-scaled fp16 SHADOW of its
//              f32 rows (half the HBM bytes and twice the tensor rate of tf32, eps = 1.25e-3), or
//              the rows themselves for an f16 index (eps = 2e-4: only accumulation error)
//
// Replaces vec_distance_cosine / vec_distance_L2 over `embeddings.embedding` blobs
// (pql/builder/filters/image_embeddings.rs:321-337, text_embeddings.rs:386-393).
//
// Pipeline per CTA (one per SM): TMA producer warp streams, per 128-byte K-chunk, 128 corpus
// rows AND the matching chunk of NQ queries (queries come from L2: a 256-query tile does not fit
// in shared memory next to the stages); MMA warp issues M=128 x N=NQ UMMAs into a double-buffered
// TMEM accumulator; four epilogue warps threshold it.
//
// Algorithmic bytes per row per pass: the bytes of the image that is scanned (4*D for tf32, 2*D
// for the fp16 image) + 4 (row norm); flops: 2*NQ*D per row.
#include "pkv_tc.cuh"

namespace pkv {

namespace {

constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = 128;  // 32 f32 or 64 f16 components
constexpr int A_BYTES = TILE_M * CHUNK_BYTES;
constexpr int MAX_STAGES = 8;
constexpr int TC_THREADS = 192;
constexpr int KIND_TF32 = 0, KIND_F16 = 1;

template <int NQ>
struct FShared {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    // per query (structure of arrays so that one 128-bit load serves four queries):
    //   qx = threshold in accumulator units (cosine), qy = -2*c (c: accumulator -> true dot),
    //   qz = |q|^2 - thr_f (L2) or -thr_f (DOT), qw = eps*|q|
    alignas(16) float qx[NQ];
    alignas(16) float qy[NQ];
    alignas(16) float qz[NQ];
    alignas(16) float qw[NQ];
};

struct FloatScan {
    const float *q_scale;  // [nq] accumulator -> true dot factor c_q; NULL = 1 (tf32)
    float eps;             // |dot~ - dot| <= eps |a||q|
    float tiny_mag;        // rows with |a|^2 below this always go to re-scoring (fp16 image underflow)
    int prefetch_tiles;
};


// Thresholds one 32-column chunk of the accumulator (thread = row).  t >= 0 (or NaN) <=> the pair
// may be in the top-k; the sign bits are AND-ed so that the common "nothing passes" case costs
// ~2 instructions per element and one branch per 32.
template <int METRIC, int NQ>
__device__ __forceinline__ void threshold_chunk(uint32_t (&v)[32], const FShared<NQ> *sh, int cbase, float rinv,
                                                float bias, float am_b, float sa2, const ScanArgs &a,
                                                const PendDev &pend, int q0, uint32_t row, bool row_ok) {
    unsigned allneg = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 x4 = *reinterpret_cast<const float4 *>(&sh->qx[cbase + j]);
        float t[4];
        if (METRIC == PKV_COSINE) {
            // -acc*rinv + bias <= thr'  <=>  acc*rinv + (thr' - bias) >= 0
            const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) t[e] = fmaf(__uint_as_float(v[j + e]), rinv, xs[e] - bias);
        } else {
            const float4 y4 = *reinterpret_cast<const float4 *>(&sh->qy[cbase + j]);
            const float4 z4 = *reinterpret_cast<const float4 *>(&sh->qz[cbase + j]);
            const float4 w4 = *reinterpret_cast<const float4 *>(&sh->qw[cbase + j]);
            const float ys[4] = {y4.x, y4.y, y4.z, y4.w}, zs[4] = {z4.x, z4.y, z4.z, z4.w},
                        ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float acc = __uint_as_float(v[j + e]);
                if (METRIC == PKV_L2)  // |a|^2 + |q|^2 - 2 dot - slack <= thr
                    t[e] = sa2 * ws[e] - fmaf(acc, ys[e], am_b + zs[e]);
                else  // -dot - slack <= thr
                    t[e] = 0.5f * sa2 * ws[e] - fmaf(acc, 0.5f * ys[e], zs[e] + bias);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[j + e] = __float_as_uint(t[e]);
            allneg &= v[j + e];
        }
    }
    if ((int)allneg >= 0) {  // some element has its sign bit clear
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int q = q0 + cbase + j;
            if (!(__uint_as_float(v[j]) < 0.f) && row_ok && q < a.nq && topk_member(a.topk, q, row)) {
                const uint32_t slot = atomicAdd(pend.cnt + q, 1u);
                if (slot < pend.cap) pend.rows[(size_t)q * pend.cap + slot] = row;
            }
        }
    }
}

template <int KIND, int NQ, int METRIC>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_float_tc_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                     const ScanArgs a, const PendDev pend, const FloatScan fsn, const int q0, const int kchunks,
                     const int stages) {
    constexpr int Q_BYTES = NQ * CHUNK_BYTES;
    constexpr int STAGE = A_BYTES + Q_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    FShared<NQ> *sh = reinterpret_cast<FShared<NQ> *>(smem + (size_t)stages * STAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + TILE_M - 1) / TILE_M;
    const float INF = __int_as_float(0x7f800000);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&sh->full[s], 1);
            tc::mbar_init(&sh->empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&sh->tmem_full[b], 1);
            tc::mbar_init(&sh->tmem_empty[b], 4);
        }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_rows);
        tc::prefetch_tmap(&tmap_q);
    }
    if (warp == 1) {
        tc::tmem_alloc(&sh->tmem_base, 2 * NQ);
        tc::tmem_relinquish();
    }
    if (warp >= 2) {
        for (int col = threadIdx.x - 64; col < NQ; col += 128) {
            const int q = q0 + col;
            float4 v = make_float4(-INF, 0.f, INF, 0.f);  // padded query: keep nothing
            if (q < a.nq) {
                const float bm = __ldg(a.q_mag_f + q);
                const float thr = __ldg(a.topk.thr_f + q);
                const float c = fsn.q_scale ? __ldg(fsn.q_scale + q) : 1.0f;
                v.x = thr / c;  // -acc*c*rinv <= thr  <=>  -acc*rinv <= thr/c  (c > 0; inf stays inf)
                if (v.x != v.x) v.x = INF;
                v.y = -2.0f * c;
                v.z = METRIC == PKV_L2 ? bm - thr : -thr;  // -inf while there is no threshold
                if (v.z != v.z) v.z = -INF;
                v.w = fsn.eps * sqrtf(bm);
            }
            sh->qx[col] = v.x;
            sh->qy[col] = v.y;
            sh->qz[col] = v.z;
            sh->qw[col] = v.w;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row0 = (int)(a.row_begin + tile * TILE_M);
                const uint32_t ptile = tile + (uint32_t)fsn.prefetch_tiles * gridDim.x;
                if (fsn.prefetch_tiles > 0 && ptile < ntiles)
                    for (int kc = 0; kc < kchunks; ++kc)
                        tc::tma_prefetch_2d(&tmap_rows, kc * CHUNK_BYTES, (int)(a.row_begin + ptile * TILE_M));
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    tc::mbar_expect_tx(&sh->full[s], STAGE);
                    uint8_t *st = smem + (size_t)s * STAGE;
                    tc::tma_load_2d(st, &tmap_rows, &sh->full[s], kc * CHUNK_BYTES, row0);
                    tc::tma_load_2d(st + A_BYTES, &tmap_q, &sh->full[s], kc * CHUNK_BYTES, q0);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform descriptors, one elected lane issues (see pkv_scan_tc.cu)
        {
            constexpr uint32_t idesc = KIND == KIND_TF32 ? tc::make_idesc(/*F32*/ 1, /*TF32*/ 2, TILE_M, NQ)
                                                         : tc::make_idesc(/*F32*/ 1, /*F16*/ 0, TILE_M, NQ);
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0, t = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1, bph = (t >> 1) & 1;
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * NQ;
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
                    const uint64_t a_desc = tc::smem_desc_sw128(tc::smem_u32(smem) + s * (uint32_t)STAGE);
                    const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(smem) + s * (uint32_t)STAGE + A_BYTES);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < CHUNK_BYTES / 32; ++k) {  // K = 32 bytes per UMMA: 8 tf32 / 16 f16
                            if (KIND == KIND_TF32)
                                tc::mma_tf32(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                            else
                                tc::mma_f16(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kc | k) != 0);
                        }
                        tc::mma_commit(&sh->empty[s]);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
                if (issuer) tc::mma_commit(&sh->tmem_full[buf]);
                __syncwarp();
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t t = 0;
        for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
            const uint32_t buf = t & 1, bph = (t >> 1) & 1;
            const uint32_t row = a.row_begin + tile * TILE_M + quarter * 32 + lane;
            const bool row_ok = row < a.row_end;
            const float am = row_ok ? __ldg(a.row_mag_f + row) : 0.f;
            // rows whose norm is too small for the fp16 image's relative error bound always pass
            const float bias = (am < fsn.tiny_mag) ? -INF : 0.f;
            const float rinv = rsqrtf(am);
            const float sa2 = 2.0f * sqrtf(am);
            const float am_b = am + bias;
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * NQ;
#pragma unroll 1
            for (int c = 0; c < NQ / 32; ++c) {
                uint32_t v[32];
                tc::tmem_ld_32x32(taddr + c * 32, v);
                tc::tmem_ld_wait();
                threshold_chunk<METRIC, NQ>(v, sh, c * 32, rinv, bias, am_b, sa2, a, pend, q0, row, row_ok);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sh->tmem_empty[buf]);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 2 * NQ);
}

// ---------------------------------------------------------------------------------------------
// 2-CTA (cta_group::2) variant for 256 queries per pass: M=256 rows x N=256 queries per UMMA; each
// CTA of the pair streams its own 128 rows AND its own 128-query half of every K-chunk, so the
// per-SM L2->smem traffic per corpus byte is halved versus the 1-CTA kernel (which is L2-bound at
// NQ=256).  Barrier topology as in pkv_scan_tc2.cu.
constexpr int P_EPI_WARPS = 16;
constexpr int P_THREADS = 64 + P_EPI_WARPS * 32;
constexpr int P_NQ = 256;
constexpr int P_STAGE = A_BYTES + 128 * CHUNK_BYTES;  // 16 KB rows + 16 KB query half per CTA

template <int KIND, int METRIC>
__global__ void __launch_bounds__(P_THREADS, 1)
scan_float_tc2_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                      const ScanArgs a, const PendDev pend, const FloatScan fsn, const int q0, const int kchunks,
                      const int stages) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    FShared<P_NQ> *sh = reinterpret_cast<FShared<P_NQ> *>(smem + (size_t)stages * P_STAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + 2 * TILE_M - 1) / (2 * TILE_M);
    const float INF = __int_as_float(0x7f800000);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&sh->full[s], 1);
            tc::mbar_init(&sh->empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc::mbar_init(&sh->tmem_full[b], 1);
            tc::mbar_init(&sh->tmem_empty[b], 2 * P_EPI_WARPS);
        }
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_rows);
        tc::prefetch_tmap(&tmap_q);
    }
    if (warp == 1) {
        tc::tmem_alloc_cta2(&sh->tmem_base, 2 * P_NQ);
        tc::tmem_relinquish_cta2();
    }
    if (warp >= 2) {
        for (int col = threadIdx.x - 64; col < P_NQ; col += P_EPI_WARPS * 32) {
            const int q = q0 + col;
            float4 v = make_float4(-INF, 0.f, INF, 0.f);
            if (q < a.nq) {
                const float bm = __ldg(a.q_mag_f + q);
                const float thr = __ldg(a.topk.thr_f + q);
                const float c = fsn.q_scale ? __ldg(fsn.q_scale + q) : 1.0f;
                v.x = thr / c;
                if (v.x != v.x) v.x = INF;
                v.y = -2.0f * c;
                v.z = METRIC == PKV_L2 ? bm - thr : -thr;
                if (v.z != v.z) v.z = -INF;
                v.w = fsn.eps * sqrtf(bm);
            }
            sh->qx[col] = v.x;
            sh->qy[col] = v.y;
            sh->qz[col] = v.z;
            sh->qw[col] = v.w;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t tile = pair; tile < ntiles; tile += npairs) {
                const int row0 = (int)(a.row_begin + tile * 2 * TILE_M + rank * TILE_M);
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    if (rank == 0) tc::mbar_expect_tx(&sh->full[s], 2 * P_STAGE);
                    const uint32_t full0 = tc::mapa(tc::smem_u32(&sh->full[s]), 0);
                    uint8_t *st = smem + (size_t)s * P_STAGE;
                    tc::tma_load_2d_cta2(st, &tmap_rows, full0, kc * CHUNK_BYTES, row0);
                    tc::tma_load_2d_cta2(st + A_BYTES, &tmap_q, full0, kc * CHUNK_BYTES, q0 + (int)rank * 128);
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform descriptors, one elected lane issues (see pkv_scan_tc.cu)
        if (rank == 0) {
            constexpr uint32_t idesc = KIND == KIND_TF32 ? tc::make_idesc(/*F32*/ 1, /*TF32*/ 2, 2 * TILE_M, P_NQ)
                                                         : tc::make_idesc(/*F32*/ 1, /*F16*/ 0, 2 * TILE_M, P_NQ);
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0, t = 0;
            for (uint32_t tile = pair; tile < ntiles; tile += npairs, ++t) {
                const uint32_t buf = t & 1, bph = (t >> 1) & 1;
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * P_NQ;
                for (int kc = 0; kc < kchunks; ++kc) {
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
                    const uint64_t a_desc = tc::smem_desc_sw128(tc::smem_u32(smem) + s * (uint32_t)P_STAGE);
                    const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(smem) + s * (uint32_t)P_STAGE + A_BYTES);
                    if (issuer) {
#pragma unroll
                        for (int k = 0; k < CHUNK_BYTES / 32; ++k) {
                            if (KIND == KIND_TF32)
                                tc::mma_tf32_cta2(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                                  (kc | k) != 0);
                            else
                                tc::mma_f16_cta2(d_tmem, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                                 (kc | k) != 0);
                        }
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
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col0 = (ew >> 2) * (P_NQ / 4);
        const uint32_t empty0 = tc::mapa(tc::smem_u32(&sh->tmem_empty[0]), 0);
        const uint32_t empty1 = tc::mapa(tc::smem_u32(&sh->tmem_empty[1]), 0);
        uint32_t t = 0;
        for (uint32_t tile = pair; tile < ntiles; tile += npairs, ++t) {
            const uint32_t buf = t & 1, bph = (t >> 1) & 1;
            const uint32_t row = a.row_begin + tile * 2 * TILE_M + rank * TILE_M + quarter * 32 + lane;
            const bool row_ok = row < a.row_end;
            const float am = row_ok ? __ldg(a.row_mag_f + row) : 0.f;
            const float bias = (am < fsn.tiny_mag) ? -INF : 0.f;
            const float rinv = rsqrtf(am);
            const float sa2 = 2.0f * sqrtf(am);
            const float am_b = am + bias;
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * P_NQ + col0;
#pragma unroll 1
            for (int c = 0; c < P_NQ / 4 / 32; ++c) {
                uint32_t v[32];
                tc::tmem_ld_32x32(taddr + c * 32, v);
                tc::tmem_ld_wait();
                threshold_chunk<METRIC, P_NQ>(v, sh, col0 + c * 32, rinv, bias, am_b, sa2, a, pend, q0, row, row_ok);
            }
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(buf ? empty1 : empty0);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();
    if (warp == 1) tc::tmem_dealloc_cta2(tmem_base, 2 * P_NQ);
}

// Exact re-scoring of the pending (row, query) pairs from the STORED rows: one warp per pair, lanes
// stride the components with the same order and the same xor-reduction tree as the CUDA-core scan.
template <int METRIC, bool ROWS_F16>
__global__ void __launch_bounds__(256) rescore_kernel(const ScanArgs a, const PendDev pend, SearchStatus *status) {
    extern __shared__ float4 s_q[];  // one query, dim_pad/4 float4 (f32, widened for an f16 index)
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int nvec = a.dim_pad >> 2;
    const uint32_t raw = pend.cnt[q];
    const uint32_t n = raw < pend.cap ? raw : pend.cap;
    if (blockIdx.y == 0 && threadIdx.x == 0 && raw > pend.cap) {
        atomicOr(&status->any_overflow, 1u);
        atomicOr(&status->sticky_overflow, 1u);
    }
    if (n == 0) return;
    if (blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&status->rescored, n);  // search statistics (pkv_counters)
    const float4 *gq = (const float4 *)a.queries + (size_t)q * nvec;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) s_q[i] = gq[i];
    __syncthreads();
    const uint32_t wstride = gridDim.y * (blockDim.x >> 5);
    const float qmag = __ldg(a.q_mag_f + q);
    for (uint32_t e = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += wstride) {
        const uint32_t row = pend.rows[(size_t)q * pend.cap + e];
        const uint8_t *rb = (const uint8_t *)a.data + (size_t)row * (size_t)a.pitch_bytes;
        float acc = 0.f, nrm = 0.f;
        if (!ROWS_F16) {
            const float4 *rp = (const float4 *)rb;
            auto body = [&](const float4 av, const float4 qv) {
                if (METRIC == PKV_L2) {
                    float t;
                    t = av.x - qv.x; acc = fmaf(t, t, acc);
                    t = av.y - qv.y; acc = fmaf(t, t, acc);
                    t = av.z - qv.z; acc = fmaf(t, t, acc);
                    t = av.w - qv.w; acc = fmaf(t, t, acc);
                } else {
                    acc = fmaf(av.x, qv.x, acc);
                    acc = fmaf(av.y, qv.y, acc);
                    acc = fmaf(av.z, qv.z, acc);
                    acc = fmaf(av.w, qv.w, acc);
                }
                if (METRIC == PKV_COSINE) {
                    nrm = fmaf(av.x, av.x, nrm);
                    nrm = fmaf(av.y, av.y, nrm);
                    nrm = fmaf(av.z, av.z, nrm);
                    nrm = fmaf(av.w, av.w, nrm);
                }
            };
            // the row's first 8 loads per lane (D <= 1024: all of them) are issued before any is consumed: one memory
            // latency per gathered row; the element order of the sums is unchanged
            float4 rv[8];
#pragma unroll
            for (int it = 0; it < 8; ++it)
                if (lane + it * 32 < nvec) rv[it] = __ldg(rp + lane + it * 32);
#pragma unroll
            for (int it = 0; it < 8; ++it)
                if (lane + it * 32 < nvec) body(rv[it], s_q[lane + it * 32]);
            for (int j = lane + 256; j < nvec; j += 32) body(__ldg(rp + j), s_q[j]);
        } else {
            // same element order as scan_f16_simt_kernel: 8 halfs per lane per step
            const uint4 *rp = (const uint4 *)rb;
            const int nvec8 = a.dim_pad >> 3;
            for (int j = lane; j < nvec8; j += 32) {
                const uint4 raw8 = __ldg(rp + j);
                const __half2 *h = reinterpret_cast<const __half2 *>(&raw8);
                float av[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 p = __half22float2(h[i]);
                    av[2 * i] = p.x;
                    av[2 * i + 1] = p.y;
                }
                const float4 q0v = s_q[2 * j], q1v = s_q[2 * j + 1];
                const float qq[8] = {q0v.x, q0v.y, q0v.z, q0v.w, q1v.x, q1v.y, q1v.z, q1v.w};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (METRIC == PKV_L2) {
                        const float t = av[i] - qq[i];
                        acc = fmaf(t, t, acc);
                    } else {
                        acc = fmaf(av[i], qq[i], acc);
                    }
                }
                if (METRIC == PKV_COSINE) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) nrm = fmaf(av[i], av[i], nrm);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (METRIC == PKV_COSINE) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
        }
        if (lane == 0) {
            float d;
            if (METRIC == PKV_COSINE)
                d = cosine_key((double)acc, (double)nrm, (double)qmag);
            else if (METRIC == PKV_L2)
                d = l2_key_from_sum(acc);
            else
                d = -acc;
            topk_push(a.topk, q, row, d);
        }
    }
}

template <int KIND, int NQ, int METRIC>
int launch_one(const Index &ix, const ScanArgs &a, const PendDev &pend, const FloatScan &fsn, const CUtensorMap &mrows,
               const CUtensorMap &mq, int q0, int kchunks, cudaStream_t s) {
    constexpr int STAGE = A_BYTES + NQ * CHUNK_BYTES;
    const size_t ctrl = sizeof(FShared<NQ>);
    int stages = (int)((227 * 1024 - 1024 - ctrl) / STAGE);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const size_t smem = 1024 + (size_t)stages * STAGE + ctrl;
    auto kernel = scan_float_tc_kernel<KIND, NQ, METRIC>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + TILE_M - 1) / TILE_M;
    const unsigned grid = ntiles < (uint32_t)ix.sm_count ? ntiles : (unsigned)ix.sm_count;
    kernel<<<grid, TC_THREADS, smem, s>>>(mrows, mq, a, pend, fsn, q0, kchunks, stages);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

template <int KIND, int METRIC>
int launch_pair(const Index &ix, const ScanArgs &a, const PendDev &pend, const FloatScan &fsn, const CUtensorMap &mrows,
                const CUtensorMap &mq128, int q0, int kchunks, cudaStream_t s) {
    const size_t ctrl = sizeof(FShared<P_NQ>);
    int stages = (int)((227 * 1024 - 1024 - ctrl) / P_STAGE);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const size_t smem = 1024 + (size_t)stages * P_STAGE + ctrl;
    auto kernel = scan_float_tc2_kernel<KIND, METRIC>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + 2 * TILE_M - 1) / (2 * TILE_M);
    const uint32_t max_pairs = (uint32_t)ix.sm_count / 2;
    const unsigned pairs = ntiles < max_pairs ? ntiles : max_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PKV_CUDA(cudaLaunchKernelEx(&cfg, kernel, mrows, mq128, a, pend, fsn, q0, kchunks, stages));
    return PKV_OK;
}

template <int KIND, int METRIC>
int launch_metric(const Index &ix, const ScanArgs &a, const PendDev &pend, const FloatScan &fsn,
                  const CUtensorMap &mrows, const CUtensorMap &mq128, const CUtensorMap &mq256, int kchunks,
                  SearchStatus *status, cudaStream_t s, int *launches) {
    for (int q0 = 0; q0 < a.nq;) {
        const int left = a.nq - q0;
        *launches += 1;
        if (left > 128 && ix.opt.tc_cta2 && (ix.sm_count % 2) == 0) {
            PKV_TRY((launch_pair<KIND, METRIC>(ix, a, pend, fsn, mrows, mq128, q0, kchunks, s)));
            q0 += 256;
        } else if (left > 128) {
            PKV_TRY((launch_one<KIND, 256, METRIC>(ix, a, pend, fsn, mrows, mq256, q0, kchunks, s)));
            q0 += 256;
        } else {
            PKV_TRY((launch_one<KIND, 128, METRIC>(ix, a, pend, fsn, mrows, mq128, q0, kchunks, s)));
            q0 += 128;
        }
    }
    PKV_TRY(launch_rescore(ix, a, pend, status, s));
    *launches += 1;
    return PKV_OK;
}

template <int KIND>
int launch_kind(const Index &ix, const ScanArgs &a, const PendDev &pend, const FloatScan &fsn, const CUtensorMap &mrows,
                const CUtensorMap &mq128, const CUtensorMap &mq256, int kchunks, SearchStatus *status, cudaStream_t s,
                int *launches) {
    switch (a.metric) {
        case PKV_COSINE:
            return launch_metric<KIND, PKV_COSINE>(ix, a, pend, fsn, mrows, mq128, mq256, kchunks, status, s, launches);
        case PKV_L2:
            return launch_metric<KIND, PKV_L2>(ix, a, pend, fsn, mrows, mq128, mq256, kchunks, status, s, launches);
        default:
            return launch_metric<KIND, PKV_DOT>(ix, a, pend, fsn, mrows, mq128, mq256, kchunks, status, s, launches);
    }
}

// ---- fp16 image helpers ----
// q16[q] = half(q * 2^e_q) with 2^e_q the power of two that puts |q|_inf * 2^e_q in [4096, 8192);
// q_scale[q] = 1 / (row_scale * 2^e_q) turns an accumulator into the true dot product.
__global__ void prep_queries_f16_kernel(const float *q, int dim, int dim_pad, int dim_pad_h, float row_scale,
                                        __half *q16, float *q_scale) {
    const int qi = blockIdx.x;
    __shared__ float s_max[32];
    const float *src = q + (size_t)qi * dim_pad;
    float m = 0.f;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) {
        const float v = fabsf(src[i]);
        if (v > m && v < __int_as_float(0x7f800000)) m = v;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0.f;
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) s_max[0] = m;
    }
    __syncthreads();
    m = s_max[0];
    int e = 0;
    if (m > 0.f) {
        e = 13 - (ilogbf(m) + 1);
        if (e > 100) e = 100;
        if (e < -100) e = -100;
    }
    const float qs = ldexpf(1.0f, e);
    for (int i = threadIdx.x; i < dim_pad_h; i += blockDim.x)
        q16[(size_t)qi * dim_pad_h + i] = __float2half_rn(i < dim ? src[i] * qs : 0.f);
    if (threadIdx.x == 0) q_scale[qi] = 1.0f / (row_scale * qs);
}

__global__ void shadow_convert_kernel(const float *rows, int64_t pitch_f, int dim, int dim_pad_h, float scale,
                                      int64_t row_begin, int64_t row_end, __half *out) {
    const int64_t total = (row_end - row_begin) * dim_pad_h;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_begin + i / dim_pad_h;
        const int c = (int)(i % dim_pad_h);
        out[r * dim_pad_h + c] = __float2half_rn(c < dim ? rows[r * pitch_f + c] * scale : 0.f);
    }
}

}  // namespace

// Exact re-scoring of everything the filter kernels parked in `pend` (one launch for all queries).
int launch_rescore(const Index &ix, const ScanArgs &a, const PendDev &pend, SearchStatus *status, cudaStream_t s) {
    const size_t smem = (size_t)a.dim_pad * 4;
    int ry = (8 * ix.sm_count + a.nq - 1) / a.nq;  // ~8 CTAs (64 warps) per SM in total: the gathers are latency-bound
    if (ry < 1) ry = 1;
    if (ry > 64) ry = 64;
    const dim3 grid((unsigned)a.nq, (unsigned)ry);
#define PKV_RESCORE(M)                                                                                   \
    do {                                                                                                 \
        if (ix.dtype == PKV_F16) {                                                                       \
            auto rk = rescore_kernel<M, true>;                                                           \
            PKV_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            rk<<<grid, 256, smem, s>>>(a, pend, status);                                                 \
        } else {                                                                                         \
            auto rk = rescore_kernel<M, false>;                                                          \
            PKV_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            rk<<<grid, 256, smem, s>>>(a, pend, status);                                                 \
        }                                                                                                \
    } while (0)
    if (a.metric == PKV_COSINE) PKV_RESCORE(PKV_COSINE);
    else if (a.metric == PKV_L2) PKV_RESCORE(PKV_L2);
    else PKV_RESCORE(PKV_DOT);
#undef PKV_RESCORE
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

// TF32 keeps 10 explicit mantissa bits; whether the tensor core truncates or rounds the f32
// operands, each is off by < 2^-10 relative, a product by < 2^-9, so
// |dot~ - dot| <= 2^-9 * sum|a_i q_i| <= 2^-9 |a||q| (Cauchy-Schwarz); f32 accumulation adds ~1e-6.
static constexpr float TF32_EPS = 2.2e-3f;
// fp16 image, round-to-nearest of power-of-two-scaled values: each normal operand is off by <= 2^-11
// relative, a product by < 2^-10 (9.8e-4); components that fall into the fp16 subnormal range add
// at most 2^-25*sqrt(D)/(scale*|a|) relative, < 2^-13 for every row that is not flagged "tiny"
// (those are always re-scored); queries are scaled per query so the same holds for them.
static constexpr float F16_SHADOW_EPS = 1.25e-3f;
// true f16 index: operands are exact, products are exact in f32, only the accumulation rounds
static constexpr float F16_EXACT_EPS = 2.0e-4f;

// which operand the filter scans: 2 int8 image, 1 fp16 image, 0 the rows themselves (tf32 / f16)
static int image_choice(const Index &ix, int nq) {
    const int want = ix.opt.use_shadow;
    const bool f16_img = ix.dtype == PKV_F32 && ix.d_shadow && ix.shadow_scale != 0.f;
    // best available: the int8 image halves the bytes of a pass but its error band passes ~5x more rows to the
    // re-scorer than the fp16 image; past img8_max_queries per batch the fp16 image (if built) is the faster filter
    if (want < 0 && img8_usable(ix) && f16_img && nq > ix.opt.img8_max_queries) return 1;
    if ((want == 2 || want < 0) && img8_usable(ix)) return 2;
    if ((want == 1 || want < 0 || want == 2) && ix.dtype == PKV_F32 && ix.d_shadow && ix.shadow_scale != 0.f) return 1;
    return 0;
}
static bool use_shadow(const Index &ix, int nq) { return image_choice(ix, nq) == 1; }

bool scan_tc_f32_supported(const Index &ix, int nq) {
    if (ix.opt.force_simt) return false;
    // with the fp16 image the tensor-core kernel reads half the bytes of the FFMA kernel, so it
    // wins from the very first query; the tf32 path reads the f32 rows and only pays off once the
    // FFMA kernel would need a second pass over the corpus
    if (image_choice(ix, nq) != 0) return nq >= ix.opt.tc_min_queries_img;
    if (ix.dtype == PKV_F32 || ix.dtype == PKV_F16) return nq >= ix.opt.tc_min_queries_f32;
    return false;
}


static float float_eps(const Index &ix, int nq) {
    if (image_choice(ix, nq) == 2) return 0.f;  // the int8-image kernel applies its own per-pair bound
    if (ix.dtype == PKV_F16) return F16_EXACT_EPS;
    return use_shadow(ix, nq) ? F16_SHADOW_EPS : TF32_EPS;
}

FilterSpec filter_spec_tc_f32(const Index &ix, int metric, int nq) {
    FilterSpec fs = filter_spec_simt(PKV_F32, metric);
    if (metric == PKV_COSINE) fs.abs = float_eps(ix, nq);  // in units of dot/|a|: scaled by |q| in filter_threshold
    return fs;  // L2 / DOT apply the per-row bound inside the kernel
}

int scan_tc_f32_kind(const Index &ix, int nq) {
    if (image_choice(ix, nq) == 2) return 8;
    return ix.dtype == PKV_F16 ? 6 : (use_shadow(ix, nq) ? 7 : 4);
}

int build_shadow(Index &ix, int64_t row_begin, int64_t row_end, cudaStream_t s) {
    if (row_end <= row_begin) return PKV_OK;
    const int64_t total = (row_end - row_begin) * ix.dim_pad_h;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    shadow_convert_kernel<<<(unsigned)blocks, 256, 0, s>>>((const float *)ix.d_data, ix.pitch / 4, ix.dim, ix.dim_pad_h,
                                                          ix.shadow_scale, row_begin, row_end, ix.d_shadow);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_scan_tc_f32(const Index &ix, const ScanArgs &a, Workspace &ws, cudaStream_t s, int *launches) {
    if (a.row_end <= a.row_begin || a.nq <= 0) return PKV_OK;
    if (image_choice(ix, a.nq) == 2) return launch_scan_img8(ix, a, ws, s, launches);
    PendDev pend{ws.d_pend_rows, ws.d_pend_cnt, (uint32_t)ws.pend_cap, nullptr, nullptr};
    PKV_CUDA(cudaMemsetAsync(ws.d_pend_cnt, 0, sizeof(uint32_t) * a.nq, s));
    FloatScan fsn;
    fsn.eps = float_eps(ix, a.nq);
    fsn.prefetch_tiles = ix.opt.tc_prefetch_tiles;
    fsn.q_scale = nullptr;
    fsn.tiny_mag = 0.f;
    CUtensorMap mrows, mq128, mq256;
    // Query maps span whole 128-row tiles of the (larger) workspace buffer: an out-of-bounds TMA box is
    // zero-filled row by row and measurably slower; rows past nq hold stale queries whose accumulator
    // columns the epilogue ignores.
    uint64_t qrows = ((uint64_t)a.nq + 127) / 128 * 128;
    if (qrows > (uint64_t)ws.nq_cap) qrows = (uint64_t)ws.nq_cap;
    if (qrows < (uint64_t)a.nq) qrows = (uint64_t)a.nq;
    const bool shadow = use_shadow(ix, a.nq);
    if (ix.dtype == PKV_F32 && !shadow) {
        PKV_TRY(make_tmap_bytes(&mrows, ix.d_data, (uint64_t)ix.pitch, (uint64_t)ix.sealed_rows, (uint64_t)ix.pitch, TILE_M));
        PKV_TRY(make_tmap_bytes(&mq128, a.queries, (uint64_t)ix.pitch, qrows, (uint64_t)ix.pitch, 128));
        PKV_TRY(make_tmap_bytes(&mq256, a.queries, (uint64_t)ix.pitch, qrows, (uint64_t)ix.pitch, 256));
        return launch_kind<KIND_TF32>(ix, a, pend, fsn, mrows, mq128, mq256, (int)(ix.pitch / CHUNK_BYTES), ws.d_status, s,
                                      launches);
    }
    // fp16 image: the shadow of an f32 index, or the rows of an f16 index
    const uint64_t pitch_h = (uint64_t)ix.dim_pad_h * 2;
    const void *img = shadow ? (const void *)ix.d_shadow : (const void *)ix.d_data;
    const float row_scale = shadow ? ix.shadow_scale : 1.0f;
    prep_queries_f16_kernel<<<a.nq, 128, 0, s>>>((const float *)a.queries, ix.dim, ix.dim_pad, ix.dim_pad_h, row_scale,
                                                 ws.d_q16, ws.d_q_scale);
    PKV_CUDA(cudaGetLastError());
    *launches += 1;
    fsn.q_scale = ws.d_q_scale;
    if (shadow) {
        // |a| * scale below sqrt(D) * 2^-12: components sit in or near the fp16 subnormal range
        const float lim = sqrtf((float)ix.dim) * 0.000244140625f / ix.shadow_scale;
        fsn.tiny_mag = lim * lim;
    }
    PKV_TRY(make_tmap_bytes(&mrows, img, pitch_h, (uint64_t)ix.sealed_rows, pitch_h, TILE_M));
    PKV_TRY(make_tmap_bytes(&mq128, ws.d_q16, pitch_h, qrows, pitch_h, 128));
    PKV_TRY(make_tmap_bytes(&mq256, ws.d_q16, pitch_h, qrows, pitch_h, 256));
    return launch_kind<KIND_F16>(ix, a, pend, fsn, mrows, mq128, mq256, (int)(pitch_h / CHUNK_BYTES), ws.d_status, s,
                                 launches);
}

}  // namespace pkv
