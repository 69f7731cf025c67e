// pkv_scan_tc_f32.cu — tensor-core scan for f32 rows: tcgen05.mma kind::tf32 as a conservative
// FILTER, exact f32 re-scoring of the survivors.
//
// Why a filter: at 256 queries per pass the f32 scan needs 2*256 flop per 4 corpus bytes, ~20x what
// the FFMA pipe delivers at HBM speed, so the contraction has to run on tensor cores; but TF32
// keeps 10 mantissa bits and cannot meet the 1e-5 relative tolerance on the score.  So the TF32
// dot product only decides which (row, query) pairs MAY belong to the top-k — with a rigorous
// error bound |dot~ - dot| <= eps * |a| * |b| folded into the threshold — and those pairs (a few
// thousand per query per pass) are re-scored by rescore_f32_kernel with the same FFMA summation
// order as the CUDA-core scan, so scores and ids do not depend on which path ran.
//
// Replaces vec_distance_cosine / vec_distance_L2 over `embeddings.embedding` blobs
// (pql/builder/filters/image_embeddings.rs:321-337, text_embeddings.rs:386-393).
//
// Pipeline per CTA (one per SM): TMA producer warp streams, per 128-byte K-chunk, 128 corpus
// rows AND the matching chunk of NQ queries (queries come from L2: a 256-query f32 tile is
// 786 KB and cannot stay resident in shared memory); MMA warp issues M=128 x N=NQ x K=8 UMMAs
// into a double-buffered TMEM accumulator; four epilogue warps threshold it.
//
// Algorithmic bytes per row per pass: dim_pad*4 (+4 row norm); flops: 2*NQ*dim_pad per row.
#include "pkv_tc.cuh"

namespace pkv {

namespace {

constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = 128;  // 32 f32 components
constexpr int A_BYTES = TILE_M * CHUNK_BYTES;
constexpr int MAX_STAGES = 8;
constexpr int TC_THREADS = 192;

template <int NQ>
struct F32Shared {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
    uint32_t pad;
    alignas(16) float thr[NQ];   // filter threshold per query (see FilterSpec)
    alignas(16) float qm[NQ];    // |q|^2
    alignas(16) float qs[NQ];    // eps * |q|  (error-bound scale)
};

struct PendDev {
    uint32_t *rows;  // [nq][cap] rows awaiting exact re-scoring
    uint32_t *cnt;   // [nq]
    uint32_t cap;
};

template <int NQ, int METRIC>
__global__ void __launch_bounds__(TC_THREADS, 1)
scan_f32_tc_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_q,
                   const ScanArgs a, const PendDev pend, const int q0, const int kchunks, const int stages,
                   const float eps, const int prefetch_tiles) {
    constexpr int Q_BYTES = NQ * CHUNK_BYTES;
    constexpr int STAGE = A_BYTES + Q_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    F32Shared<NQ> *sh = reinterpret_cast<F32Shared<NQ> *>(smem + (size_t)stages * STAGE);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + TILE_M - 1) / TILE_M;

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
            const bool ok = q < a.nq;
            const float bm = ok ? __ldg(a.q_mag_f + q) : 0.f;
            sh->thr[col] = ok ? __ldg(a.topk.thr_f + q) : -__int_as_float(0x7f800000);
            sh->qm[col] = bm;
            sh->qs[col] = eps * sqrtf(bm);
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int row0 = (int)(a.row_begin + tile * TILE_M);
                const uint32_t ptile = tile + (uint32_t)prefetch_tiles * gridDim.x;
                if (prefetch_tiles > 0 && ptile < ntiles)
                    for (int kc = 0; kc < kchunks; ++kc)
                        tc::tma_prefetch_2d(&tmap_rows, kc * CHUNK_BYTES, (int)(a.row_begin + ptile * TILE_M));
                for (int kc = 0; kc < kchunks; ++kc, ++it) {
                    const uint32_t s = it % stages, ph = (it / stages) & 1;
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    tc::mbar_expect_tx(&sh->full[s], STAGE);
                    uint8_t *st = smem + (size_t)s * STAGE;
                    tc::tma_load_2d(st, &tmap_rows, &sh->full[s], kc * CHUNK_BYTES, row0);
                    tc::tma_load_2d(st + A_BYTES, &tmap_q, &sh->full[s], kc * CHUNK_BYTES, q0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = tc::make_idesc(/*F32*/ 1, /*TF32*/ 2, TILE_M, NQ);
            uint32_t it = 0, t = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++t) {
                const uint32_t buf = t & 1, bph = (t >> 1) & 1;
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * NQ;
                for (int kc = 0; kc < kchunks; ++kc, ++it) {
                    const uint32_t s = it % stages, ph = (it / stages) & 1;
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
                    const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * STAGE);
                    const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
                    for (int k = 0; k < CHUNK_BYTES / 32; ++k) {
                        tc::mma_tf32(d_tmem, tc::smem_desc_sw128(a_addr + k * 32),
                                     tc::smem_desc_sw128(b_addr + k * 32), idesc, (kc | k) != 0);
                    }
                    tc::mma_commit(&sh->empty[s]);
                }
                tc::mma_commit(&sh->tmem_full[buf]);
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
            const float rinv = rsqrtf(am);
            const float sa2 = 2.0f * sqrtf(am);
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * NQ;
#pragma unroll 1
            for (int c = 0; c < NQ / 32; ++c) {
                uint32_t v[32];
                tc::tmem_ld_32x32(taddr + c * 32, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 th = *reinterpret_cast<const float4 *>(&sh->thr[c * 32 + j]);
                    const float thv[4] = {th.x, th.y, th.z, th.w};
                    bool pass[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float dot = __uint_as_float(v[j + e]);
                        float f;
                        if (METRIC == PKV_COSINE) {
                            f = -dot * rinv;  // the eps slack is folded into thr (FilterSpec.abs)
                        } else if (METRIC == PKV_L2) {
                            // lower bound of the squared distance given |dot~ - dot| <= eps|a||q|
                            f = fmaf(-2.0f, dot, am + sh->qm[c * 32 + j + e]) - sa2 * sh->qs[c * 32 + j + e];
                        } else {
                            f = -dot - 0.5f * sa2 * sh->qs[c * 32 + j + e];
                        }
                        pass[e] = !(f > thv[e]);
                    }
                    if (pass[0] | pass[1] | pass[2] | pass[3]) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int q = q0 + c * 32 + j + e;
                            if (pass[e] && row_ok && q < a.nq && topk_member(a.topk, q, row)) {
                                const uint32_t slot = atomicAdd(pend.cnt + q, 1u);
                                if (slot < pend.cap) pend.rows[(size_t)q * pend.cap + slot] = row;
                            }
                        }
                    }
                }
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

// Exact re-scoring of the pending (row, query) pairs: one warp per pair, lanes stride the
// components with the same float4 order and the same xor-reduction tree as scan_f32_simt_kernel.
template <int METRIC>
__global__ void __launch_bounds__(256) rescore_f32_kernel(const ScanArgs a, const PendDev pend, SearchStatus *status) {
    extern __shared__ float4 s_q[];  // one query, dim_pad/4 float4
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int nvec = a.dim_pad >> 2;
    const uint32_t raw = pend.cnt[q];
    const uint32_t n = raw < pend.cap ? raw : pend.cap;
    if (blockIdx.y == 0 && threadIdx.x == 0 && raw > pend.cap) atomicOr(&status->any_overflow, 1u);
    if (n == 0) return;
    const float4 *gq = (const float4 *)a.queries + (size_t)q * nvec;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) s_q[i] = gq[i];
    __syncthreads();
    const uint32_t wstride = gridDim.y * (blockDim.x >> 5);
    const float qmag = __ldg(a.q_mag_f + q);
    for (uint32_t e = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5); e < n; e += wstride) {
        const uint32_t row = pend.rows[(size_t)q * pend.cap + e];
        const float4 *rp = (const float4 *)((const uint8_t *)a.data + (size_t)row * (size_t)a.pitch_bytes);
        float acc = 0.f, nrm = 0.f;
        for (int j = lane; j < nvec; j += 32) {
            const float4 av = __ldg(rp + j);
            const float4 qv = s_q[j];
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int NQ, int METRIC>
int launch_one(const Index &ix, const ScanArgs &a, const PendDev &pend, const CUtensorMap &mrows,
               const CUtensorMap &mq, int q0, float eps, cudaStream_t s) {
    constexpr int STAGE = A_BYTES + NQ * CHUNK_BYTES;
    const size_t ctrl = sizeof(F32Shared<NQ>);
    int stages = (int)((227 * 1024 - 1024 - ctrl) / STAGE);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const size_t smem = 1024 + (size_t)stages * STAGE + ctrl;
    auto kernel = scan_f32_tc_kernel<NQ, METRIC>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + TILE_M - 1) / TILE_M;
    const unsigned grid = ntiles < (uint32_t)ix.sm_count ? ntiles : (unsigned)ix.sm_count;
    kernel<<<grid, TC_THREADS, smem, s>>>(mrows, mq, a, pend, q0, ix.dim_pad * 4 / CHUNK_BYTES, stages, eps,
                                          ix.opt.tc_prefetch_tiles);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

template <int METRIC>
int launch_metric(const Index &ix, const ScanArgs &a, const PendDev &pend, const CUtensorMap &mrows,
                  const CUtensorMap &mq128, const CUtensorMap &mq256, float eps, SearchStatus *status, cudaStream_t s,
                  int *launches) {
    for (int q0 = 0; q0 < a.nq;) {
        const int left = a.nq - q0;
        *launches += 1;
        if (left > 128) {
            PKV_TRY((launch_one<256, METRIC>(ix, a, pend, mrows, mq256, q0, eps, s)));
            q0 += 256;
        } else {
            PKV_TRY((launch_one<128, METRIC>(ix, a, pend, mrows, mq128, q0, eps, s)));
            q0 += 128;
        }
    }
    auto rk = rescore_f32_kernel<METRIC>;
    const size_t smem = (size_t)a.dim_pad * 4;
    PKV_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ry = (4 * ix.sm_count + a.nq - 1) / a.nq;  // ~4 CTAs per SM in total
    if (ry < 1) ry = 1;
    if (ry > 64) ry = 64;
    rk<<<dim3((unsigned)a.nq, (unsigned)ry), 256, smem, s>>>(a, pend, status);
    PKV_CUDA(cudaGetLastError());
    *launches += 1;
    return PKV_OK;
}

}  // namespace

int make_tmap_bytes(CUtensorMap *m, const void *base, uint64_t inner_bytes, uint64_t rows, uint64_t pitch_bytes,
                    uint32_t box_rows);

bool scan_tc_f32_supported(const Index &ix, int nq) {
    if (ix.dtype != PKV_F32 || ix.opt.force_simt) return false;
    return nq >= ix.opt.tc_min_queries_f32;
}

// TF32 keeps 10 explicit mantissa bits; whether the tensor core truncates or rounds the f32
// operands, each is off by < 2^-10 relative, a product by < 2^-9, so
// |dot~ - dot| <= 2^-9 * sum|a_i q_i| <= 2^-9 |a||q| (Cauchy-Schwarz); the f32 accumulation adds
// ~1e-6.  2.2e-3 > 2^-9 = 1.953e-3 leaves margin.  tests/test_gpu_tc_f32.py measures the real error.
static constexpr float TF32_EPS = 2.2e-3f;

FilterSpec filter_spec_tc_f32(int metric) {
    FilterSpec fs = filter_spec_simt(PKV_F32, metric);
    if (metric == PKV_COSINE) fs.abs = TF32_EPS;  // in units of dot/|a|: scaled by |q| in filter_threshold
    return fs;  // L2 / DOT apply the per-row bound inside the kernel
}

int launch_scan_tc_f32(const Index &ix, const ScanArgs &a, uint32_t *d_pend_rows, uint32_t *d_pend_cnt, uint32_t pend_cap,
                       SearchStatus *d_status, cudaStream_t s, int *launches) {
    if (a.row_end <= a.row_begin || a.nq <= 0) return PKV_OK;
    PendDev pend{d_pend_rows, d_pend_cnt, pend_cap};
    PKV_CUDA(cudaMemsetAsync(d_pend_cnt, 0, sizeof(uint32_t) * a.nq, s));
    CUtensorMap mrows, mq128, mq256;
    PKV_TRY(make_tmap_bytes(&mrows, ix.d_data, (uint64_t)ix.pitch, (uint64_t)ix.sealed_rows, (uint64_t)ix.pitch, TILE_M));
    PKV_TRY(make_tmap_bytes(&mq128, a.queries, (uint64_t)ix.pitch, (uint64_t)a.nq, (uint64_t)ix.pitch, 128));
    PKV_TRY(make_tmap_bytes(&mq256, a.queries, (uint64_t)ix.pitch, (uint64_t)a.nq, (uint64_t)ix.pitch, 256));
    switch (a.metric) {
        case PKV_COSINE:
            return launch_metric<PKV_COSINE>(ix, a, pend, mrows, mq128, mq256, TF32_EPS, d_status, s, launches);
        case PKV_L2: return launch_metric<PKV_L2>(ix, a, pend, mrows, mq128, mq256, TF32_EPS, d_status, s, launches);
        default: return launch_metric<PKV_DOT>(ix, a, pend, mrows, mq128, mq256, TF32_EPS, d_status, s, launches);
    }
}

}  // namespace pkv
