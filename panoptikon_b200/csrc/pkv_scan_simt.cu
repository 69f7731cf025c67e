// pkv_scan_simt.cu — CUDA-core scan kernels (FFMA for f32/f16 rows, DP4A for int8 rows).
//
// These are the small-batch path (HBM-bound up to ~8-16 queries per pass on f32) and the
// always-correct path every other kernel is checked against.  Replaces, for a whole batch of
// queries at once, the per-row scalar calls vec_distance_cosine / vec_distance_L2 that the
// reference's SQL makes (pql/builder/filters/image_embeddings.rs:321-362,
// text_embeddings.rs:386-418) plus the ORDER BY ... LIMIT of builder.rs:551-582.
//
// Mapping: one warp owns R consecutive rows; the 32 lanes split the components of a row with
// coalesced 128-bit loads (512 B per warp request); QB queries sit in shared memory; every
// lane keeps R*QB partial dot products (+R partial row norms) which a transpose-reduction
// turns into one finished (row, query) pair per lane; that lane applies the per-query
// threshold and, for the rare survivor, computes the exact key and pushes the candidate.
//
// Algorithmic bytes per (row, pass): dim_pad * elem_size (+4 for the int8 row norm).
#include "pkv_device.cuh"

namespace pkv {

namespace {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ITERS = 4;  // row groups per warp per work item

template <int R, int QB>
struct Tile {
    static constexpr int V = R * QB;
    static constexpr int SLOTS = V >= 32 ? V / 32 : 1;
    static constexpr int ROWS_PER_ITEM = WARPS * R * ITERS;
};

__device__ __forceinline__ float dot4(const float4 &a, const float4 &b, float acc) {
    acc = fmaf(a.x, b.x, acc);
    acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc);
    acc = fmaf(a.w, b.w, acc);
    return acc;
}
__device__ __forceinline__ float sqdiff4(const float4 &a, const float4 &b, float acc) {
    float t;
    t = a.x - b.x; acc = fmaf(t, t, acc);
    t = a.y - b.y; acc = fmaf(t, t, acc);
    t = a.z - b.z; acc = fmaf(t, t, acc);
    t = a.w - b.w; acc = fmaf(t, t, acc);
    return acc;
}

// ------------------------------------------------------------------ f32 rows
template <int METRIC, int R, int QB>
__global__ void __launch_bounds__(THREADS) scan_f32_simt_kernel(const ScanArgs a) {
    using T = Tile<R, QB>;
    extern __shared__ float4 s_q[];  // [QB][dim_pad/4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = a.dim_pad >> 2;
    const int ngroups = (a.nq + QB - 1) / QB;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t nblocks = (nrows + T::ROWS_PER_ITEM - 1) / T::ROWS_PER_ITEM;
    const uint64_t items = (uint64_t)nblocks * ngroups;
    const uint8_t *base = (const uint8_t *)a.data;
    int cur_group = -1;

    for (uint64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int g = (int)(item % ngroups);
        const uint32_t rb = (uint32_t)(item / ngroups);
        if (g != cur_group) {
            __syncthreads();
            const float4 *gq = (const float4 *)a.queries;
            for (int i = threadIdx.x; i < QB * nvec; i += THREADS) {
                int b = i / nvec, j = i - b * nvec;
                int q = g * QB + b;
                s_q[i] = q < a.nq ? gq[(size_t)q * nvec + j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            cur_group = g;
        }
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t row0 = a.row_begin + rb * T::ROWS_PER_ITEM + (uint32_t)(it * WARPS + warp) * R;
            if (row0 >= a.row_end) break;
            float acc[T::V];
            float nrm[R];
#pragma unroll
            for (int i = 0; i < T::V; ++i) acc[i] = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) nrm[r] = 0.f;
            const float4 *rp[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t row = row0 + r < a.row_end ? row0 + r : a.row_end - 1;
                rp[r] = (const float4 *)(base + (size_t)row * (size_t)a.pitch_bytes);
            }
#pragma unroll 2
            for (int j = lane; j < nvec; j += 32) {
                float4 av[R];
#pragma unroll
                for (int r = 0; r < R; ++r) av[r] = ldg_stream_f4(rp[r] + j);
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const float4 q = s_q[b * nvec + j];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (METRIC == PKV_L2)
                            acc[r * QB + b] = sqdiff4(av[r], q, acc[r * QB + b]);
                        else
                            acc[r * QB + b] = dot4(av[r], q, acc[r * QB + b]);
                    }
                }
                if (METRIC == PKV_COSINE) {
#pragma unroll
                    for (int r = 0; r < R; ++r) nrm[r] = dot4(av[r], av[r], nrm[r]);
                }
            }
            warp_transpose_reduce<T::V>(acc, lane);
            if (METRIC == PKV_COSINE) warp_transpose_reduce<R>(nrm, lane);
            const int low = transpose_owned_low<T::V>(lane);
            const bool owner = transpose_is_owner<T::V>(lane);
#pragma unroll
            for (int t = 0; t < T::SLOTS; ++t) {
                const int i = t * 32 + low;
                const int r = i / QB, b = i - r * QB;
                float am = 0.f;
                if (METRIC == PKV_COSINE) am = __shfl_sync(0xffffffffu, nrm[0], transpose_owner_lane(r));
                const uint32_t row = row0 + r;
                const int q = g * QB + b;
                if (!owner || row >= a.row_end || q >= a.nq) continue;
                const float v = acc[t];
                if (a.dense_out) {  // full scoring: the reference's untruncated dist CTE
                    float d;
                    if (METRIC == PKV_COSINE)
                        d = cosine_key((double)v, (double)am, (double)__ldg(a.q_mag_f + q));
                    else if (METRIC == PKV_L2)
                        d = l2_key_from_sum(v);
                    else
                        d = 0.0f - v;
                    a.dense_out[(size_t)q * a.dense_stride + row] = d;
                    continue;
                }
                const float tf = __ldg(a.topk.thr_f + q);
                const uint32_t lrow = row;  // local row number inside this index
                if (METRIC == PKV_COSINE) {
                    const float f = -v * rsqrtf(am);
                    if (f > tf) continue;
                    if (!topk_member(a.topk, q, lrow)) continue;
                    const float d = cosine_key((double)v, (double)am, (double)__ldg(a.q_mag_f + q));
                    topk_push(a.topk, q, lrow, d);
                } else if (METRIC == PKV_L2) {
                    if (v > tf) continue;
                    if (!topk_member(a.topk, q, lrow)) continue;
                    topk_push(a.topk, q, lrow, l2_key_from_sum(v));
                } else {
                    const float d = -v;
                    if (d > tf) continue;
                    if (!topk_member(a.topk, q, lrow)) continue;
                    topk_push(a.topk, q, lrow, d);
                }
            }
        }
    }
}

// ----------------------------------------------------------------- int8 rows
__device__ __forceinline__ int dp4a4(const int4 &a, const int4 &b, int acc) {
    acc = __dp4a(a.x, b.x, acc);
    acc = __dp4a(a.y, b.y, acc);
    acc = __dp4a(a.z, b.z, acc);
    acc = __dp4a(a.w, b.w, acc);
    return acc;
}

template <int R, int QB>
__global__ void __launch_bounds__(THREADS) scan_i8_simt_kernel(const ScanArgs a) {
    using T = Tile<R, QB>;
    extern __shared__ float4 s_q[];
    int4 *s_qi = reinterpret_cast<int4 *>(s_q);  // [QB][dim_pad/16]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec = a.dim_pad >> 4;
    const int ngroups = (a.nq + QB - 1) / QB;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t nblocks = (nrows + T::ROWS_PER_ITEM - 1) / T::ROWS_PER_ITEM;
    const uint64_t items = (uint64_t)nblocks * ngroups;
    const uint8_t *base = (const uint8_t *)a.data;
    int cur_group = -1;

    for (uint64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int g = (int)(item % ngroups);
        const uint32_t rb = (uint32_t)(item / ngroups);
        if (g != cur_group) {
            __syncthreads();
            const int4 *gq = (const int4 *)a.queries;
            for (int i = threadIdx.x; i < QB * nvec; i += THREADS) {
                int b = i / nvec, j = i - b * nvec;
                int q = g * QB + b;
                s_qi[i] = q < a.nq ? gq[(size_t)q * nvec + j] : make_int4(0, 0, 0, 0);
            }
            __syncthreads();
            cur_group = g;
        }
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t row0 = a.row_begin + rb * T::ROWS_PER_ITEM + (uint32_t)(it * WARPS + warp) * R;
            if (row0 >= a.row_end) break;
            int acc[T::V];
#pragma unroll
            for (int i = 0; i < T::V; ++i) acc[i] = 0;
            const int4 *rp[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t row = row0 + r < a.row_end ? row0 + r : a.row_end - 1;
                rp[r] = (const int4 *)(base + (size_t)row * (size_t)a.pitch_bytes);
            }
            for (int j = lane; j < nvec; j += 32) {
                int4 av[R];
#pragma unroll
                for (int r = 0; r < R; ++r) av[r] = ldg_stream_i4(rp[r] + j);
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const int4 q = s_qi[b * nvec + j];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r * QB + b] = dp4a4(av[r], q, acc[r * QB + b]);
                }
            }
            warp_transpose_reduce<T::V>(acc, lane);
            const int low = transpose_owned_low<T::V>(lane);
            const bool owner = transpose_is_owner<T::V>(lane);
#pragma unroll
            for (int t = 0; t < T::SLOTS; ++t) {
                const int i = t * 32 + low;
                const int r = i / QB, b = i - r * QB;
                const uint32_t row = row0 + r;
                const int q = g * QB + b;
                if (!owner || row >= a.row_end || q >= a.nq) continue;
                const int dot = acc[t];
                const int am = __ldg(a.row_mag_i + row);
                const int bm = __ldg(a.q_mag_i + q);
                if (a.dense_out) {
                    const int8_t *rowp = (const int8_t *)(base + (size_t)row * (size_t)a.pitch_bytes);
                    const int8_t *qp = (const int8_t *)a.queries + (size_t)q * a.dim_pad;
                    a.dense_out[(size_t)q * a.dense_stride + row] = i8_key(a.metric, dot, am, bm, a.dim, rowp, qp) + 0.0f;
                    continue;
                }
                const float tf = __ldg(a.topk.thr_f + q);
                float f;
                if (a.metric == PKV_COSINE)
                    f = -(float)dot * rsqrtf((float)am);
                else if (a.metric == PKV_L2)
                    f = (float)(am + bm - 2 * dot);
                else
                    f = -(float)dot;
                if (f > tf) continue;
                if (!topk_member(a.topk, q, row)) continue;
                const int8_t *rowp = (const int8_t *)(base + (size_t)row * (size_t)a.pitch_bytes);
                const int8_t *qp = (const int8_t *)a.queries + (size_t)q * a.dim_pad;
                topk_push(a.topk, q, row, i8_key(a.metric, dot, am, bm, a.dim, rowp, qp));
            }
        }
    }
}

// ------------------------------------------------------------------ f16 rows
// Framework extension (BASELINE config 4): rows and queries are IEEE half; arithmetic is
// the f32 formulas on the widened values (oracle: ORC_F16).  Queries are widened to f32 by
// the prep kernel, so shared memory holds f32 queries.
__device__ __forceinline__ void half8_to_float(const uint4 &u, float (&f)[8]) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 p = __half22float2(h[i]);
        f[2 * i] = p.x;
        f[2 * i + 1] = p.y;
    }
}

template <int METRIC, int R, int QB>
__global__ void __launch_bounds__(THREADS) scan_f16_simt_kernel(const ScanArgs a) {
    using T = Tile<R, QB>;
    extern __shared__ float4 s_q[];  // [QB][dim_pad/4] f32
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvec8 = a.dim_pad >> 3;  // 8 halfs per 16-byte load
    const int nvec4 = a.dim_pad >> 2;
    const int ngroups = (a.nq + QB - 1) / QB;
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t nblocks = (nrows + T::ROWS_PER_ITEM - 1) / T::ROWS_PER_ITEM;
    const uint64_t items = (uint64_t)nblocks * ngroups;
    const uint8_t *base = (const uint8_t *)a.data;
    int cur_group = -1;

    for (uint64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int g = (int)(item % ngroups);
        const uint32_t rb = (uint32_t)(item / ngroups);
        if (g != cur_group) {
            __syncthreads();
            const float4 *gq = (const float4 *)a.queries;
            for (int i = threadIdx.x; i < QB * nvec4; i += THREADS) {
                int b = i / nvec4, j = i - b * nvec4;
                int q = g * QB + b;
                s_q[i] = q < a.nq ? gq[(size_t)q * nvec4 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
            cur_group = g;
        }
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t row0 = a.row_begin + rb * T::ROWS_PER_ITEM + (uint32_t)(it * WARPS + warp) * R;
            if (row0 >= a.row_end) break;
            float acc[T::V];
            float nrm[R];
#pragma unroll
            for (int i = 0; i < T::V; ++i) acc[i] = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) nrm[r] = 0.f;
            const uint4 *rp[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                uint32_t row = row0 + r < a.row_end ? row0 + r : a.row_end - 1;
                rp[r] = (const uint4 *)(base + (size_t)row * (size_t)a.pitch_bytes);
            }
            for (int j = lane; j < nvec8; j += 32) {
                float av[R][8];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    int4 raw = ldg_stream_i4((const int4 *)(rp[r] + j));
                    half8_to_float(*reinterpret_cast<uint4 *>(&raw), av[r]);
                }
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const float4 q0 = s_q[b * nvec4 + 2 * j], q1 = s_q[b * nvec4 + 2 * j + 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float4 a0 = make_float4(av[r][0], av[r][1], av[r][2], av[r][3]);
                        const float4 a1 = make_float4(av[r][4], av[r][5], av[r][6], av[r][7]);
                        if (METRIC == PKV_L2) {
                            acc[r * QB + b] = sqdiff4(a0, q0, acc[r * QB + b]);
                            acc[r * QB + b] = sqdiff4(a1, q1, acc[r * QB + b]);
                        } else {
                            acc[r * QB + b] = dot4(a0, q0, acc[r * QB + b]);
                            acc[r * QB + b] = dot4(a1, q1, acc[r * QB + b]);
                        }
                    }
                }
                if (METRIC == PKV_COSINE) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) nrm[r] = fmaf(av[r][e], av[r][e], nrm[r]);
                    }
                }
            }
            warp_transpose_reduce<T::V>(acc, lane);
            if (METRIC == PKV_COSINE) warp_transpose_reduce<R>(nrm, lane);
            const int low = transpose_owned_low<T::V>(lane);
            const bool owner = transpose_is_owner<T::V>(lane);
#pragma unroll
            for (int t = 0; t < T::SLOTS; ++t) {
                const int i = t * 32 + low;
                const int r = i / QB, b = i - r * QB;
                float am = 0.f;
                if (METRIC == PKV_COSINE) am = __shfl_sync(0xffffffffu, nrm[0], transpose_owner_lane(r));
                const uint32_t row = row0 + r;
                const int q = g * QB + b;
                if (!owner || row >= a.row_end || q >= a.nq) continue;
                const float v = acc[t];
                if (a.dense_out) {
                    float d;
                    if (METRIC == PKV_COSINE)
                        d = cosine_key((double)v, (double)am, (double)__ldg(a.q_mag_f + q));
                    else if (METRIC == PKV_L2)
                        d = l2_key_from_sum(v);
                    else
                        d = 0.0f - v;
                    a.dense_out[(size_t)q * a.dense_stride + row] = d;
                    continue;
                }
                const float tf = __ldg(a.topk.thr_f + q);
                if (METRIC == PKV_COSINE) {
                    const float f = -v * rsqrtf(am);
                    if (f > tf) continue;
                    if (!topk_member(a.topk, q, row)) continue;
                    topk_push(a.topk, q, row, cosine_key((double)v, (double)am, (double)__ldg(a.q_mag_f + q)));
                } else if (METRIC == PKV_L2) {
                    if (v > tf) continue;
                    if (!topk_member(a.topk, q, row)) continue;
                    topk_push(a.topk, q, row, l2_key_from_sum(v));
                } else {
                    const float d = -v;
                    if (d > tf) continue;
                    if (!topk_member(a.topk, q, row)) continue;
                    topk_push(a.topk, q, row, d);
                }
            }
        }
    }
}

template <typename K>
int launch_one(K kernel, const Index &ix, const ScanArgs &a, int rows_per_item, int qb, size_t smem,
               cudaStream_t s) {
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    PKV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, THREADS, smem));
    if (occ < 1) return fail(PKV_ERR_UNSUPPORTED, "scan kernel does not fit on an SM (smem %zu)", smem);
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint64_t items = (uint64_t)((nrows + rows_per_item - 1) / rows_per_item) * ((a.nq + qb - 1) / qb);
    uint64_t grid = (uint64_t)ix.sm_count * occ;  // persistent: one wave of resident CTAs
    if (grid > items) grid = items;
    if (grid < 1) grid = 1;
    kernel<<<(unsigned)grid, THREADS, smem, s>>>(a);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

template <int R, int QB>
int launch_f32(const Index &ix, const ScanArgs &a, cudaStream_t s) {
    const size_t smem = (size_t)QB * a.dim_pad * 4;
    const int rpi = Tile<R, QB>::ROWS_PER_ITEM;
    switch (a.metric) {
        case PKV_L2: return launch_one(scan_f32_simt_kernel<PKV_L2, R, QB>, ix, a, rpi, QB, smem, s);
        case PKV_COSINE: return launch_one(scan_f32_simt_kernel<PKV_COSINE, R, QB>, ix, a, rpi, QB, smem, s);
        default: return launch_one(scan_f32_simt_kernel<PKV_DOT, R, QB>, ix, a, rpi, QB, smem, s);
    }
}
template <int R, int QB>
int launch_f16(const Index &ix, const ScanArgs &a, cudaStream_t s) {
    const size_t smem = (size_t)QB * a.dim_pad * 4;
    const int rpi = Tile<R, QB>::ROWS_PER_ITEM;
    switch (a.metric) {
        case PKV_L2: return launch_one(scan_f16_simt_kernel<PKV_L2, R, QB>, ix, a, rpi, QB, smem, s);
        case PKV_COSINE: return launch_one(scan_f16_simt_kernel<PKV_COSINE, R, QB>, ix, a, rpi, QB, smem, s);
        default: return launch_one(scan_f16_simt_kernel<PKV_DOT, R, QB>, ix, a, rpi, QB, smem, s);
    }
}
template <int R, int QB>
int launch_i8(const Index &ix, const ScanArgs &a, cudaStream_t s) {
    const size_t smem = (size_t)QB * a.dim_pad;
    return launch_one(scan_i8_simt_kernel<R, QB>, ix, a, Tile<R, QB>::ROWS_PER_ITEM, QB, smem, s);
}

}  // namespace

int launch_scan_simt(const Index &ix, const ScanArgs &a, cudaStream_t s, int *launches) {
    if (a.row_end <= a.row_begin || a.nq <= 0) return PKV_OK;
    *launches += 1;
    if (ix.dtype == PKV_F32) {
        if (a.nq == 1) return launch_f32<8, 1>(ix, a, s);
        if (a.nq <= 4) return launch_f32<8, 4>(ix, a, s);
        return launch_f32<4, 8>(ix, a, s);
    }
    if (ix.dtype == PKV_F16) {
        if (a.nq == 1) return launch_f16<4, 1>(ix, a, s);
        if (a.nq <= 4) return launch_f16<4, 4>(ix, a, s);
        return launch_f16<4, 8>(ix, a, s);
    }
    if (a.nq == 1) return launch_i8<8, 1>(ix, a, s);
    if (a.nq <= 4) return launch_i8<8, 4>(ix, a, s);
    return launch_i8<4, 8>(ix, a, s);
}

FilterSpec filter_spec_simt(int dtype, int metric) {
    FilterSpec fs;
    if (metric == PKV_DOT) {
        fs.kind = FK_EXACT_DIST;
        fs.rel = 0.f;
        fs.abs = 0.f;
    } else if (metric == PKV_COSINE) {
        fs.kind = FK_COS_RATIO;
        fs.rel = 1e-5f;  // rsqrtf + one multiply are good to a few ulp; 1e-5 is ample
        fs.abs = 0.f;
    } else {
        fs.kind = FK_L2_SQUARED;
        fs.rel = dtype == PKV_I8 ? 1e-4f : 1e-6f;
        fs.abs = dtype == PKV_I8 ? 1.0f : 1e-30f;
    }
    return fs;
}

}  // namespace pkv
