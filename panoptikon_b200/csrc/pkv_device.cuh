// pkv_device.cuh — device helpers shared by the scan kernels: exact distance keys,
// candidate push, warp transpose-reduction.
#pragma once
#include "pkv_internal.cuh"

namespace pkv {

// 128-bit streaming load: the corpus is read once per pass, keep it out of L1.
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ldg_stream_i4(const int4 *p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ---- exact keys: the same operation sequence as the sqlite-vec scalar code
// (oracle/pkv_oracle.c): every step is one IEEE round-to-nearest operation.
__device__ __forceinline__ float cosine_key(double dot, double a_mag, double b_mag) {
    double den = __dmul_rn(__dsqrt_rn(a_mag), __dsqrt_rn(b_mag));
    return __double2float_rn(__dsub_rn(1.0, __ddiv_rn(dot, den)));
}
__device__ __forceinline__ float l2_key_from_sum(float sum) {
    // (float)sqrt((double)sum) == sqrtf(sum): 53 >= 2*24+2 makes the double rounding innocuous
    return __fsqrt_rn(sum);
}

// Sequential f32 replay of the int8 scalar loops, used only when an integer sum is too
// large for an f32 accumulator to have stayed exact (>= 2^24), so that even those pairs
// reproduce the reference's rounding.
static __device__ __noinline__ float replay_i8(const int8_t *a, const int8_t *b, int dim, int metric) {
    if (metric == PKV_L2) {
        float res = 0.f;
        for (int i = 0; i < dim; i++) {
            float t = (float)((int)a[i] - (int)b[i]);
            res = __fadd_rn(res, __fmul_rn(t, t));
        }
        return __double2float_rn(__dsqrt_rn((double)res));
    }
    float dot = 0.f, am = 0.f, bm = 0.f;
    for (int i = 0; i < dim; i++) {
        int x = a[i], y = b[i];
        dot = __fadd_rn(dot, (float)(x * y));
        am = __fadd_rn(am, (float)(x * x));
        bm = __fadd_rn(bm, (float)(y * y));
    }
    if (metric == PKV_DOT) return -dot;
    return cosine_key((double)dot, (double)am, (double)bm);
}

// Exact int8 key from the integer dot product and the two integer squared norms.
__device__ __forceinline__ float i8_key(int metric, int dot, int a_mag, int b_mag, int dim, const int8_t *row,
                                        const int8_t *query) {
    const int LIM = 1 << 24;
    if (metric == PKV_L2) {
        int n = a_mag + b_mag - 2 * dot;
        if (n < LIM) return __double2float_rn(__dsqrt_rn((double)n));
        return replay_i8(row, query, dim, metric);
    }
    if (dim > 1024 && (a_mag >= LIM || b_mag >= LIM || (double)a_mag * (double)b_mag >= 281474976710656.0))
        return replay_i8(row, query, dim, metric);
    if (metric == PKV_DOT) return -(float)dot;
    return cosine_key((double)dot, (double)a_mag, (double)b_mag);
}

__device__ __forceinline__ bool topk_member(const TopkDev &t, int q, uint32_t row) {
    if (!t.bitmap) return true;
    const uint64_t *bm = t.bitmap + (size_t)q * (size_t)t.bitmap_stride;
    return (__ldg(bm + (row >> 6)) >> (row & 63)) & 1ull;
}

// Appends (dist,row) to query q's candidate buffer when it beats the current k-th best.
__device__ __forceinline__ void topk_push(const TopkDev &t, int q, uint32_t row, float dist) {
    uint64_t key = pack_key(dist, row);
    if (key >= __ldg(t.thr_key + q)) return;
    uint32_t slot = atomicAdd(t.cnt + q, 1u);
    if (slot < t.cap) t.cand[(size_t)q * t.cap + slot] = key;
}

// ---- warp transpose-reduction --------------------------------------------
// Every lane holds V partial sums; afterwards each of the V totals lives in exactly one
// place: total i (low 5 bits = bit-reversed lane) sits in v[i >> 5] of that lane.
// Costs V-1 (+ a few) shuffles instead of 5*V.
constexpr __host__ __device__ int ilog2_c(int v) { return v <= 1 ? 0 : 1 + ilog2_c(v >> 1); }

template <int V, typename T>
__device__ __forceinline__ void warp_transpose_reduce(T (&v)[V], const int lane) {
    static_assert(V >= 1 && (V & (V - 1)) == 0, "V must be a power of two");
    constexpr int LOGV = ilog2_c(V);
#pragma unroll
    for (int step = 0; step < 5; ++step) {
        const int s = 16 >> step;
        if (step < LOGV) {
            const int n = V >> step;
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < n / 2; ++i) {
                T keep = up ? v[2 * i + 1] : v[2 * i];
                T send = up ? v[2 * i] : v[2 * i + 1];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], s);
        }
    }
}
// low bits of the total index owned by `lane` after warp_transpose_reduce<V>
template <int V>
__device__ __forceinline__ int transpose_owned_low(const int lane) {
    constexpr int NB = ilog2_c(V) < 5 ? ilog2_c(V) : 5;
    return (int)(__brev((unsigned)lane) >> 27) & ((1 << NB) - 1);
}
// for V < 32 several lanes hold the same total; exactly one of them acts
template <int V>
__device__ __forceinline__ bool transpose_is_owner(const int lane) {
    constexpr int NB = ilog2_c(V) < 5 ? ilog2_c(V) : 5;
    return (lane & ((1 << (5 - NB)) - 1)) == 0;
}
// lane that owns total index r after warp_transpose_reduce<V> with V <= 32
__device__ __forceinline__ int transpose_owner_lane(const int r) { return (int)(__brev((unsigned)r) >> 27); }

}  // namespace pkv
