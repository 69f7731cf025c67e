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

// ---- live search state: per-query thresholds that other CTAs tighten while a scan runs ----------
// Reads that must see those updates go to L2 (relaxed, gpu scope) instead of the non-coherent path.
__device__ __forceinline__ uint64_t ld_live_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_live_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_live_f32(const float *p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
// *addr = min(*addr, v) for floats of either sign (never NaN): non-negative floats order like signed
// ints, negative ones like reversed unsigned ints.
__device__ __forceinline__ void atomic_min_f32(float *addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

// Exact k-th best distance d_k -> the cheap threshold a scan kernel compares its own per-pair figure
// against (FilterKind, pkv_internal.cuh).  Monotone in d_k; rounds towards "keep".
__device__ __forceinline__ float filter_threshold(FilterSpec fs, float d_k, float b_mag) {
    const float INF = __int_as_float(0x7f800000);
    if (d_k != d_k) return INF;
    if (fs.kind == FK_EXACT_DIST) return d_k;
    if (fs.kind == FK_COS_RATIO) {
        // in top-k only if dot/(sqrt(a)sqrt(b)) >= 1 - d_k - delta, delta covering the f32 rounding of d
        double sb = sqrt((double)b_mag);
        double T = (1.0 - (double)d_k - 2.4e-7) * sb;
        double tf = -T + fabs(T) * (double)fs.rel + (double)fs.abs * sb + 1e-30;
        float f = (float)tf;
        if ((double)f < tf) f = nextafterf(f, INF);
        return f;
    }
    double tf = (double)d_k * (double)d_k * (1.0 + (double)fs.rel) + (double)fs.abs;
    float f = (float)tf;
    if ((double)f < tf) f = nextafterf(f, INF);
    return f;
}

// Appends (dist,row) to query q's candidate buffer when it beats the current k-th best.
__device__ __forceinline__ void topk_push(const TopkDev &t, int q, uint32_t row, float dist) {
    uint64_t key = pack_key(dist, row);
    if (key >= __ldg(t.thr_key + q)) return;
    uint32_t slot = atomicAdd(t.cnt + q, 1u);
    if (slot < t.cap) t.cand[(size_t)q * t.cap + slot] = key;
}

// The same against the LIVE threshold.  Returns true when this push is the query's refresh trigger:
// the caller's warp then re-selects the threshold (live_refresh).
__device__ __forceinline__ bool topk_push_live(const TopkDev &t, int q, uint32_t row, float dist) {
    const uint64_t key = pack_key(dist, row);
    if (key >= ld_live_u64(t.thr_key + q)) return false;
    const uint32_t slot = atomicAdd(t.cnt + q, 1u);
    if (slot >= t.cap) return false;  // overflow: the final select sees cnt > cap and the search is redone chunked
    t.cand[(size_t)q * t.cap + slot] = key;
    return ((slot + 1u) % t.refresh_every) == 0u;
}

// In-kernel threshold maintenance (whole warp, converged).  The candidate buffer of query q is
// append-only during a live launch: slots [0, cnt) hold keys of real (row, distance) pairs that beat the
// threshold of their time, or KEY_MAX where the push is still in flight.  The k-th smallest key of ANY
// set of >= k distinct real pairs bounds the final k-th best from above, so the result may be published
// with atomicMin whatever other warps do meanwhile.  Only keys <= the current threshold can be among the
// k smallest; they are gathered into registers (<= R per lane, else the refresh is skipped: thresholds
// only ever tighten, a skipped refresh costs candidates, never correctness) and the k-th is found by
// bisection on the distance bits, then on the row bits among the ties.
template <int R>
__device__ __noinline__ void live_refresh(const TopkDev &t, int q, int lane) {
    // one lane reads the moving state: every loop below must run the same number of times in all lanes
    uint32_t raw = 0;
    uint64_t told = 0;
    if (lane == 0) {
        raw = ld_live_u32(t.cnt + q);
        told = ld_live_u64(t.thr_key + q);
    }
    raw = __shfl_sync(0xffffffffu, raw, 0);
    told = __shfl_sync(0xffffffffu, told, 0);
    const uint32_t n = raw < t.cap ? raw : t.cap;
    const uint64_t *mine = t.cand + (size_t)q * t.cap;
    uint32_t hi_w[R], lo_w[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        hi_w[j] = 0xFFFFFFFFu;
        lo_w[j] = 0xFFFFFFFFu;
    }
    int c = 0;
    for (uint32_t base = 0; base < n; base += 256) {
        uint64_t kk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // eight independent L2 reads in flight per lane
            const uint32_t i = base + u * 32 + lane;
            kk[u] = i < n ? ld_live_u64(mine + i) : KEY_MAX;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (kk[u] <= told && kk[u] != KEY_MAX) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    if (j == c) {
                        hi_w[j] = (uint32_t)(kk[u] >> 32);
                        lo_w[j] = (uint32_t)kk[u];
                    }
                }
                ++c;
            }
        }
    }
    if (__any_sync(0xffffffffu, c > R)) {
        if (lane == 0) atomicAdd(&t.status->live_skips, 1u);
        return;
    }
    const int total = __reduce_add_sync(0xffffffffu, c);
    if (total < t.k) return;
    // k-th smallest distance word: smallest v with count(hi <= v) >= k.  The live keys span a narrow band of
    // distances, so the bisection starts from their own range instead of [0, threshold]
    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < R; ++j) mn = (j < c && hi_w[j] < mn) ? hi_w[j] : mn;
    uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = (uint32_t)(told >> 32);
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        int cl = 0;
#pragma unroll
        for (int j = 0; j < R; ++j) cl += (j < c && hi_w[j] <= mid) ? 1 : 0;
        if (__reduce_add_sync(0xffffffffu, cl) >= t.k) hi = mid;
        else lo = mid + 1u;
    }
    const uint32_t dk = lo;
    int less = 0, ties = 0;
    uint32_t tie_row = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < R; ++j) {
        less += (j < c && hi_w[j] < dk) ? 1 : 0;
        if (j < c && hi_w[j] == dk) {
            ++ties;
            tie_row = lo_w[j] < tie_row ? lo_w[j] : tie_row;
        }
    }
    const int need = t.k - __reduce_add_sync(0xffffffffu, less);  // >= 1: rank of the k-th among the keys at dk
    ties = __reduce_add_sync(0xffffffffu, ties);
    uint32_t rlo = __reduce_min_sync(0xffffffffu, tie_row), rhi = 0xFFFFFFFFu;
    if (ties > 1 && need > 1) {  // several rows share the k-th distance: bisect on the row word among them
        while (rlo < rhi) {
            const uint32_t mid = rlo + ((rhi - rlo) >> 1);
            int cl = 0;
#pragma unroll
            for (int j = 0; j < R; ++j) cl += (j < c && hi_w[j] == dk && lo_w[j] <= mid) ? 1 : 0;
            if (__reduce_add_sync(0xffffffffu, cl) >= need) rhi = mid;
            else rlo = mid + 1u;
        }
    }
    const uint64_t kth = ((uint64_t)dk << 32) | (uint64_t)rlo;
    if (lane == 0) atomicAdd(&t.status->live_refreshes, 1u);
    if (lane == 0 && kth < told) {
        atomicMin(reinterpret_cast<unsigned long long *>(t.thr_key + q), (unsigned long long)kth);
        atomic_min_f32(t.thr_f + q, filter_threshold(t.fs, unordered_bits(dk), __ldg(t.q_mag_f + q)));
    }
}

// Runs live_refresh for every lane whose last push was a trigger (q >= 0), one query at a time.
template <int R>
__device__ __forceinline__ void live_refresh_pending(const TopkDev &t, int trig_q, int lane) {
    unsigned need = __ballot_sync(0xffffffffu, trig_q >= 0);
    while (need) {
        const int src = __ffs(need) - 1;
        const int qq = __shfl_sync(0xffffffffu, trig_q, src);
        live_refresh<R>(t, qq, lane);
        need &= need - 1u;
    }
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
