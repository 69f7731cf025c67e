// pkv_topk.cu — everything around the scan: query preparation, the per-query candidate
// select (sort + dedupe + threshold update), result finalisation, cross-shard merge,
// per-item aggregation, and the int8 codec kernels.
#include "pkv_device.cuh"

namespace pkv {

namespace {

// ------------------------------------------------------------ query prep
// One CTA per query: lays the query out in the scan dtype with the index row pitch and
// computes its squared norm the way the reference accumulates bMag (sequential f32 for
// f32/f16 data; exact integer for int8 codes).
__global__ void prep_queries_kernel(const void *qraw, int query_dtype, int index_dtype, int dim, int dim_pad,
                                    float scale, void *qout, float *q_mag_f, int32_t *q_mag_i) {
    const int q = blockIdx.x;
    __shared__ int s_red[32];
    if (index_dtype == PKV_I8) {
        int8_t *out = (int8_t *)qout + (size_t)q * dim_pad;
        int local = 0;
        for (int i = threadIdx.x; i < dim_pad; i += blockDim.x) {
            int code = 0;
            if (i < dim) {
                if (query_dtype == PKV_I8) {
                    code = ((const int8_t *)qraw)[(size_t)q * dim + i];
                } else {
                    // quantize_int8 (db/vector_quants.rs:1489-1497)
                    float v = ((const float *)qraw)[(size_t)q * dim + i];
                    float r = rintf(__fdiv_rn(v, scale));
                    if (r != r) {
                        code = 0;
                    } else {
                        r = fminf(fmaxf(r, -128.0f), 127.0f);
                        code = (int)r;
                    }
                }
            }
            out[i] = (int8_t)code;
            local += code * code;
        }
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_red[w];
            q_mag_i[q] = tot;
            q_mag_f[q] = (float)tot;
        }
    } else {
        float *out = (float *)qout + (size_t)q * dim_pad;
        for (int i = threadIdx.x; i < dim_pad; i += blockDim.x) {
            float v = 0.f;
            if (i < dim) {
                if (query_dtype == PKV_F16)
                    v = __half2float(((const __half *)qraw)[(size_t)q * dim + i]);
                else
                    v = ((const float *)qraw)[(size_t)q * dim + i];
            }
            out[i] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = 0.f;
            for (int i = 0; i < dim; ++i) m = __fadd_rn(m, __fmul_rn(out[i], out[i]));
            q_mag_f[q] = m;
            q_mag_i[q] = 0;
        }
    }
}

__global__ void reset_state_kernel(uint32_t *cnt, uint64_t *thr_key, float *thr_f, SearchStatus *st, int nq,
                                   uint32_t *pend_cnt, uint32_t *defer_cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) {
        cnt[i] = 0;
        if (pend_cnt) pend_cnt[i] = 0;
        if (defer_cnt) defer_cnt[i] = 0;
        thr_key[i] = KEY_MAX;
        thr_f[i] = __int_as_float(0x7f800000);
    }
    if (i == 0) {
        st->max_raw_cnt = 0;
        st->any_overflow = 0;
        st->min_filled = 0xFFFFFFFFu;
        st->sticky_overflow = 0;
        st->live_refreshes = 0;
        st->live_skips = 0;
        st->rescored = 0;
        st->deferred = 0;
        st->defer_overflow = 0;
    }
}

__global__ void reset_status_kernel(SearchStatus *st) {
    st->max_raw_cnt = 0;
    st->any_overflow = 0;
    st->min_filled = 0xFFFFFFFFu;
}

// ----------------------------------------------------------------- select
// One CTA per query.  Sorts the candidate buffer (bitonic, shared memory), drops duplicate
// keys (a chunk re-scanned after an overflow pushes the same rows again), keeps the best k
// in cand[0..k) sorted ascending, and publishes the new exact and filter thresholds.
//   clear_tail: the next launch is a live scan (append-only buffer): every slot behind the kept keys becomes KEY_MAX.
//   out_ids != NULL: this is the last select of the search; the result rows are written here too (no finalize launch).
__global__ void __launch_bounds__(512) select_kernel(uint64_t *cand, uint32_t *cnt, uint64_t *thr_key, float *thr_f,
                                                     const float *q_mag_f, SearchStatus *status, uint32_t cap, int k,
                                                     FilterSpec fs, uint32_t *pend_cnt, uint32_t *defer_cnt, int clear_tail,
                                                     const int64_t *row_ids, int64_t row_base, int64_t *out_ids,
                                                     float *out_dist, int32_t *out_counts, int guess_rank) {
    extern __shared__ uint64_t s_keys[];
    __shared__ uint64_t s_kth;
    __shared__ int s_m;
    const int q = blockIdx.x;
    const uint32_t raw = cnt[q];
    const uint32_t n = raw < cap ? raw : cap;
    uint64_t *mine = cand + (size_t)q * cap;
    uint32_t P = 2;
    while (P < n) P <<= 1;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) s_keys[i] = i < n ? mine[i] : KEY_MAX;
    __syncthreads();
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < (P >> 1); i += blockDim.x) {
                uint32_t pos = 2 * i - (i & (stride - 1));
                uint64_t x = s_keys[pos], y = s_keys[pos + stride];
                bool up = (pos & size) == 0;
                if ((x > y) == up) {
                    s_keys[pos] = y;
                    s_keys[pos + stride] = x;
                }
            }
            __syncthreads();
        }
    }
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t out = 0;
        uint64_t kth = KEY_MAX;
        for (uint32_t base = 0; base < n && out < (uint32_t)k; base += 32) {
            uint32_t i = base + lane;
            uint64_t key = i < n ? s_keys[i] : KEY_MAX;
            bool uniq = i < n && (i == 0 || key != s_keys[i - 1]);
            unsigned mask = __ballot_sync(0xffffffffu, uniq);
            uint32_t pos = out + __popc(mask & ((1u << lane) - 1u));
            if (uniq && pos < (uint32_t)k) {
                mine[pos] = key;
                if (pos == (uint32_t)k - 1) kth = key;
            }
            out += __popc(mask);
        }
        // the lane that wrote slot k-1 knows the k-th key
        for (int o = 16; o > 0; o >>= 1) {
            uint64_t other = __shfl_xor_sync(0xffffffffu, kth, o);
            kth = other < kth ? other : kth;
        }
        if (lane == 0) {
            uint32_t m = out < (uint32_t)k ? out : (uint32_t)k;
            s_m = (int)m;
            s_kth = m == (uint32_t)k ? kth : KEY_MAX;
            // guess mode: the keys are those of a SAMPLE of the corpus; the threshold the scan starts from is the
            // guess_rank-th best of the sample (a real row's distance: about guess_rank * N / sample rows beat it) and
            // the candidate list starts empty - the sample's rows are found again by the scan itself
            if (guess_rank > 0) {
                s_kth = (uint32_t)guess_rank <= n ? s_keys[guess_rank - 1] | 0xFFFFFFFFull : KEY_MAX;
                s_m = 0;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t m = (uint32_t)s_m;
        cnt[q] = m;
        if (pend_cnt) pend_cnt[q] = 0;  // the re-scorer has consumed this chunk's survivors
        if (defer_cnt) defer_cnt[q] = 0;  // ... and the deferred pass the parked pairs
        thr_key[q] = s_kth;
        float tf = __int_as_float(0x7f800000);
        if (s_kth != KEY_MAX) tf = filter_threshold(fs, unordered_bits((uint32_t)(s_kth >> 32)), q_mag_f[q]);
        thr_f[q] = tf;
        atomicMax(&status->max_raw_cnt, raw);
        if (raw > cap) {
            atomicOr(&status->any_overflow, 1u);
            atomicOr(&status->sticky_overflow, 1u);
        }
        if (guess_rank <= 0) atomicMin(&status->min_filled, m);  // (a guess select starts the list empty on purpose)
    }
    const uint32_t m_all = (uint32_t)s_m;
    if (clear_tail)
        for (uint32_t i = m_all + threadIdx.x; i < cap; i += blockDim.x) mine[i] = KEY_MAX;
    if (out_ids) {
        for (uint32_t i = threadIdx.x; i < (uint32_t)k; i += blockDim.x) {
            int64_t id = -1;
            float d = __int_as_float(0x7fc00000);
            if (i < m_all) {
                const uint64_t key = mine[i];
                const uint32_t row = (uint32_t)key;
                id = row_ids ? row_ids[row] : row_base + (int64_t)row;
                d = unordered_bits((uint32_t)(key >> 32));
            }
            out_ids[(size_t)q * k + i] = id;
            out_dist[(size_t)q * k + i] = d;
        }
        if (threadIdx.x == 0) out_counts[q] = (int32_t)m_all;
    }
}

__global__ void finalize_kernel(const uint64_t *cand, const uint32_t *cnt, uint32_t cap, int k, const int64_t *row_ids,
                                int64_t row_base, int64_t *out_ids, float *out_dist, int32_t *out_counts) {
    const int q = blockIdx.x;
    const uint32_t m = cnt[q] < (uint32_t)k ? cnt[q] : (uint32_t)k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        int64_t id = -1;
        float d = __int_as_float(0x7fc00000);
        if ((uint32_t)i < m) {
            uint64_t key = cand[(size_t)q * cap + i];
            uint32_t row = (uint32_t)key;
            id = row_ids ? row_ids[row] : row_base + (int64_t)row;
            d = unordered_bits((uint32_t)(key >> 32));
        }
        out_ids[(size_t)q * k + i] = id;
        out_dist[(size_t)q * k + i] = d;
    }
    if (threadIdx.x == 0) out_counts[q] = (int32_t)m;
}

// ------------------------------------------------------- per-row side arrays
__global__ void row_mags_kernel(const uint8_t *data, int64_t pitch, int dim_pad, int dtype, int64_t row_begin,
                                int64_t row_end, int32_t *mag_i, float *mag_f) {
    const int lane = threadIdx.x & 31;
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_end) return;
    const uint8_t *rp = data + row * pitch;
    if (dtype == PKV_I8) {
        int acc = 0;
        const int *p = (const int *)rp;
        for (int j = lane; j < (dim_pad >> 2); j += 32) {
            int v = p[j];
            acc = __dp4a(v, v, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) mag_i[row] = acc;
    } else if (dtype == PKV_F32) {
        float acc = 0.f;
        const float4 *p = (const float4 *)rp;
        for (int j = lane; j < (dim_pad >> 2); j += 32) {
            float4 v = p[j];
            acc = fmaf(v.x, v.x, acc);
            acc = fmaf(v.y, v.y, acc);
            acc = fmaf(v.z, v.z, acc);
            acc = fmaf(v.w, v.w, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) mag_f[row] = acc;
    } else {
        float acc = 0.f;
        const __half2 *p = (const __half2 *)rp;
        for (int j = lane; j < (dim_pad >> 1); j += 32) {
            float2 v = __half22float2(p[j]);
            acc = fmaf(v.x, v.x, acc);
            acc = fmaf(v.y, v.y, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) mag_f[row] = acc;
    }
}

__global__ void fill_ids_kernel(int64_t *ids, int64_t begin, int64_t end, int64_t base) {
    int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < end) ids[i] = base + i;
}

// ------------------------------------------------------------ shard merge
// Every (part, i) entry finds its rank in the union by binary searches over the other
// parts' (already sorted) lists.  Order: distance (NaN last), then part, then position —
// which is ascending global row for contiguous row shards gathered in rank order.
__device__ __forceinline__ int valid_len(const int64_t *ids, int k) {
    int lo = 0, hi = k;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ids[mid] == -1) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__global__ void merge_kernel(const int64_t *ids, const float *dist, int parts, int nq, int k, int64_t *o_ids,
                             float *o_dist, int32_t *o_counts) {
    const int q = blockIdx.x;
    extern __shared__ int s_len[];  // [parts]
    if ((int)threadIdx.x < parts) s_len[threadIdx.x] = valid_len(ids + ((size_t)threadIdx.x * nq + q) * k, k);
    __syncthreads();
    int total = 0;
    for (int p = 0; p < parts; ++p) total += s_len[p];
    const int m = total < k ? total : k;
    for (int e = threadIdx.x; e < parts * k; e += blockDim.x) {
        const int p = e / k, i = e - p * k;
        if (i >= s_len[p]) continue;
        const size_t off = ((size_t)p * nq + q) * k;
        const uint32_t mine = ordered_bits(dist[off + i]);
        int rank = i;
        for (int o = 0; o < parts; ++o) {
            if (o == p) continue;
            const float *od = dist + ((size_t)o * nq + q) * k;
            int lo = 0, hi = s_len[o];
            while (lo < hi) {  // o < p: count entries <= mine; o > p: count entries < mine
                int mid = (lo + hi) >> 1;
                uint32_t v = ordered_bits(od[mid]);
                bool before = o < p ? v <= mine : v < mine;
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            o_ids[(size_t)q * k + rank] = ids[off + i];
            o_dist[(size_t)q * k + rank] = dist[off + i];
        }
    }
    for (int i = m + threadIdx.x; i < k; i += blockDim.x) {
        o_ids[(size_t)q * k + i] = -1;
        o_dist[(size_t)q * k + i] = __int_as_float(0x7fc00000);
    }
    if (threadIdx.x == 0) o_counts[q] = m;
}

// Packed shard results: 12 bytes per entry {id low word, id high word, distance bits}, the payload of the ONE
// all-gather; the merge reads the gathered buffer directly (no unpacking pass).
__global__ void pack_topk_kernel(const int64_t *ids, const float *dist, int64_t n, uint32_t *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t id = (uint64_t)ids[i];
    out[3 * i] = (uint32_t)id;
    out[3 * i + 1] = (uint32_t)(id >> 32);
    out[3 * i + 2] = __float_as_uint(dist[i]);
}
__device__ __forceinline__ int64_t packed_id(const uint32_t *e) { return (int64_t)((uint64_t)e[0] | ((uint64_t)e[1] << 32)); }
__device__ int packed_valid_len(const uint32_t *list, int k) {
    int lo = 0, hi = k;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (packed_id(list + 3 * (size_t)mid) == -1) hi = mid; else lo = mid + 1;
    }
    return lo;
}
__global__ void merge_packed_kernel(const uint32_t *packed, int parts, int nq, int k, int64_t *o_ids, float *o_dist,
                                    int32_t *o_counts) {
    const int q = blockIdx.x;
    extern __shared__ int s_len[];  // [parts]
    if ((int)threadIdx.x < parts) s_len[threadIdx.x] = packed_valid_len(packed + 3 * (((size_t)threadIdx.x * nq + q) * k), k);
    __syncthreads();
    int total = 0;
    for (int p = 0; p < parts; ++p) total += s_len[p];
    const int m = total < k ? total : k;
    for (int e = threadIdx.x; e < parts * k; e += blockDim.x) {
        const int p = e / k, i = e - p * k;
        if (i >= s_len[p]) continue;
        const uint32_t *mine_e = packed + 3 * (((size_t)p * nq + q) * k + i);
        const uint32_t mine = ordered_bits(__uint_as_float(mine_e[2]));
        int rank = i;
        for (int o = 0; o < parts; ++o) {
            if (o == p) continue;
            const uint32_t *ol = packed + 3 * (((size_t)o * nq + q) * k);
            int lo = 0, hi = s_len[o];
            while (lo < hi) {  // o < p: count entries <= mine; o > p: count entries < mine
                int mid = (lo + hi) >> 1;
                uint32_t v = ordered_bits(__uint_as_float(ol[3 * (size_t)mid + 2]));
                bool before = o < p ? v <= mine : v < mine;
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            o_ids[(size_t)q * k + rank] = packed_id(mine_e);
            o_dist[(size_t)q * k + rank] = __uint_as_float(mine_e[2]);
        }
    }
    for (int i = m + threadIdx.x; i < k; i += blockDim.x) {
        o_ids[(size_t)q * k + i] = -1;
        o_dist[(size_t)q * k + i] = __int_as_float(0x7fc00000);
    }
    if (threadIdx.x == 0) o_counts[q] = m;
}

// -------------------------------------------------------------- aggregation
__device__ __forceinline__ void atomic_min_f64(double *addr, double v) {
    unsigned long long *a = (unsigned long long *)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) <= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__device__ __forceinline__ void atomic_max_f64(double *addr, double v) {
    unsigned long long *a = (unsigned long long *)addr;
    unsigned long long old = *a, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}
__global__ void agg_init_kernel(double *out, double *den, unsigned long long *cnt, int64_t n_items, int agg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    out[i] = agg == PKV_AGG_MIN ? INF : (agg == PKV_AGG_MAX ? -INF : 0.0);
    den[i] = 0.0;
    cnt[i] = 0ull;
}
__global__ void agg_accum_kernel(const float *dist, const int64_t *item, const float *w, int64_t n, int64_t n_items,
                                 int agg, double *out, double *den, unsigned long long *cnt) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int64_t it = item[r];
    if (it < 0 || it >= n_items) return;
    float df = dist[r];
    if (df != df) return;  // SQL NULL
    double d = (double)df;
    if (w) {
        double ww = (double)w[r];
        atomicAdd(out + it, d * ww);
        atomicAdd(den + it, ww);
    } else if (agg == PKV_AGG_AVG) {
        atomicAdd(out + it, d);
    } else if (agg == PKV_AGG_MIN) {
        atomic_min_f64(out + it, d);
    } else {
        atomic_max_f64(out + it, d);
    }
    atomicAdd(cnt + it, 1ull);
}
__global__ void agg_finish_kernel(double *out, const double *den, const unsigned long long *cnt, int64_t n_items,
                                  int agg, int weighted) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    if (cnt[i] == 0ull)
        out[i] = __longlong_as_double(0x7ff8000000000000ll);
    else if (weighted)
        out[i] = out[i] / den[i];
    else if (agg == PKV_AGG_AVG)
        out[i] = out[i] / (double)cnt[i];
}

// -------------------------------------------------------------------- codec
__global__ void absmax_kernel(const float *v, int64_t n, unsigned int *out_bits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = fabsf(v[i]);
        if (x > m) m = x;  // NaN never replaces the running max (vector_quants.rs:1478)
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order as uints
}
__global__ void quantize_kernel(const float *v, int64_t n, float scale, int8_t *codes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float r = rintf(__fdiv_rn(v[i], scale));
        int code;
        if (r != r) {
            code = 0;
        } else {
            r = fminf(fmaxf(r, -128.0f), 127.0f);
            code = (int)r;
        }
        codes[i] = (int8_t)code;
    }
}

}  // namespace

// ================================================================ launchers
int launch_prep_queries(const Index &ix, Workspace &ws, const void *d_qraw, int nq, int query_dtype, cudaStream_t s) {
    if (nq <= 0) return PKV_OK;
    prep_queries_kernel<<<nq, 256, 0, s>>>(d_qraw, query_dtype, ix.dtype, ix.dim, ix.dim_pad, ix.scale, ws.d_q,
                                           ws.d_q_mag_f, ws.d_q_mag_i);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_reset_state(Workspace &ws, int nq, cudaStream_t s) {
    reset_state_kernel<<<(nq + 255) / 256 > 0 ? (nq + 255) / 256 : 1, 256, 0, s>>>(ws.d_cnt, ws.d_thr_key, ws.d_thr_f,
                                                                                  ws.d_status, nq, ws.d_pend_cnt, ws.d_defer_cnt);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_reset_status(Workspace &ws, cudaStream_t s) {
    reset_status_kernel<<<1, 1, 0, s>>>(ws.d_status);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_select(const Index &ix, Workspace &ws, int nq, int k, FilterSpec fs, bool clear_tail, bool clear_deferred,
                  int64_t *d_ids, float *d_dist, int32_t *d_counts, cudaStream_t s, int guess_rank) {
    if (nq <= 0) return PKV_OK;
    const size_t smem = (size_t)ws.cap * sizeof(uint64_t);
    PKV_CUDA(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    select_kernel<<<nq, 512, smem, s>>>(ws.d_cand, ws.d_cnt, ws.d_thr_key, ws.d_thr_f, ws.d_q_mag_f, ws.d_status,
                                        (uint32_t)ws.cap, k, fs, ws.d_pend_cnt, clear_deferred ? ws.d_defer_cnt : nullptr,
                                        clear_tail ? 1 : 0, ix.d_ids,
                                        ix.row_base, d_ids, d_dist, d_counts, guess_rank);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_finalize(const Index &ix, Workspace &ws, int nq, int k, int64_t *d_ids, float *d_dist, int32_t *d_counts,
                    cudaStream_t s) {
    if (nq <= 0) return PKV_OK;
    finalize_kernel<<<nq, 128, 0, s>>>(ws.d_cand, ws.d_cnt, (uint32_t)ws.cap, k, ix.d_ids, ix.row_base, d_ids, d_dist,
                                       d_counts);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_row_mags(Index &ix, int64_t row_begin, int64_t row_end, cudaStream_t s) {
    if (row_end <= row_begin) return PKV_OK;
    const int warps = 8;
    const int64_t blocks = (row_end - row_begin + warps - 1) / warps;
    row_mags_kernel<<<(unsigned)blocks, warps * 32, 0, s>>>(ix.d_data, ix.pitch, ix.dim_pad, ix.dtype, row_begin,
                                                            row_end, ix.d_mag_i, ix.d_mag_f);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

// ---- threshold-guess sample: rows 0, stride, 2 stride, ... of the sealed corpus, copied into one contiguous block
__global__ void sample_copy_kernel(const uint8_t *data, int64_t pitch, int64_t stride, int64_t n, uint8_t *out) {
    const int64_t r = blockIdx.x;
    if (r >= n) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(data + (size_t)(r * stride) * (size_t)pitch);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (size_t)r * (size_t)pitch);
    for (int i = threadIdx.x; i < (int)(pitch / 16); i += blockDim.x) dst[i] = src[i];
}

constexpr int64_t SAMPLE_ROWS = 3968;  // <= candidate capacity - k for every k the guessed start serves (k <= 128)

int build_sample(Index &ix, cudaStream_t s) {
    const int64_t N = ix.rows;
    if (N < 16 * SAMPLE_ROWS) {  // small corpora use the plain schedule
        ix.sample_rows = 0;
        return PKV_OK;
    }
    if (!ix.d_sample) {
        PKV_CUDA(cudaMalloc((void **)&ix.d_sample, (size_t)SAMPLE_ROWS * ix.pitch));
        if (ix.dtype == PKV_I8) PKV_CUDA(cudaMalloc((void **)&ix.d_sample_mag_i, sizeof(int32_t) * SAMPLE_ROWS));
    }
    const int64_t stride = N / SAMPLE_ROWS;
    sample_copy_kernel<<<(unsigned)SAMPLE_ROWS, 128, 0, s>>>(ix.d_data, ix.pitch, stride, SAMPLE_ROWS, ix.d_sample);
    if (ix.dtype == PKV_I8) {
        const int warps = 8;
        row_mags_kernel<<<(unsigned)((SAMPLE_ROWS + warps - 1) / warps), warps * 32, 0, s>>>(
            ix.d_sample, ix.pitch, ix.dim_pad, ix.dtype, 0, SAMPLE_ROWS, ix.d_sample_mag_i, nullptr);
    }
    PKV_CUDA(cudaGetLastError());
    ix.sample_rows = SAMPLE_ROWS;
    ix.sample_of_rows = N;
    return PKV_OK;
}

int launch_fill_ids(int64_t *d_ids, int64_t begin, int64_t end, int64_t base, cudaStream_t s) {
    if (end <= begin) return PKV_OK;
    fill_ids_kernel<<<(unsigned)((end - begin + 255) / 256), 256, 0, s>>>(d_ids, begin, end, base);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_merge(const int64_t *d_ids, const float *d_dist, int parts, int nq, int k, int64_t *o_ids, float *o_dist,
                 int32_t *o_counts, cudaStream_t s) {
    if (nq <= 0) return PKV_OK;
    merge_kernel<<<nq, 256, parts * sizeof(int), s>>>(d_ids, d_dist, parts, nq, k, o_ids, o_dist, o_counts);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_pack_topk(const int64_t *d_ids, const float *d_dist, int64_t n, void *d_packed, cudaStream_t s) {
    if (n <= 0) return PKV_OK;
    pack_topk_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_ids, d_dist, n, (uint32_t *)d_packed);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_merge_packed(const void *d_packed, int parts, int nq, int k, int64_t *o_ids, float *o_dist, int32_t *o_counts,
                        cudaStream_t s) {
    if (nq <= 0) return PKV_OK;
    merge_packed_kernel<<<nq, 256, parts * sizeof(int), s>>>((const uint32_t *)d_packed, parts, nq, k, o_ids, o_dist,
                                                             o_counts);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_aggregate(const float *d_dist, const int64_t *d_item, const float *d_w, int64_t n, int64_t n_items, int agg,
                     double *d_out, cudaStream_t s) {
    if (n_items <= 0) return PKV_OK;
    double *den = nullptr;
    unsigned long long *cnt = nullptr;
    PKV_CUDA(cudaMallocAsync((void **)&den, sizeof(double) * n_items, s));
    PKV_CUDA(cudaMallocAsync((void **)&cnt, sizeof(unsigned long long) * n_items, s));
    const unsigned ib = (unsigned)((n_items + 255) / 256);
    agg_init_kernel<<<ib, 256, 0, s>>>(d_out, den, cnt, n_items, d_w ? PKV_AGG_AVG : agg);
    if (n > 0)
        agg_accum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_dist, d_item, d_w, n, n_items, agg, d_out, den,
                                                                    cnt);
    agg_finish_kernel<<<ib, 256, 0, s>>>(d_out, den, cnt, n_items, agg, d_w ? 1 : 0);
    PKV_CUDA(cudaGetLastError());
    PKV_CUDA(cudaFreeAsync(den, s));
    PKV_CUDA(cudaFreeAsync(cnt, s));
    return PKV_OK;
}

int launch_absmax(const float *d_values, int64_t n, float *d_out, cudaStream_t s) {
    PKV_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float), s));
    if (n <= 0) return PKV_OK;
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    absmax_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_values, n, (unsigned int *)d_out);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_quantize(const float *d_values, int64_t n, float scale, int8_t *d_codes, cudaStream_t s) {
    if (n <= 0) return PKV_OK;
    int64_t blocks = (n + 1023) / 1024;
    if (blocks > 148 * 16) blocks = 148 * 16;
    quantize_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_values, n, scale, d_codes);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

}  // namespace pkv
