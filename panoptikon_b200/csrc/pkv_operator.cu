// pkv_operator.cu — the PQL vector operator's grouped form: score, aggregate per group, rank.
//
// What the reference's filter compilers render as SQL:
//   dist_{cte} AS MATERIALIZED (SELECT ..., vec_distance_*(payload, ?) AS d [, w] FROM candidates)
//   SELECT ..., row_number() OVER (ORDER BY AGG(d) ASC) AS order_rank FROM dist_{cte} GROUP BY file_id
// (builder/filters/exact.rs:67-80,106-165; builder.rs:757-771; item_similarity.rs:503-581 for
// `similar_to`, whose AGG runs over every (target vector, candidate vector) pair).
// Here: untruncated distances (CUDA-core dense scan) -> atomic per-group MIN/MAX/AVG or
// SUM(d*w)/SUM(w) in double, skipping NULL (NaN) like SQLite -> ranking of the groups that own at
// least one candidate row, ascending aggregate, NULL aggregates last, ties by group id.
#include "pkv_device.cuh"

namespace pkv {

namespace {

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

__global__ void grp_init_kernel(double *acc, double *den, uint32_t *valid, uint32_t *present, int64_t n, int agg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double INF = __longlong_as_double(0x7ff0000000000000ll);
    acc[i] = agg == PKV_AGG_MIN ? INF : (agg == PKV_AGG_MAX ? -INF : 0.0);
    den[i] = 0.0;
    valid[i] = 0;
    present[i] = 0;
}

// one thread per (query, row) pair of a [nq][rows] distance block
__global__ void grp_accum_kernel(const float *dist, int64_t rows, int nq, const int64_t *group_of_row, const float *w,
                                 int64_t n_groups, int agg, double *acc, double *den, uint32_t *valid,
                                 uint32_t *present, const PairRules rules) {
    const int64_t total = rows * nq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i % rows;
        const int64_t g = group_of_row[r];
        if (g < 0 || g >= n_groups) continue;  // not a candidate (context filter, the similar_to target itself)
        const int64_t qi = i / rows;
        if (rules.row_modality) {
            // similar_to's pair rules (item_similarity.rs:468-488): without clip_xmodal only the image setter's rows
            // exist on either side; with it, image-image pairs go when !xmodal_i2i, text-text pairs when !xmodal_t2t
            const int rm = rules.row_modality[r], qm = rules.q_modality[qi];
            if (!rules.clip_xmodal && (rm != 0 || qm != 0)) continue;
            if (rules.skip_i2i && rm == 0 && qm == 0) continue;
            if (rules.skip_t2t && rm == 1 && qm == 1) continue;
        }
        present[g] = 1;
        const float df = dist[i];
        if (df != df) continue;  // SQL NULL is skipped by every aggregate
        const double d = (double)df;
        if (w) {
            // pow(conf_main * conf_other, wc) * pow(lang_main * lang_other, wl) factorises into a per-row and a
            // per-target weight (item_similarity.rs:523-560)
            const double ww = (double)w[r] * (rules.q_weights ? (double)rules.q_weights[qi] : 1.0);
            atomicAdd(acc + g, d * ww);
            atomicAdd(den + g, ww);
        } else if (agg == PKV_AGG_AVG) {
            atomicAdd(acc + g, d);
            atomicAdd(den + g, 1.0);
        } else if (agg == PKV_AGG_MIN) {
            atomic_min_f64(acc + g, d);
        } else {
            atomic_max_f64(acc + g, d);
        }
        atomicAdd(valid + g, 1u);
    }
}

// ordered 64-bit image of a double: ascending u64 == ascending double, NaN (NULL) after +inf
__device__ __forceinline__ uint64_t ordered_f64(double d) {
    if (d != d) return 0xFFFFFFFFFFFFFFFEull;
    d += 0.0;
    uint64_t b = (uint64_t)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unordered_f64(uint64_t o) {
    if (o == 0xFFFFFFFFFFFFFFFEull) return __longlong_as_double(0x7ff8000000000000ll);
    uint64_t b = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
    return __longlong_as_double((long long)b);
}

__global__ void grp_keys_kernel(const double *acc, const double *den, const uint32_t *valid, const uint32_t *present,
                                int64_t n, int agg, int weighted, uint64_t *keys, uint32_t *ids) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t key = 0xFFFFFFFFFFFFFFFFull;  // absent group: never ranked
    if (present[i]) {
        double v;
        if (valid[i] == 0)
            v = __longlong_as_double(0x7ff8000000000000ll);  // only NULL distances: NULL aggregate, ranked last
        else if (weighted || agg == PKV_AGG_AVG)
            v = acc[i] / den[i];
        else
            v = acc[i];
        key = ordered_f64(v);
    }
    keys[i] = key;
    ids[i] = (uint32_t)i;
}

// Each CTA sorts a slice of SLICE (key, id) pairs and keeps its best `keep`; applied repeatedly
// until one slice is left.  Order: key ascending, then id ascending.
constexpr int SLICE = 4096;
__global__ void __launch_bounds__(512) rank_pass_kernel(const uint64_t *keys, const uint32_t *ids, int64_t n, int keep,
                                                        uint64_t *okeys, uint32_t *oids) {
    __shared__ uint64_t s_k[SLICE];
    __shared__ uint32_t s_i[SLICE];
    const int64_t base = (int64_t)blockIdx.x * SLICE;
    for (int i = threadIdx.x; i < SLICE; i += blockDim.x) {
        const int64_t g = base + i;
        s_k[i] = g < n ? keys[g] : 0xFFFFFFFFFFFFFFFFull;
        s_i[i] = g < n ? ids[g] : 0xFFFFFFFFu;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= SLICE; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = threadIdx.x; i < SLICE / 2; i += blockDim.x) {
                const uint32_t pos = 2 * i - (i & (stride - 1));
                const uint64_t ka = s_k[pos], kb = s_k[pos + stride];
                const uint32_t ia = s_i[pos], ib = s_i[pos + stride];
                const bool up = (pos & size) == 0;
                const bool gt = ka > kb || (ka == kb && ia > ib);
                if (gt == up) {
                    s_k[pos] = kb;
                    s_k[pos + stride] = ka;
                    s_i[pos] = ib;
                    s_i[pos + stride] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < keep; i += blockDim.x) {
        okeys[(int64_t)blockIdx.x * keep + i] = s_k[i];
        oids[(int64_t)blockIdx.x * keep + i] = s_i[i];
    }
}

__global__ void rank_emit_kernel(const uint64_t *keys, const uint32_t *ids, int avail, int offset, int limit,
                                 int64_t *out_groups, double *out_agg, int32_t *out_count) {
    int produced = 0;
    for (int i = threadIdx.x; i < limit; i += blockDim.x) {
        const int src = offset + i;
        int64_t g = -1;
        double v = __longlong_as_double(0x7ff8000000000000ll);
        if (src < avail && keys[src] != 0xFFFFFFFFFFFFFFFFull) {
            g = (int64_t)ids[src];
            v = unordered_f64(keys[src]);
        }
        out_groups[i] = g;
        out_agg[i] = v;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < limit; ++i)
            if (offset + i < avail && keys[offset + i] != 0xFFFFFFFFFFFFFFFFull) produced++;
        *out_count = produced;
    }
}

}  // namespace

// dist: [nq][rows] on the device.  Ranks groups and writes `limit` entries starting at `offset`.
int rank_groups(const float *d_dist, int64_t rows, int nq, const int64_t *d_group_of_row, const float *d_weights,
                int64_t n_groups, int agg, int offset, int limit, int64_t *d_out_groups, double *d_out_agg,
                int32_t *d_out_count, cudaStream_t s, const PairRules *rules_or_null) {
    PairRules rules;
    memset(&rules, 0, sizeof(rules));
    if (rules_or_null) rules = *rules_or_null;
    const int keep = offset + limit;
    if (keep > SLICE / 2) return fail(PKV_ERR_INVALID, "offset + limit must not exceed %d", SLICE / 2);
    double *acc = nullptr, *den = nullptr;
    uint32_t *valid = nullptr, *present = nullptr, *ids[2] = {nullptr, nullptr};
    uint64_t *keys[2] = {nullptr, nullptr};
    const int64_t ng = n_groups > 0 ? n_groups : 1;
    cudaError_t e = cudaMallocAsync((void **)&acc, sizeof(double) * ng, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&den, sizeof(double) * ng, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&valid, sizeof(uint32_t) * ng, s);
    if (e == cudaSuccess) e = cudaMallocAsync((void **)&present, sizeof(uint32_t) * ng, s);
    const int64_t slices0 = (ng + SLICE - 1) / SLICE;
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
        const int64_t cap = b == 0 ? ng : slices0 * keep;
        e = cudaMallocAsync((void **)&keys[b], sizeof(uint64_t) * (cap > 0 ? cap : 1), s);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&ids[b], sizeof(uint32_t) * (cap > 0 ? cap : 1), s);
    }
    int st = PKV_OK;
    if (e != cudaSuccess) {
        st = fail(e == cudaErrorMemoryAllocation ? PKV_ERR_OOM : PKV_ERR_CUDA, "rank_groups allocation failed: %s",
                  cudaGetErrorString(e));
    } else {
        const unsigned gb = (unsigned)((ng + 255) / 256);
        grp_init_kernel<<<gb, 256, 0, s>>>(acc, den, valid, present, ng, d_weights ? PKV_AGG_AVG : agg);
        const int64_t total = rows * nq;
        if (total > 0) {
            int64_t blocks = (total + 255) / 256;
            if (blocks > 148 * 32) blocks = 148 * 32;
            grp_accum_kernel<<<(unsigned)blocks, 256, 0, s>>>(d_dist, rows, nq, d_group_of_row, d_weights, n_groups, agg,
                                                             acc, den, valid, present, rules);
        }
        grp_keys_kernel<<<gb, 256, 0, s>>>(acc, den, valid, present, ng, agg, d_weights ? 1 : 0, keys[0], ids[0]);
        int64_t n = ng;
        int cur = 0;
        for (;;) {
            const int64_t slices = (n + SLICE - 1) / SLICE;
            rank_pass_kernel<<<(unsigned)slices, 512, 0, s>>>(keys[cur], ids[cur], n, keep, keys[cur ^ 1], ids[cur ^ 1]);
            n = slices * keep;
            cur ^= 1;
            if (slices == 1) break;
        }
        rank_emit_kernel<<<1, 128, 0, s>>>(keys[cur], ids[cur], (int)n, offset, limit, d_out_groups, d_out_agg,
                                           d_out_count);
        cudaError_t le = cudaGetLastError();
        if (le != cudaSuccess) st = fail(PKV_ERR_CUDA, "rank_groups launch failed: %s", cudaGetErrorString(le));
    }
    cudaFreeAsync(acc, s);
    cudaFreeAsync(den, s);
    cudaFreeAsync(valid, s);
    cudaFreeAsync(present, s);
    for (int b = 0; b < 2; ++b) {
        cudaFreeAsync(keys[b], s);
        cudaFreeAsync(ids[b], s);
    }
    return st;
}

// gather stored rows by position (similar_to reads its target's own stored vectors)
__global__ void gather_rows_kernel(const uint8_t *data, int64_t pitch, int row_bytes, const int64_t *rows, int n,
                                   int64_t n_rows, uint8_t *out) {
    const int r = blockIdx.x;
    if (r >= n) return;
    const int64_t src = rows[r];
    for (int i = threadIdx.x; i < row_bytes; i += blockDim.x)
        out[(size_t)r * row_bytes + i] = (src >= 0 && src < n_rows) ? data[(size_t)src * pitch + i] : 0;
}

// modality and weight of the target's rows (the "main" side of similar_to's self-join)
__global__ void gather_attrs_kernel(const int64_t *rows, int n, int64_t n_rows, const uint8_t *modality, const float *weights,
                                    uint8_t *q_mod, float *q_w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = rows[i];
    const bool ok = r >= 0 && r < n_rows;
    q_mod[i] = (ok && modality) ? modality[r] : 0;
    q_w[i] = (ok && weights) ? weights[r] : 1.0f;
}

int launch_gather_attrs(const int64_t *d_rows, int n, int64_t n_rows, const uint8_t *d_modality, const float *d_weights,
                        uint8_t *d_q_mod, float *d_q_w, cudaStream_t s) {
    if (n <= 0) return PKV_OK;
    gather_attrs_kernel<<<(n + 127) / 128, 128, 0, s>>>(d_rows, n, n_rows, d_modality, d_weights, d_q_mod, d_q_w);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

int launch_gather_rows(const Index &ix, const int64_t *d_rows, int n, void *d_out, cudaStream_t s) {
    if (n <= 0) return PKV_OK;
    gather_rows_kernel<<<n, 128, 0, s>>>(ix.d_data, ix.pitch, ix.dim * ix.elem, d_rows, n, ix.sealed_rows,
                                         (uint8_t *)d_out);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

}  // namespace pkv
