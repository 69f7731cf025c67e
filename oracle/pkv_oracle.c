/*
 * pkv_oracle.c — see pkv_oracle.h.  TEST INFRASTRUCTURE ONLY (checker / CPU baseline).
 *
 * Build: gcc -O3 -std=c11 -fPIC -shared -ffp-contract=off -fno-fast-math -pthread
 * (no -march=native, no FMA contraction: the accumulation order and the
 * single rounding per operation are part of what is being restated).
 */
#define _GNU_SOURCE
#include "pkv_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ codec */

/* vector_quants.rs:1446 */
static const float INT8_MAX_CODE = 127.0f;

/* vector_quants.rs:1465-1471 */
float orc_scale_from_absmax(float absmax) {
    if (absmax > 0.0f && isfinite(absmax)) return absmax / INT8_MAX_CODE;
    return 1.0f;
}

/* vector_quants.rs:1449-1451 (f32::to_le_bytes) */
void orc_scale_artifact(float scale, uint8_t out[4]) {
    uint32_t bits;
    memcpy(&bits, &scale, 4);
    out[0] = (uint8_t)(bits & 0xff);
    out[1] = (uint8_t)((bits >> 8) & 0xff);
    out[2] = (uint8_t)((bits >> 16) & 0xff);
    out[3] = (uint8_t)((bits >> 24) & 0xff);
}

static float f32_from_le(const uint8_t *p) {
    uint32_t bits = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
                    ((uint32_t)p[3] << 24);
    float v;
    memcpy(&v, &bits, 4);
    return v;
}

/* vector_quants.rs:1456-1460 */
int orc_artifact_scale(const uint8_t *artifact, size_t len, float *scale) {
    if (len != 4) return 0;
    float s = f32_from_le(artifact);
    if (isfinite(s) && s > 0.0f) {
        *scale = s;
        return 1;
    }
    return 0;
}

/* vector_quants.rs:1474-1483 — `>` never lets a NaN replace the running max */
float orc_blob_absmax(const uint8_t *blob, size_t len) {
    float absmax = 0.0f;
    for (size_t i = 0; i + 4 <= len; i += 4) {
        float v = fabsf(f32_from_le(blob + i));
        if (v > absmax) absmax = v;
    }
    return absmax;
}

/* vector_quants.rs:1489-1497: (x / s).round_ties_even().clamp(-128, 127) as i8.
 * nearbyintf under the default FE_TONEAREST mode is round-half-to-even; Rust's
 * clamp keeps NaN and `as i8` maps NaN to 0. */
void orc_quantize_int8(const uint8_t *blob, size_t len, float scale, uint8_t *out) {
    size_t n = len / 4;
    for (size_t i = 0; i < n; i++) {
        float value = f32_from_le(blob + 4 * i);
        float r = nearbyintf(value / scale);
        int8_t code;
        if (isnan(r)) {
            code = 0;
        } else {
            if (r < -128.0f) r = -128.0f;
            if (r > INT8_MAX_CODE) r = INT8_MAX_CODE;
            code = (int8_t)r;
        }
        out[i] = (uint8_t)code;
    }
}

/* -------------------------------------------------------------- distances */
/* sqlite-vec 0.1.9 scalar loops: three sequential f32 accumulators; the C
 * `sqrt` promotes to double, the whole return expression is evaluated in
 * double and rounded once to f32. */

float orc_distance_cosine_f32(const float *a, const float *b, size_t d) {
    float dot = 0, aMag = 0, bMag = 0;
    for (size_t i = 0; i < d; i++) {
        dot += a[i] * b[i];
        aMag += a[i] * a[i];
        bMag += b[i] * b[i];
    }
    return (float)(1 - (dot / (sqrt(aMag) * sqrt(bMag))));
}

float orc_distance_l2_f32(const float *a, const float *b, size_t d) {
    float res = 0;
    for (size_t i = 0; i < d; i++) {
        float t = a[i] - b[i];
        res += t * t;
    }
    return (float)sqrt(res);
}

float orc_distance_cosine_i8(const int8_t *a, const int8_t *b, size_t d) {
    float dot = 0, aMag = 0, bMag = 0;
    for (size_t i = 0; i < d; i++) {
        dot += (float)((int)a[i] * (int)b[i]);
        aMag += (float)((int)a[i] * (int)a[i]);
        bMag += (float)((int)b[i] * (int)b[i]);
    }
    return (float)(1 - (dot / (sqrt(aMag) * sqrt(bMag))));
}

float orc_distance_l2_i8(const int8_t *a, const int8_t *b, size_t d) {
    float res = 0;
    for (size_t i = 0; i < d; i++) {
        float t = (float)((int)a[i] - (int)b[i]);
        res += t * t;
    }
    return (float)sqrt(res);
}

float orc_distance_dot_f32(const float *a, const float *b, size_t d) {
    float dot = 0;
    for (size_t i = 0; i < d; i++) dot += a[i] * b[i];
    return 0.0f - dot; /* +0.0 for a zero dot product: equal distances must compare and print alike */
}

float orc_distance_dot_i8(const int8_t *a, const int8_t *b, size_t d) {
    float dot = 0;
    for (size_t i = 0; i < d; i++) dot += (float)((int)a[i] * (int)b[i]);
    return 0.0f - dot;
}

float orc_half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal: normalise */
            int e = -1;
            do {
                man <<= 1;
                e++;
            } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

/* ------------------------------------------------------------------ top-k */

typedef struct {
    float d;
    int64_t row;
} hit_t;

/* Total order of SURVEY App. A.4: finite/inf ascending, NaN last, ties by row. */
static int hit_less(hit_t x, hit_t y) {
    int xn = isnan(x.d), yn = isnan(y.d);
    if (xn != yn) return yn; /* non-NaN first */
    if (!xn) {
        if (x.d < y.d) return 1;
        if (x.d > y.d) return 0;
    }
    return x.row < y.row;
}

/* max-heap on hit_less: root = worst kept hit */
static void heap_sift_down(hit_t *h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < n && hit_less(h[m], h[l])) m = l;
        if (r < n && hit_less(h[m], h[r])) m = r;
        if (m == i) return;
        hit_t t = h[i];
        h[i] = h[m];
        h[m] = t;
        i = m;
    }
}
static void heap_sift_up(hit_t *h, int i) {
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!hit_less(h[p], h[i])) return;
        hit_t t = h[i];
        h[i] = h[p];
        h[p] = t;
        i = p;
    }
}

static float row_distance(const void *corpus, int64_t row, int dim, int dtype, const void *query,
                          int metric, float *scratch_a, const float *query_f32) {
    if (dtype == ORC_F32) {
        const float *a = (const float *)corpus + (size_t)row * dim;
        const float *b = (const float *)query;
        if (metric == ORC_COSINE) return orc_distance_cosine_f32(a, b, dim);
        if (metric == ORC_L2) return orc_distance_l2_f32(a, b, dim);
        return orc_distance_dot_f32(a, b, dim);
    } else if (dtype == ORC_I8) {
        const int8_t *a = (const int8_t *)corpus + (size_t)row * dim;
        const int8_t *b = (const int8_t *)query;
        if (metric == ORC_COSINE) return orc_distance_cosine_i8(a, b, dim);
        if (metric == ORC_L2) return orc_distance_l2_i8(a, b, dim);
        return orc_distance_dot_i8(a, b, dim);
    } else {
        /* fp16 corpus extension (no reference storage analogue):
         * f32 formulas on the half-rounded values widened back to f32 */
        const uint16_t *a = (const uint16_t *)corpus + (size_t)row * dim;
        for (int i = 0; i < dim; i++) scratch_a[i] = orc_half_to_float(a[i]);
        if (metric == ORC_COSINE) return orc_distance_cosine_f32(scratch_a, query_f32, dim);
        if (metric == ORC_L2) return orc_distance_l2_f32(scratch_a, query_f32, dim);
        return orc_distance_dot_f32(scratch_a, query_f32, dim);
    }
}

static size_t elem_size(int dtype) { return dtype == ORC_F32 ? 4 : (dtype == ORC_I8 ? 1 : 2); }

typedef struct {
    const void *corpus;
    int64_t n;
    int dim, dtype;
    const void *queries;
    int nq, metric, k;
    const uint64_t *bitmap;
    int64_t bitmap_stride;
    int64_t *out_rows;
    float *out_dist;
    int32_t *out_counts;
    int q_begin, q_step;
} job_t;

static int cmp_hits(const void *x, const void *y) {
    hit_t a = *(const hit_t *)x, b = *(const hit_t *)y;
    if (hit_less(a, b)) return -1;
    if (hit_less(b, a)) return 1;
    return 0;
}

static void *topk_worker(void *arg) {
    job_t *j = (job_t *)arg;
    hit_t *heap = (hit_t *)malloc(sizeof(hit_t) * (size_t)(j->k > 0 ? j->k : 1));
    float *scratch = (float *)malloc(sizeof(float) * (size_t)j->dim * 2);
    float *qf = scratch + j->dim;
    size_t es = elem_size(j->dtype);
    for (int q = j->q_begin; q < j->nq; q += j->q_step) {
        const void *query = (const uint8_t *)j->queries + (size_t)q * j->dim * es;
        if (j->dtype == ORC_F16)
            for (int i = 0; i < j->dim; i++) qf[i] = orc_half_to_float(((const uint16_t *)query)[i]);
        const uint64_t *bm = j->bitmap ? j->bitmap + (size_t)q * j->bitmap_stride : NULL;
        int cnt = 0;
        for (int64_t r = 0; r < j->n; r++) {
            if (bm && !((bm[r >> 6] >> (r & 63)) & 1)) continue;
            hit_t h;
            h.d = row_distance(j->corpus, r, j->dim, j->dtype, query, j->metric, scratch, qf);
            h.row = r;
            if (cnt < j->k) {
                heap[cnt] = h;
                heap_sift_up(heap, cnt);
                cnt++;
            } else if (hit_less(h, heap[0])) {
                heap[0] = h;
                heap_sift_down(heap, cnt, 0);
            }
        }
        qsort(heap, (size_t)cnt, sizeof(hit_t), cmp_hits);
        for (int i = 0; i < j->k; i++) {
            size_t o = (size_t)q * j->k + i;
            if (i < cnt) {
                j->out_rows[o] = heap[i].row;
                j->out_dist[o] = heap[i].d;
            } else {
                j->out_rows[o] = -1;
                j->out_dist[o] = NAN;
            }
        }
        j->out_counts[q] = cnt;
    }
    free(heap);
    free(scratch);
    return NULL;
}

int orc_topk(const void *corpus, int64_t n, int dim, int dtype, const void *queries, int nq,
             int metric, int k, const uint64_t *bitmap, int64_t bitmap_stride, int threads,
             int64_t *out_rows, float *out_dist, int32_t *out_counts) {
    if (n < 0 || dim < 1 || nq < 0 || k < 1 || dtype < 0 || dtype > 2 || metric < 0 || metric > 2)
        return -1;
    if (threads < 1) threads = 1;
    if (threads > nq) threads = nq > 0 ? nq : 1;
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * (size_t)threads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) {
        job_t j = {corpus,        n,      dim,      dtype,    queries,  nq,         metric, k,
                   bitmap,        bitmap_stride, out_rows, out_dist, out_counts, t,      threads};
        jobs[t] = j;
    }
    if (threads == 1) {
        topk_worker(&jobs[0]);
    } else {
        for (int t = 0; t < threads; t++) pthread_create(&tids[t], NULL, topk_worker, &jobs[t]);
        for (int t = 0; t < threads; t++) pthread_join(tids[t], NULL);
    }
    free(jobs);
    free(tids);
    return 0;
}

int orc_distances(const void *corpus, int64_t n, int dim, int dtype, const void *query, int metric,
                  float *out) {
    if (n < 0 || dim < 1 || dtype < 0 || dtype > 2 || metric < 0 || metric > 2) return -1;
    float *scratch = (float *)malloc(sizeof(float) * (size_t)dim * 2);
    float *qf = scratch + dim;
    if (dtype == ORC_F16)
        for (int i = 0; i < dim; i++) qf[i] = orc_half_to_float(((const uint16_t *)query)[i]);
    for (int64_t r = 0; r < n; r++)
        out[r] = row_distance(corpus, r, dim, dtype, query, metric, scratch, qf);
    free(scratch);
    return 0;
}

/* ------------------------------------------------------------ aggregation */

int orc_aggregate(const float *dist, const int64_t *item_of_row, const float *weights, int64_t n,
                  int64_t n_items, int agg, double *out) {
    if (n < 0 || n_items < 0 || agg < 0 || agg > 2) return -1;
    double *den = (double *)calloc((size_t)(n_items > 0 ? n_items : 1), sizeof(double));
    int64_t *cnt = (int64_t *)calloc((size_t)(n_items > 0 ? n_items : 1), sizeof(int64_t));
    for (int64_t i = 0; i < n_items; i++) out[i] = 0.0;
    for (int64_t r = 0; r < n; r++) {
        int64_t it = item_of_row[r];
        if (it < 0 || it >= n_items) continue;
        double d = (double)dist[r];
        if (isnan(d)) continue; /* SQL NULL is skipped by every aggregate */
        if (weights) { /* SUM(d*w)/SUM(w), exact.rs:73-79 */
            double w = (double)weights[r];
            out[it] += d * w;
            den[it] += w;
            cnt[it]++;
        } else if (agg == ORC_AGG_AVG) {
            out[it] += d;
            cnt[it]++;
        } else if (agg == ORC_AGG_MIN) {
            if (cnt[it] == 0 || d < out[it]) out[it] = d;
            cnt[it]++;
        } else {
            if (cnt[it] == 0 || d > out[it]) out[it] = d;
            cnt[it]++;
        }
    }
    for (int64_t i = 0; i < n_items; i++) {
        if (cnt[i] == 0)
            out[i] = NAN;
        else if (weights)
            out[i] = out[i] / den[i];
        else if (agg == ORC_AGG_AVG)
            out[i] = out[i] / (double)cnt[i];
    }
    free(den);
    free(cnt);
    return 0;
}
