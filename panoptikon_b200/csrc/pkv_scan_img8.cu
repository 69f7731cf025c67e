// pkv_scan_img8.cu — tensor-core scan of floating-point indexes (f32 / f16 rows) over an INT8 IMAGE:
// a conservative tcgen05 kind::i8 FILTER plus exact re-scoring of the survivors from the stored rows.
//
// Replaces vec_distance_cosine / vec_distance_L2 over `embeddings.embedding` blobs
// (pql/builder/filters/image_embeddings.rs:321-337, text_embeddings.rs:386-393).  It is the reference's
// own idea of an int8 "quant" shadow (db/vector_quants.rs:1446-1503, docs/vector-int8-quant.md) turned
// into a filter whose results stay exact: every score that is reported is computed from the f32 rows
// by rescore_kernel (pkv_scan_tc_f32.cu) with the same summation order as the CUDA-core scan.
//
// Image: every row a is stored as codes c = clamp(rint(a / s_a), -127, 127) with ITS OWN scale s_a = |a| / U,
// i.e. the DIRECTION of the row is what is quantised: |a|/s_a = U is the same for every row, so a bound that must hold
// for the 32 rows an epilogue warp sees at once loses nothing to the spread of the rows' scales.  U = 127 / p_ref
// where p_ref is a high quantile (mean + 1.3 sigma over the first sealed batch) of the rows' peakiness
// max|a_i| / |a|; components of peakier rows saturate, which only enlarges that row's own error term.  Per row four
// floats {u = |a|/s_a, v = |c|, w = |a - s_a c|/s_a, r = 1/s_a}.  Queries are quantised per query with
// s_q = max|q_i| / 127 (e_q = q - s_q c_q).  With acc = c . c_q (exact, s32):
//     a.q = s_a s_q acc + s_a (c . e_q) + (a - s_a c) . q
//     |a.q - s_a s_q acc| <= s_a |c| |e_q| + |a - s_a c| |q|                     (Cauchy-Schwarz)
// A pair can be in the top-k only if a.q >= X(a, q) (X from the exact k-th best distance, per metric), so
// it may be dropped only if      acc < X / (s_a s_q) - v |e_q|/s_q - w |q|/s_q   (minus rounding slack).
// For unit-norm 768-d Gaussian-like data the slack is ~1.5 % of |a||q|: ~5 survivors per true top-k row.
//
// Kernel = the int8 scan with the queries resident in TMEM (pkv_scan_ts.cu): A = 128 queries per CTA in
// tensor memory, B = row tiles streamed by TMA, accumulator lane = query / column = row.  Two shapes:
//   PAIR   cta_group::2, M = 256 queries, each CTA stages 64 rows of a 128-row tile; up to 4 query groups
//          per launch share row tiles through L2                                    (batches > 128)
//   single cta_group::1, M = 128 queries, one CTA per SM stages whole 128-row tiles  (batches <= 128:
//          HBM-bound, half the bytes of the fp16 image and a quarter of the f32 rows)
//
// Algorithmic bytes per row per pass: dim_pad8 (+16 B row figures); ops: 2 * queries * dim_pad8.
#include "pkv_tc.cuh"

namespace pkv {

namespace {

constexpr int QM_CTA = 128;
constexpr int CHUNK_BYTES = 128;
constexpr int MAX_STAGES = 26;
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int RS_WARPS = 2;  // exact re-scoring warps (fused mode): consume the survivors of this CTA's epilogue
constexpr int RS_SLOTS = 2;  // rows in flight per re-scoring warp (shared-memory buffers filled by cp.async.bulk)
constexpr int RS_CLAIM = 16; // survivors a re-scoring warp takes off the ring per round
constexpr int RS_WARP0 = 2 + EPI_WARPS;
constexpr int IMG_THREADS = 64 + EPI_THREADS + RS_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int HOLD_CAP = 64;
constexpr int HOLD_FLUSH = 32;
constexpr float BOUND_CLAMP = 1.07e9f;
constexpr uint32_t RING_CAP = 256;   // filter survivors in flight between the epilogue and the re-scoring warps
constexpr uint32_t RING_MASK = RING_CAP - 1;

struct ImgArgs {
    const float4 *row_meta;  // [rows] {|a|/s_a, |c|, |a - s_a c|/s_a, 1/s_a}
    const int8_t *q8;        // [nq][dim_pad8] query codes
    const float4 *q_meta;    // [nq] {1/s_q, |e_q|/s_q, |q|/s_q, |q|^2}
    PendDev pend;            // unfused mode: survivors parked for rescore_kernel (this chunk's)
    PendDev dlist;           // defer mode: pairs parked for the whole search (rescore_deferred_kernel)
    int dim_pad8;
    int fused;               // survivors are re-scored by this kernel's own warps (no pend lists, no second launch)
    int defer;               // only the LIKELY candidates (approximate score beats the threshold) are re-scored right
                             // away (in-kernel when fused, else by rescore_kernel behind the chunk); pairs inside the
                             // filter's error band (and, fused, whatever the ring cannot take) are parked with their dot
                             // product and re-checked against the FINAL thresholds at the end of the search
    int rows_f16;            // stored rows are fp16 (f16 index), else f32
    unsigned long long *dbg; // PKV_TILE_TIMING builds only: per-warp clock sums of the accumulator hand-off (tools/gpu/tile_timing.py)
    int park_lean;           // parked pairs carry (row, dot) only: the deferred pass reads the row figures itself
    int epi_exact;           // the epilogue culls with the EXACT per-pair bound in registers (row figures shuffled from the
                             // lane that prefetched them) before a pair enters the hold list: the flush has no
                             // dependent global load left and sees ~half the entries
};

struct ImgShared {
    uint64_t full[MAX_STAGES];
    uint64_t empty[MAX_STAGES];
    uint64_t tmem_full[3];
    uint64_t tmem_empty[3];
    uint32_t tmem_base;
    uint32_t ring_head;  // tickets handed to producers (epilogue lanes)
    uint32_t ring_tail;  // tickets handed to consumers (re-scoring warps)
    uint32_t epi_done;   // epilogue warps that have pushed their last survivor
    uint32_t hold_cnt[EPI_WARPS];  // (img8_epi = 0 only: the atomically filled hold lists of round 1)
    uint64_t rs_bar[RS_WARPS][RS_SLOTS];  // completion of the re-scoring warps' row copies
    alignas(16) float4 qm[QM_CTA];  // per query {1/s_q, |e_q|/s_q, |q|/s_q, |q|^2} (fixed for the launch)
    float thr[QM_CTA];              // per query: filter threshold (live mode: refreshed tile by tile)
    uint32_t hold_row[EPI_WARPS][HOLD_CAP];
    int hold_dot[EPI_WARPS][HOLD_CAP];
    uint32_t hold_col[EPI_WARPS][HOLD_CAP];
    // bounded multi-producer multi-consumer ring (per-slot sequence numbers): slot i is writable by ticket t
    // (t % RING_CAP == i) when seq[i] == t, readable when seq[i] == t + 1, and handed on with seq[i] = t + RING_CAP
    uint32_t ring_seq[RING_CAP];
    uint32_t ring_row[RING_CAP];
    int ring_dot[RING_CAP];  // the filter's integer dot product: re-checked against the fresher threshold at dequeue time
    uint16_t ring_col[RING_CAP];
};

// Bound on acc below which a pair is certainly outside the top-k:   lead(row, query) - ty*v - tz*w - slack
//   COSINE  lead = c1 * x1,          x1 = |a|/s_a,               c1 = -thr_f/s_q
//   DOT     lead = c1 * x1,          x1 = 1/s_a,                 c1 = -thr_f/s_q
//   L2      lead = c1 * x2 * (x1 + c2),  x1 = |a|^2/2, x2 = 1/s_a,   c1 = 1/s_q, c2 = (|q|^2 - thr_f)/2
// (L2 keeps the product form: bounding |a|^2/(2 s_a) and (|q|^2 - thr_f)/(2 s_a) separately over a warp's rows
// would subtract two loosened terms of almost equal size.)
template <int METRIC>
__device__ __forceinline__ void row_figures(const float4 m, float &x1, float &x2) {
    x2 = m.w;
    if (METRIC == PKV_COSINE) x1 = m.x;
    else if (METRIC == PKV_DOT) x1 = m.w;
    else {
        const float na = m.x / m.w;  // |a|
        x1 = 0.5f * na * na;
    }
}

// [x1lo, x1hi], [x2lo, x2hi]: range of the row figures over the rows the bound must hold for (a single row: lo = hi);
// v, w, u: their largest |c|, |a - s_a c|/s_a, |a|/s_a.  Rounding slack: 3e-5 relative on every term (row and query
// figures are good to ~1e-5), qs per unit of x2 (L2: |q|^2 is a sequential f32 sum), 4e-6 |a||q|/(s_a s_q) (the
// re-scorer's own f32 summation), + 1.
template <int METRIC>
__device__ __forceinline__ float pair_bound(const float4 qc, float qs, float x1lo, float x1hi, float x2lo, float x2hi,
                                            float v, float w, float u, float *lead_out = nullptr) {
    float lead, mag;
    if (METRIC == PKV_L2) {
        const float sum = x1lo + qc.y;  // smallest |a|^2/2 + (|q|^2 - thr)/2 over the rows
        lead = qc.x * (sum >= 0.f ? x2lo : x2hi) * sum;
        mag = qc.x * x2hi * (x1hi + fabsf(qc.y));
    } else {
        lead = qc.x * (qc.x >= 0.f ? x1lo : x1hi);
        mag = fabsf(qc.x) * x1hi;
    }
    const float neg = qc.z * v + qc.w * w;
    if (lead_out) *lead_out = lead;  // acc >= lead: the APPROXIMATE score already beats the threshold
    return lead - neg - (3e-5f * (mag + neg) + qs * x2hi + 4e-6f * qc.w * u + 1.0f);
}

// The threshold-dependent query constants of pair_bound from the fixed query figures qm and the filter threshold.
template <int METRIC>
__device__ __forceinline__ void query_consts(const float4 qm, float thr, float4 &qc, float &qs) {
    if (METRIC == PKV_L2) {
        qc.x = qm.x;
        qc.y = 0.5f * (qm.w - thr);
        qs = 1e-4f * 0.5f * qm.x * (qm.w + fabsf(thr));  // |q|^2 is a sequential f32 sum: good to ~6e-5
    } else {
        qc.x = -thr * qm.x;
        qc.y = 0.f;
        qs = 0.f;
    }
    qc.z = qm.y;
    qc.w = qm.z;
}

__device__ __forceinline__ void ring_push(ImgShared *sh, uint32_t row, uint32_t col, int d) {
    const uint32_t t = atomicAdd(&sh->ring_head, 1u);
    const uint32_t i = t & RING_MASK;
    volatile uint32_t *seq = &sh->ring_seq[i];
    // ring full: wait for the re-scoring warps (they never wait for us); a wait that never ends is a bug: trap
    for (uint32_t spins = 0; *seq != t; ++spins) {
        __nanosleep(64);
        if (spins > (1u << 24)) __trap();
    }
    sh->ring_row[i] = row;
    sh->ring_dot[i] = d;
    sh->ring_col[i] = (uint16_t)col;
    __threadfence_block();
    *seq = t + 1u;
}

// Non-blocking variant: false when the ring is full.
__device__ __forceinline__ bool ring_try_push(ImgShared *sh, uint32_t row, uint32_t col, int d) {
    for (;;) {
        const uint32_t t = *(volatile uint32_t *)&sh->ring_head;
        const uint32_t i = t & RING_MASK;
        if (*(volatile uint32_t *)&sh->ring_seq[i] != t) return false;  // the slot's previous entry is still unread
        if (atomicCAS(&sh->ring_head, t, t + 1u) != t) continue;
        sh->ring_row[i] = row;
        sh->ring_dot[i] = d;
        sh->ring_col[i] = (uint16_t)col;
        __threadfence_block();
        *(volatile uint32_t *)&sh->ring_seq[i] = t + 1u;
        return true;
    }
}

constexpr uint32_t COL_CHECKED = 0x80000000u, COL_LIKELY = 0x40000000u;
#ifndef RING_PUBLISH_FENCE
#define RING_PUBLISH_FENCE() __threadfence_block()
#endif
// Warp-cooperative, non-blocking push of up to 32 survivors (one per lane in `want`) into the CTA's ring: ONE
// compare-and-swap reserves consecutive tickets for as many of them as have a free slot ahead (lane-serial pushes cost
// one shared-memory CAS round trip each - measured 10 000 clocks per 32-entry flush, with the accumulator hand-off of
// the whole CTA pair waiting behind it).  Returns the lanes that did not get in (ring full): they are parked.
__device__ __forceinline__ unsigned ring_push_warp(ImgShared *sh, unsigned want, int lane, uint32_t row, uint32_t col, int d) {
    const uint32_t lt = (1u << lane) - 1u;
    for (int attempt = 0; attempt < 16 && want; ++attempt) {
        uint32_t t = 0;
        const int leader = __ffs(want) - 1;
        if (lane == leader) t = *(volatile uint32_t *)&sh->ring_head;
        t = __shfl_sync(0xffffffffu, t, leader);
        const bool mine = (want >> lane) & 1u;
        const uint32_t rank = (uint32_t)__popc(want & lt);
        const uint32_t ticket = t + rank, i = ticket & RING_MASK;
        // the slot's previous entry must have been read (tickets are handed out in order, slots are freed in any order)
        const bool busy = mine && *(volatile uint32_t *)&sh->ring_seq[i] != ticket;
        const unsigned blocked = __ballot_sync(0xffffffffu, busy);
        // lanes ranked below the first blocked one get in
        const uint32_t m = blocked ? (uint32_t)__popc(want & ((1u << (__ffs(blocked) - 1)) - 1u)) : (uint32_t)__popc(want);
        if (m == 0) break;  // full
        uint32_t ok = 0;
        if (lane == leader) ok = atomicCAS(&sh->ring_head, t, t + m) == t ? 1u : 0u;
        ok = __shfl_sync(0xffffffffu, ok, leader);
        if (!ok) continue;  // another warp moved the head: look again
        const bool in = mine && rank < m;
        if (in) {
            sh->ring_row[i] = row;
            sh->ring_dot[i] = d;
            sh->ring_col[i] = (uint16_t)col;
            RING_PUBLISH_FENCE();
            *(volatile uint32_t *)&sh->ring_seq[i] = ticket + 1u;
        }
        want &= ~__ballot_sync(0xffffffffu, in);
        if (blocked) break;  // the ring is full behind what just got in
    }
    return want;
}


// One pre-filter survivor: exact per-pair bound, membership, then hand the row on for exact re-scoring.
// `col` bit 31 set: the pair already passed the per-pair bound in the epilogue's registers (epi_exact) and bit 30 says
// whether it is a LIKELY candidate; nothing is left to check but membership.
template <int METRIC>
__device__ __noinline__ void consider_img(const ScanArgs &a, const ImgArgs &im, int qbase, uint32_t colw, int d, uint32_t row,
                                          ImgShared *sh) {
    const int col = (int)(colw & 0xffffu);
    const int q = qbase + col;
    if (q >= a.nq || row >= a.row_end) return;
    bool likely;
    float4 m;
    const bool checked = (colw & COL_CHECKED) != 0;
    if (checked) {
        likely = (colw & COL_LIKELY) != 0;
    } else {
        m = __ldg(im.row_meta + row);
        float x1, x2;
        row_figures<METRIC>(m, x1, x2);
        float4 qc;
        float qs;
        query_consts<METRIC>(sh->qm[col], *(volatile const float *)&sh->thr[col], qc, qs);
        float lead;
        const float b = pair_bound<METRIC>(qc, qs, x1, x1, x2, x2, m.y, m.z, m.x, &lead);
        if ((float)d < b) return;  // NaN bound (non-finite row or query): kept
        likely = !((float)d < lead);  // the approximate score beats the threshold (or the pair is unfilterable)
    }
    if (!topk_member(a.topk, q, row)) return;
    if (im.fused) {
        if (!im.defer) {
            ring_push(sh, row, (uint32_t)col, d);
            return;
        }
        if (likely && ring_try_push(sh, row, (uint32_t)col, d)) return;  // the in-kernel re-scorer, if it has room
    } else if (!im.defer || likely) {
        const uint32_t slot = atomicAdd(im.pend.cnt + q, 1u);
        if (slot < im.pend.cap) im.pend.rows[(size_t)q * im.pend.cap + slot] = row;
        return;
    }
    // parked: the deferred pass re-checks with the row figures (a pre-checked pair loads them only now - the load is in
    // flight together with the list's atomic, not ahead of it)
    if (checked && im.dlist.meta) m = __ldg(im.row_meta + row);
    const uint32_t slot = atomicAdd(im.dlist.cnt + q, 1u);
    if (slot < im.dlist.cap) {
        im.dlist.rows[(size_t)q * im.dlist.cap + slot] = row;
        im.dlist.dots[(size_t)q * im.dlist.cap + slot] = d;
        if (im.dlist.meta) im.dlist.meta[(size_t)q * im.dlist.cap + slot] = m;
    }
}

// ---- exact re-scoring inside the scan kernel (fused mode) -----------------------------------------
// A re-scoring warp owns RS_SLOTS row buffers in shared memory.  It claims up to RS_SLOTS survivors from the CTA's
// ring, has the TMA engine copy their stored rows (cp.async.bulk, one mbarrier per slot: no registers are tied up by
// the gathers, RS_WARPS x RS_SLOTS rows are in flight per SM), loads the query from L2 meanwhile, and computes the
// exact score with the element order, the fmaf chains and the xor tree of rescore_kernel / the CUDA-core scan, so the
// reported score of a pair does not depend on which kernel computed it.
template <int METRIC>
__device__ __forceinline__ void acc_f4(const float4 av, const float4 qv, float &acc, float &nrm) {
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
template <int METRIC>
__device__ __forceinline__ void acc_h8(const uint4 raw8, const float4 q0v, const float4 q1v, float &acc, float &nrm) {
    const __half2 *h = reinterpret_cast<const __half2 *>(&raw8);
    float av[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 p = __half22float2(h[i]);
        av[2 * i] = p.x;
        av[2 * i + 1] = p.y;
    }
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

template <int METRIC>
__device__ __forceinline__ float exact_distance(float acc, float nrm, float qmag) {
    if (METRIC == PKV_COSINE) return cosine_key((double)acc, (double)nrm, (double)qmag);
    if (METRIC == PKV_L2) return l2_key_from_sum(acc);
    return -acc;
}

// 1-D bulk copy global -> shared, completion (bytes) signalled on an mbarrier of this CTA
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tc::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}

// Sums of (row in the shared-memory slot, query q), reduced over the warp (every lane ends up with the totals).
template <int METRIC>
__device__ __forceinline__ void score_slot(const ScanArgs &a, const ImgArgs &im, const uint8_t *slot, int q, uint64_t *bar,
                                           uint32_t phase, int lane, float &acc, float &nrm) {
    const int nvec = a.dim_pad >> 2;  // float4 per padded f32 query
    const float4 *qp = (const float4 *)a.queries + (size_t)q * nvec;
    // the query (<= 4 KB, L2 resident) is fetched while the row copy is in flight.  f32 rows: step `it` of a lane uses
    // query float4 lane + 32 it; f16 rows (8 halfs per lane per step, scan_f16_simt_kernel's order): float4 2j, 2j + 1
    // of step j = lane + 32 it
    float4 qv[8];
    if (!im.rows_f16) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = lane + 32 * it;
            qv[it] = j < nvec ? __ldg(qp + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int j = lane + 32 * it;
            const bool ok = 2 * j + 1 < nvec;
            qv[2 * it] = ok ? __ldg(qp + 2 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            qv[2 * it + 1] = ok ? __ldg(qp + 2 * j + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc::mbar_wait(bar, phase);
    acc = 0.f;
    nrm = 0.f;
    if (!im.rows_f16) {
        const float4 *rp = (const float4 *)slot;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int j = lane + 32 * it;
            if (j < nvec) acc_f4<METRIC>(rp[j], qv[it], acc, nrm);
        }
    } else {
        const uint4 *rp = (const uint4 *)slot;
        const int nvec8 = a.dim_pad >> 3;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int j = lane + 32 * it;
            if (j < nvec8) acc_h8<METRIC>(rp[j], qv[2 * it], qv[2 * it + 1], acc, nrm);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (METRIC == PKV_COSINE) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    }
}

// Claims the next ring entry if one has been handed out to a producer (it is then published within a few
// instructions), reads it and returns the slot to the producers.  false: nothing to take right now.
__device__ __forceinline__ bool ring_claim(ImgShared *sh, uint32_t &row, uint32_t &col, int &d) {
    uint32_t t;
    for (;;) {
        t = *(volatile uint32_t *)&sh->ring_tail;
        if ((int32_t)(*(volatile uint32_t *)&sh->ring_head - t) <= 0) return false;
        if (atomicCAS(&sh->ring_tail, t, t + 1u) == t) break;
    }
    const uint32_t i = t & RING_MASK;
    volatile uint32_t *seq = &sh->ring_seq[i];
    for (uint32_t spins = 0; *seq != t + 1u; ++spins)  // the producer holding this ticket is a few instructions away
        if (spins > (1u << 28)) __trap();
    __threadfence_block();
    row = *(volatile uint32_t *)&sh->ring_row[i];
    d = *(volatile int *)&sh->ring_dot[i];
    col = *(volatile uint16_t *)&sh->ring_col[i];
    __threadfence_block();
    *seq = t + RING_CAP;
    return true;
}

// Re-scoring warp: consumes ring entries until the CTA's epilogue warps are done and the ring is drained.
// Per round: up to RS_CLAIM entries are claimed (lane e holds entry e); in live mode each is checked once more
// against the threshold of NOW (entries wait in the ring while every CTA keeps tightening the thresholds - during the
// first tiles of a launch most of them have become hopeless by the time they are dequeued); the rest are gathered
// RS_SLOTS at a time.
template <int METRIC>
__device__ __forceinline__ void rescore_loop(const ScanArgs &a, const ImgArgs &im, ImgShared *sh, uint8_t *slots,
                                             uint64_t *bars, int qbase, uint32_t epi_expected, int rw, int lane) {
    const uint32_t row_bytes = (uint32_t)a.pitch_bytes;
    uint32_t phase = 0;  // bit s: parity the next completion of slot s will have
    uint32_t done = 0;
    constexpr int TPL = QM_CTA / (RS_WARPS * 32);  // thresholds this lane publishes per round
    for (;;) {
        // ---- live mode: the CTA's thresholds are published in shared memory by THESE warps (the epilogue threads read
        // them there once per tile: a global load in their per-tile path stalls the accumulator hand-off).  The loads
        // are issued now and stored behind this round's work.
        float tv[TPL];
        if (a.topk.live) {
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                const int q = qbase + rw * 32 + lane + i * RS_WARPS * 32;
                tv[i] = q < a.nq ? ld_live_f32(a.topk.thr_f + q) : 0.f;
            }
        }
        // ---- claim: lane 0 waits a little for the first entry (or the end of the CTA's scan), lanes 1.. take what is there
        uint32_t row = 0, col = 0;
        int d = 0;
        bool have = false;
        int state = 0;  // 0: nothing yet, 1: an entry, 2: the scan is over and the ring is empty
        if (lane == 0) {
            for (int spin = 0; spin < 8; ++spin) {
                if (ring_claim(sh, row, col, d)) {
                    state = 1;
                    break;
                }
                if (*(volatile uint32_t *)&sh->epi_done >= epi_expected) {
                    __threadfence_block();
                    state = ring_claim(sh, row, col, d) ? 1 : 2;  // the head is final now
                    break;
                }
                __nanosleep(128);
            }
            have = state == 1;
        }
        state = __shfl_sync(0xffffffffu, state, 0);
        if (a.topk.live) {
#pragma unroll
            for (int i = 0; i < TPL; ++i) {
                const int c = rw * 32 + lane + i * RS_WARPS * 32;
                if (qbase + c < a.nq) *(volatile float *)&sh->thr[c] = tv[i];
            }
        }
        if (state == 2) break;
        if (state == 0) continue;
        if (lane > 0 && lane < RS_CLAIM) have = ring_claim(sh, row, col, d);
        // ---- late re-check against the live threshold
        if (have && a.topk.live) {
            const int q = qbase + (int)col;
            float4 qc;
            float qs;
            query_consts<METRIC>(sh->qm[col], ld_live_f32(a.topk.thr_f + q), qc, qs);
            const float4 m = __ldg(im.row_meta + row);
            float x1, x2;
            row_figures<METRIC>(m, x1, x2);
            if ((float)d < pair_bound<METRIC>(qc, qs, x1, x1, x2, x2, m.y, m.z, m.x)) have = false;
        }
        // the CTA's scan is over: what is still queued joins the parked pairs (the launch behind this one gathers with
        // every warp of the GPU and the final thresholds) instead of keeping 146 idle SMs waiting for two warps
        // (one lane reads the flag: the decision must be the same for the whole warp)
        const uint32_t scan_over =
            __shfl_sync(0xffffffffu, (uint32_t)(*(volatile uint32_t *)&sh->epi_done >= epi_expected), 0);
        if (im.defer && scan_over) {
            if (have) {
                const int q = qbase + (int)col;
                const uint32_t slot = atomicAdd(im.dlist.cnt + q, 1u);
                if (slot < im.dlist.cap) {
                    im.dlist.rows[(size_t)q * im.dlist.cap + slot] = row;
                    im.dlist.dots[(size_t)q * im.dlist.cap + slot] = d;
                    if (im.dlist.meta) im.dlist.meta[(size_t)q * im.dlist.cap + slot] = __ldg(im.row_meta + row);
                }
            }
            continue;
        }
        const unsigned todo = __ballot_sync(0xffffffffu, have);
        if (!todo) continue;
        // ---- gather + score: the warp's RS_SLOTS row buffers form a pipeline - while entry i is scored, the TMA
        // engine is already copying entries i + 1 .. i + RS_SLOTS - 1; entry i's sums stay in the lane that holds it
        const int n = __popc(todo);
        unsigned issue_rem = todo, proc_rem = todo;
        float my_acc = 0.f, my_nrm = 0.f;
        auto issue_next = [&](int slot) {
            const int src = __ffs(issue_rem) - 1;
            issue_rem &= issue_rem - 1u;
            const uint32_t r = __shfl_sync(0xffffffffu, row, src);
            if (lane == 0) {
                // the warp has finished reading this slot (generic proxy) before the TMA engine rewrites it
                tc::fence_proxy_async();
                tc::mbar_expect_tx(&bars[slot], row_bytes);
                bulk_g2s(slots + (size_t)slot * row_bytes, (const uint8_t *)a.data + (size_t)r * (size_t)a.pitch_bytes,
                         row_bytes, &bars[slot]);
            }
        };
#pragma unroll
        for (int s0 = 0; s0 < RS_SLOTS; ++s0)
            if (s0 < n) issue_next(s0);
        for (int i = 0; i < n; ++i) {
            const int slot = i % RS_SLOTS;
            const int src = __ffs(proc_rem) - 1;
            proc_rem &= proc_rem - 1u;
            const int q = qbase + (int)__shfl_sync(0xffffffffu, col, src);
            float acc, nrm;
            score_slot<METRIC>(a, im, slots + (size_t)slot * row_bytes, q, &bars[slot], (phase >> slot) & 1u, lane, acc, nrm);
            phase ^= 1u << slot;
            if (lane == src) {
                my_acc = acc;
                my_nrm = nrm;
            }
            __syncwarp();
            if (issue_rem) issue_next(slot);
        }
        done += (uint32_t)n;
        // ---- every lane finishes its own pair: exact key, push, refresh request
        int trig_q = -1;
        if (have) {
            const int q = qbase + (int)col;
            const float dist = exact_distance<METRIC>(my_acc, my_nrm, __ldg(a.q_mag_f + q));
            if (a.topk.live) {
                if (topk_push_live(a.topk, q, row, dist)) trig_q = q;
            } else {
                topk_push(a.topk, q, row, dist);
            }
        }
        __syncwarp();
        if (a.topk.live) live_refresh_pending<16>(a.topk, trig_q, lane);
    }
    if (lane == 0 && done) atomicAdd(&a.topk.status->rescored, done);
}

// ---- the epilogue warp's hold list: pre-filter survivors wait in shared memory (this warp's own slots) until ~32 of
// them can be handled lane-parallel.
// img8_epi = 0 (round 1): divergent per-pair calls fill the list through a shared-memory atomic; the flush checks every
// entry against the exact per-pair bound (row figures from global memory) and parks / queues it - two dependent global
// round trips inside the accumulator hand-off loop.
template <int METRIC>
__device__ __noinline__ void hold_img(const ScanArgs &a, const ImgArgs &im, int qbase, int col, int d, uint32_t row,
                                      ImgShared *sh, int ew) {
    const uint32_t slot = atomicAdd(&sh->hold_cnt[ew], 1u);
    if (slot < HOLD_CAP) {
        sh->hold_row[ew][slot] = row;
        sh->hold_dot[ew][slot] = d;
        sh->hold_col[ew][slot] = (uint32_t)col;
    } else {
        consider_img<METRIC>(a, im, qbase, (uint32_t)col, d, row, sh);
    }
}

template <int METRIC>
__device__ __forceinline__ void flush_img(const ScanArgs &a, const ImgArgs &im, int qbase, ImgShared *sh, int ew, int lane,
                                          uint32_t min_cnt) {
    __syncwarp();
    const uint32_t cnt = sh->hold_cnt[ew];
    if (cnt < min_cnt) return;
    const uint32_t n = cnt < HOLD_CAP ? cnt : HOLD_CAP;
    for (uint32_t e = lane; e < n; e += 32)
        consider_img<METRIC>(a, im, qbase, sh->hold_col[ew][e], sh->hold_dot[ew][e], sh->hold_row[ew][e], sh);
    __syncwarp();
    if (lane == 0) sh->hold_cnt[ew] = 0;
    __syncwarp();
}

// img8_epi = 1 (round 2, live launches): the list is filled warp-uniformly (ballot + prefix popcount, the fill count is
// a register) with pairs that already passed the exact per-pair bound in registers; the flush has no dependent global
// load left.  A pair that must be PARKED (outside the in-kernel re-scorer's reach) only ISSUES the atomic that reserves
// its slot in the query's parked list; the two stores that need the slot are made at the next flush (or at the end of the
// scan), when the atomic has long returned: the epilogue warp never waits for a global round trip between two
// accumulators.
struct ParkPending {
    uint32_t slot, row;
    int d, q;  // q < 0: nothing pending
};

__device__ __forceinline__ void park_complete(const ImgArgs &im, ParkPending &pp) {
    if (pp.q >= 0) {
        if (pp.slot < im.dlist.cap) {
            im.dlist.rows[(size_t)pp.q * im.dlist.cap + pp.slot] = pp.row;
            im.dlist.dots[(size_t)pp.q * im.dlist.cap + pp.slot] = pp.d;
        }
        pp.q = -1;
    }
}

// One round (<= 32 entries from `base`) of a hold-list flush: membership, the ring for the likely candidates; what must
// be PARKED comes back to the caller in registers - the caller issues the atomic that reserves the slot INLINE, so that
// its result can stay pending across the tile loop (a function must wait for its outstanding loads before it returns:
// measured 10 000 clocks per flush while the atomic was issued in here).
struct HeldOut {
    uint32_t row;
    int d, q;  // q < 0: nothing to park
};

template <int METRIC>
__device__ __noinline__ HeldOut flush_round(const ScanArgs &a, const ImgArgs &im, int qbase, ImgShared *sh, int ew, int lane,
                                            uint32_t base, uint32_t n) {
    const uint32_t e = base + (uint32_t)lane;
    HeldOut out{0u, 0, -1};
    uint32_t colw = 0;
    bool park = false, likely = false;
    if (e < n) {
        colw = sh->hold_col[ew][e];
        out.row = sh->hold_row[ew][e];
        out.d = sh->hold_dot[ew][e];
        const int col = (int)(colw & 0xffffu);
        const bool parkable = (colw & COL_CHECKED) && im.fused && im.defer && im.park_lean;
        if (!parkable) {
            consider_img<METRIC>(a, im, qbase, colw, out.d, out.row, sh);  // every other mode: the synchronous path
        } else if (qbase + col < a.nq && out.row < a.row_end && topk_member(a.topk, qbase + col, out.row)) {
            out.q = qbase + col;
            park = true;
            likely = (colw & COL_LIKELY) != 0;
        }
    }
    // the likely candidates go to the in-kernel re-scorer if the ring has room; everything else is parked
    const unsigned want = __ballot_sync(0xffffffffu, park && likely);
    if (want) {
        const unsigned left = ring_push_warp(sh, want, lane, out.row, colw & 0xffffu, out.d);
        if (((want & ~left) >> lane) & 1u) out.q = -1;
    }
    return out;
}

// non-negative floats order like their bit patterns: hardware warp min/max on the unsigned view
__device__ __forceinline__ float warp_min_nn(float v) {
    return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));
}
__device__ __forceinline__ float warp_max_nn(float v) {
    return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(v)));
}

// TN = rows per tile (MMA N), NBUF = accumulator buffers in TMEM (three hide the commit -> tcgen05.ld -> arrive ->
// re-issue round trip of a buffer, see pkv_scan_ts.cu).
template <int METRIC, int CPS, bool PAIR, int TN, int NBUF>
__global__ void __launch_bounds__(IMG_THREADS, 1)
scan_img8_kernel(const __grid_constant__ CUtensorMap tmap_rows, const ScanArgs a, const ImgArgs im, const int q0,
                 const int groups, const int kchunks, const int stages, const int rs_bytes) {
    constexpr int TILE_N = TN;
    constexpr int EPI_USED = (TN / 32) * 4;            // epilogue warps with work: 32 accumulator columns each
    const uint32_t acc_col0 = (uint32_t)(im.dim_pad8 / 4 + 31) / 32 * 32;  // accumulators behind the query columns
    constexpr int ROWS_CTA = PAIR ? TN / 2 : TN;       // rows this CTA stages per tile
    constexpr int BOX_BYTES = ROWS_CTA * CHUNK_BYTES;  // one TMA box
    constexpr int STAGE_BYTES = CPS * BOX_BYTES;
    constexpr int NCTA = PAIR ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_addr = tc::smem_u32(smem_raw);
    uint8_t *smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
    uint8_t *s_b = smem;  // [stages][CPS chunks][ROWS_CTA rows][128 B]
    uint8_t *rs_slots = s_b + (size_t)stages * STAGE_BYTES;  // [RS_WARPS][RS_SLOTS][row pitch] (fused mode)
    ImgShared *sh = reinterpret_cast<ImgShared *>(rs_slots + rs_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? tc::cluster_ctarank() : 0u;
    const uint32_t unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x, nunits = PAIR ? (gridDim.x >> 1) : gridDim.x;
    const uint32_t grp = unit % (uint32_t)groups, seq = unit / (uint32_t)groups, nseq = nunits / (uint32_t)groups;
    const int qbase = q0 + (int)grp * (NCTA * QM_CTA) + (int)rank * QM_CTA;  // first query of this CTA
    const uint32_t nrows = a.row_end - a.row_begin;
    const uint32_t ntiles = (nrows + TILE_N - 1) / TILE_N;
    const float INF = __int_as_float(0x7f800000);

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            tc::mbar_init(&sh->full[s], 1);
            tc::mbar_init(&sh->empty[s], 1);
        }
        for (int b = 0; b < NBUF; ++b) {
            tc::mbar_init(&sh->tmem_full[b], 1);
            tc::mbar_init(&sh->tmem_empty[b], NCTA * EPI_USED);
        }
        for (int w = 0; w < RS_WARPS; ++w)
            for (int b = 0; b < RS_SLOTS; ++b) tc::mbar_init(&sh->rs_bar[w][b], 1);
        for (int w = 0; w < EPI_WARPS; ++w) sh->hold_cnt[w] = 0;
        sh->ring_head = 0;
        sh->ring_tail = 0;
        sh->epi_done = 0;
        tc::fence_barrier_init();
        tc::prefetch_tmap(&tmap_rows);
    }
    for (uint32_t i = threadIdx.x; i < RING_CAP; i += IMG_THREADS) sh->ring_seq[i] = i;
    if (warp == 1) {
        if (PAIR) {
            tc::tmem_alloc_cta2(&sh->tmem_base, TMEM_COLS);
            tc::tmem_relinquish_cta2();
        } else {
            tc::tmem_alloc(&sh->tmem_base, TMEM_COLS);
            tc::tmem_relinquish();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = sh->tmem_base;

    float4 qm = make_float4(0.f, 0.f, 0.f, 0.f);  // this thread's query (epilogue thread = one query)
    float thr0 = 0.f;
    bool real_q = false;  // padded query lanes keep nothing
    if (warp >= 2 && warp < RS_WARP0) {
        const int ew = warp - 2, quarter = warp & 3;
        const int col = quarter * 32 + lane;
        const int q = qbase + col;
        if (q < a.nq) {
            real_q = true;
            // +inf while the query has no threshold: keep everything
            thr0 = a.topk.live ? ld_live_f32(a.topk.thr_f + q) : __ldg(a.topk.thr_f + q);
            qm = __ldg(im.q_meta + q);
        }
        if ((ew >> 2) == 0) {
            sh->qm[col] = qm;
            sh->thr[col] = thr0;
        }
        // this CTA's 128 queries -> TMEM columns [0, dim_pad8/4): lane = query, 4 codes per column
        const int qrow = q < a.nq ? q : (a.nq - 1);
        const uint8_t *qp = (const uint8_t *)im.q8 + (size_t)qrow * im.dim_pad8;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // all global loads first (<= 8 chunks of 32 codes per warp for D <= 1024), then the TMEM stores: the
        // prologue pays one memory latency instead of one per chunk
        const int n8 = im.dim_pad8 / 32;
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
    if (PAIR) tc::cluster_sync();
    tc::fence_after_sync();

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (seq < nseq) {
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0;
            const uint32_t full0 = PAIR ? tc::mapa(tc::smem_u32(&sh->full[0]), 0) : tc::smem_u32(&sh->full[0]);
            for (uint32_t tile = seq; tile < ntiles; tile += nseq) {
                const int row0 = (int)(a.row_begin + tile * TILE_N + rank * ROWS_CTA);
                for (int kc = 0; kc < kchunks; kc += CPS) {
                    const int n = kchunks - kc < CPS ? kchunks - kc : CPS;
                    tc::mbar_wait(&sh->empty[s], ph ^ 1);
                    if (issuer) {
                        if (rank == 0) tc::mbar_expect_tx(&sh->full[s], (uint32_t)(NCTA * n * BOX_BYTES));
#pragma unroll
                        for (int j = 0; j < CPS; ++j) {
                            if (j < n) {
                                uint8_t *dst = s_b + (size_t)s * STAGE_BYTES + j * BOX_BYTES;
                                if (PAIR)
                                    tc::tma_load_2d_cta2(dst, &tmap_rows, full0 + s * 8u, (kc + j) * CHUNK_BYTES, row0);
                                else
                                    tc::tma_load_2d(dst, &tmap_rows, &sh->full[s], (kc + j) * CHUNK_BYTES, row0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (whole warp, one elected lane issues) =====================
        if (rank == 0 && seq < nseq) {
            constexpr uint32_t idesc = tc::make_idesc(/*S32*/ 2, /*INT8*/ 1, NCTA * QM_CTA, TILE_N);
            const bool issuer = tc::elect_one();
            uint32_t s = 0, ph = 0, buf = 0, bph = 0;
#ifdef PKV_TILE_TIMING
            long long tt_empty = 0, tt_full = 0, tt_issue = 0, tt_n = 0;
#endif
            for (uint32_t tile = seq; tile < ntiles; tile += nseq) {
#ifdef PKV_TILE_TIMING
                const long long tt0 = clock64();
#endif
                tc::mbar_wait(&sh->tmem_empty[buf], bph ^ 1);
                tc::fence_after_sync();
#ifdef PKV_TILE_TIMING
                const long long tt1 = clock64();
                tt_empty += tt1 - tt0;
                ++tt_n;
#endif
                const uint32_t d_tmem = tmem_base + acc_col0 + buf * TILE_N;
                for (int kc = 0; kc < kchunks; kc += CPS) {
                    const int n = kchunks - kc < CPS ? kchunks - kc : CPS;
#ifdef PKV_TILE_TIMING
                    const long long tf0 = clock64();
#endif
                    tc::mbar_wait(&sh->full[s], ph);
                    tc::fence_after_sync();
#ifdef PKV_TILE_TIMING
                    tt_full += clock64() - tf0;
#endif
                    const uint64_t b_desc = tc::smem_desc_sw128(tc::smem_u32(s_b) + s * STAGE_BYTES);
                    const uint32_t a_tmem = tmem_base + (uint32_t)kc * (CHUNK_BYTES / 4);
                    if (issuer) {
#pragma unroll
                        for (int j = 0; j < CPS; ++j) {
                            if (j < n) {
#pragma unroll
                                for (int k = 0; k < CHUNK_BYTES / 32; ++k) {
                                    const uint32_t at = a_tmem + j * (CHUNK_BYTES / 4) + k * 8;
                                    const uint64_t bd = b_desc + (uint64_t)(j * (BOX_BYTES / 16) + k * 2);
                                    if (PAIR) tc::mma_i8_ts_cta2(d_tmem, at, bd, idesc, (kc | j | k) != 0);
                                    else tc::mma_i8_ts(d_tmem, at, bd, idesc, (kc | j | k) != 0);
                                }
                            }
                        }
                        if (PAIR) tc::mma_commit_cta2(&sh->empty[s]);
                        else tc::mma_commit(&sh->empty[s]);
                    }
                    __syncwarp();
                    if (++s == (uint32_t)stages) { s = 0; ph ^= 1; }
                }
                if (issuer) {
                    if (PAIR) tc::mma_commit_cta2(&sh->tmem_full[buf]);
                    else tc::mma_commit(&sh->tmem_full[buf]);
                }
                __syncwarp();
                if (++buf == NBUF) { buf = 0; bph ^= 1; }
#ifdef PKV_TILE_TIMING
                tt_issue += clock64() - tt1;
#endif
            }
#ifdef PKV_TILE_TIMING
            if (im.dbg && lane == 0) {
                unsigned long long *o = im.dbg + ((size_t)blockIdx.x * 20 + warp) * 8;
                o[0] = (unsigned long long)tt_n; o[1] = (unsigned long long)tt_empty; o[2] = (unsigned long long)tt_full;
                o[3] = (unsigned long long)tt_issue;
            }
#endif
        }
    } else if (seq < nseq && warp - 2 < EPI_USED) {
        // ===================== epilogue: own 128 queries x the tile's rows =====================
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col0 = (ew >> 2) * 32;       // accumulator columns = rows of the tile
        const int qcol = quarter * 32 + lane;  // this thread's query within the CTA
        uint32_t buf = 0, bph = 0;
        uint32_t hold_n = 0;  // entries in this warp's hold list (warp-uniform; img8_epi = 1)
        ParkPending pp{0u, 0u, 0, -1};
#ifdef PKV_TILE_TIMING
        long long te_wait = 0, te_ld = 0, te_proc = 0, te_n = 0, te_max = 0, te_slow = 0, te_pre = 0, te_pre_slow = 0;
        long long te_flush = 0, te_nflush = 0, te_fmax = 0, tf_round = 0, tf_park = 0, tf_atom = 0, te_prev = 0;
#endif
        // Flushes the first 32 held entries (one lane-parallel round: a second round would have to wait for the atomics
        // the first one has just issued) and moves the rest to the front of the list; `all`: the end of the scan.
        auto flush_held = [&](bool all) {
#ifdef PKV_TILE_TIMING
            const long long tf = clock64();
#endif
            __syncwarp();
            for (uint32_t base = 0; base < hold_n; base += 32) {
#ifdef PKV_TILE_TIMING
                const long long f0 = clock64();
#endif
                const HeldOut h = flush_round<METRIC>(a, im, qbase, sh, ew, lane, base, hold_n < base + 32 ? hold_n : base + 32);
#ifdef PKV_TILE_TIMING
                const long long f1 = clock64();
#endif
                park_complete(im, pp);
#ifdef PKV_TILE_TIMING
                const long long f2 = clock64();
#endif
                if (h.q >= 0) {
                    pp.slot = atomicAdd(im.dlist.cnt + h.q, 1u);  // issued now, consumed at the next flush
                    pp.row = h.row;
                    pp.d = h.d;
                    pp.q = h.q;
                }
#ifdef PKV_TILE_TIMING
                tf_round += f1 - f0;
                tf_park += f2 - f1;
                tf_atom += clock64() - f2;
#endif
                if (!all) break;
            }
            __syncwarp();
#ifdef PKV_TILE_TIMING
            if (!all) {
                const long long dtf = clock64() - tf;
                te_flush += dtf;
                ++te_nflush;
                if (dtf > te_fmax) te_fmax = dtf;
            }
#endif
            if (!all && hold_n > 32) {
                const uint32_t rest = hold_n - 32;
                uint32_t r0 = 0, r2 = 0;
                int r1 = 0;
                if ((uint32_t)lane < rest) {
                    r0 = sh->hold_row[ew][32 + lane];
                    r1 = sh->hold_dot[ew][32 + lane];
                    r2 = sh->hold_col[ew][32 + lane];
                }
                __syncwarp();
                if ((uint32_t)lane < rest) {
                    sh->hold_row[ew][lane] = r0;
                    sh->hold_dot[ew][lane] = r1;
                    sh->hold_col[ew][lane] = r2;
                }
                __syncwarp();
                hold_n = rest;
            } else {
                hold_n = 0;
            }
        };
        const uint32_t lt_mask = (1u << lane) - 1u;
        // lane j prefetches the figures of row col0 + j (the warp's 32 rows of the tile)
        uint32_t nrow = a.row_begin + seq * TILE_N + col0 + lane;
        bool row_ok = seq < ntiles && nrow < a.row_end;
        float4 m = row_ok ? __ldg(im.row_meta + nrow) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t empty0 = PAIR ? tc::mapa(tc::smem_u32(&sh->tmem_empty[0]), 0) : tc::smem_u32(&sh->tmem_empty[0]);
        // live mode: the query's threshold is re-read once per tile from shared memory, where this CTA's re-scoring
        // warps keep publishing what every CTA's re-scoring warps tighten in global memory while the scan runs
        const bool live = a.topk.live != 0 && real_q;
        float4 qc;
        float qs;
        query_consts<METRIC>(qm, thr0, qc, qs);
#ifdef PKV_TILE_TIMING
        te_prev = clock64();
#endif
        for (uint32_t tile = seq; tile < ntiles; tile += nseq) {
#ifdef PKV_TILE_TIMING
            const long long t_top = clock64();
#endif
            const uint32_t row_first = a.row_begin + tile * TILE_N + col0;
            if (live) query_consts<METRIC>(qm, *(volatile const float *)&sh->thr[qcol], qc, qs);
            // loosest figures over the warp's 32 rows (rows past the end must not loosen them)
            float x1, x2;
            row_figures<METRIC>(m, x1, x2);
            const float x1lo = warp_min_nn(row_ok ? x1 : INF), x1hi = warp_max_nn(row_ok ? x1 : 0.f);
            float x2lo = 0.f, x2hi = 0.f;
            if (METRIC == PKV_L2) {
                x2lo = warp_min_nn(row_ok ? x2 : INF);
                x2hi = warp_max_nn(row_ok ? x2 : 0.f);
            }
            const float vhi = warp_max_nn(m.y), whi = warp_max_nn(m.z), uhi = warp_max_nn(m.x);
            const float4 mc = m;  // figures of row col0 + lane of THIS tile (exact per-pair bounds below)
            nrow = a.row_begin + (tile + nseq) * TILE_N + col0 + lane;
            row_ok = tile + nseq < ntiles && nrow < a.row_end;
            m = row_ok ? __ldg(im.row_meta + nrow) : make_float4(0.f, 0.f, 0.f, 0.f);
            {
                // ... and the figures of the tile after next are pulled into L2 now: the load above then never waits for
                // DRAM - with 32 epilogue warps per accumulator, ONE slow DRAM access per tile stalls the whole hand-off
                const uint32_t prow = a.row_begin + (tile + 3 * nseq) * TILE_N + col0 + lane;
                if (tile + 3 * nseq < ntiles && prow < a.row_end && (lane & 7) == 0)  // one request per 128-byte line
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(im.row_meta + prow));
            }
            float bf = pair_bound<METRIC>(qc, qs, x1lo, x1hi, x2lo, x2hi, vhi, whi, uhi);
            bf = fminf(fmaxf(bf, -BOUND_CLAMP), BOUND_CLAMP);  // NaN -> -clamp: keep everything
            const int bound = real_q ? __float2int_rd(bf) : (int)BOUND_CLAMP;
#ifdef PKV_TILE_TIMING
            const long long ta = clock64();   // bound of this tile computed: (ta - te_prev) = processing of the previous tile + this bound
            te_pre += ta - t_top;
            if (ta - t_top > 1500) ++te_pre_slow;
            {
                const long long p = ta - te_prev;
                te_proc += p;
                if (p > te_max) te_max = p;
                if (p > 1500) ++te_slow;
            }
#endif
            tc::mbar_wait(&sh->tmem_full[buf], bph);
            tc::fence_after_sync();
#ifdef PKV_TILE_TIMING
            const long long tb = clock64();
            te_wait += tb - ta;
#endif
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc_col0 + buf * TILE_N + col0, v);
            tc::tmem_ld_wait();
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) tc::mbar_arrive_cluster(empty0 + buf * 8u);  // accumulator is in registers
                else tc::mbar_arrive(&sh->tmem_empty[buf]);
            }
#ifdef PKV_TILE_TIMING
            te_prev = clock64();
            te_ld += te_prev - tb;
            ++te_n;
#endif
            if (++buf == NBUF) { buf = 0; bph ^= 1; }
            if (!im.epi_exact) {
                // round 1: sign bit of (bound - 1 - d) is set iff d >= bound: OR them all, branch once per lane
                int any = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) any |= bound - (int)v[j] - 1;
                if (any < 0) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int d = (int)v[j];
                        if (d >= bound) hold_img<METRIC>(a, im, qbase, qcol, d, row_first + j, sh, ew);
                    }
                }
                flush_img<METRIC>(a, im, qbase, sh, ew, lane, HOLD_FLUSH);
                continue;
            }
            // round 2.  The largest of the lane's 32 dot products against the bound: maxima of 8 groups of 4 rows first
            // (3-input max: 20 instructions in all), one branch for the whole warp
            int g[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                g[i] = max(__vimax3_s32((int)v[4 * i], (int)v[4 * i + 1], (int)v[4 * i + 2]), (int)v[4 * i + 3]);
            const int vmax = __vimax3_s32(__vimax3_s32(g[0], g[1], g[2]), __vimax3_s32(g[3], g[4], g[5]), max(g[6], g[7]));
            if (__any_sync(0xffffffffu, vmax >= bound)) {
                // which groups hold a survivor of ANY lane (one warp-wide OR): only those are looked at row by row -
                // straight-line, predicated: each lane (query) notes how many of its rows pass and which was the last
                uint32_t gm = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) gm |= g[i] >= bound ? (1u << i) : 0u;
                const uint32_t gw = __reduce_or_sync(0xffffffffu, gm);
                int n = 0, ej = 0, ed = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (gw & (1u << i)) {
#pragma unroll
                        for (int j = 4 * i; j < 4 * i + 4; ++j) {
                            if ((int)v[j] >= bound) {
                                ej = j;
                                ed = (int)v[j];
                                ++n;
                            }
                        }
                    }
                }
                int fj = ej, fd = ed;
                const int nmax = __reduce_max_sync(0xffffffffu, n);
                if (nmax == 2) {  // some lane has two: the FIRST of each lane as well
#pragma unroll
                    for (int i = 7; i >= 0; --i) {
                        if (gw & (1u << i)) {
#pragma unroll
                            for (int j = 4 * i + 3; j >= 4 * i; --j) {
                                if ((int)v[j] >= bound) {
                                    fj = j;
                                    fd = (int)v[j];
                                }
                            }
                        }
                    }
                }
                if (nmax <= 2) {
                    // <= 2 rounds: bound THIS pair exactly - the figures of row j sit in lane j (mc) - and append what
                    // is left, warp-uniformly (ballot + prefix popcount: no atomics, no divergent calls)
                    for (int round = 0; round < nmax; ++round) {
                        const int rj = round == 0 ? ej : fj, rd = round == 0 ? ed : fd;
                        bool pass = n > round;
                        float4 mj;
                        mj.x = __shfl_sync(0xffffffffu, mc.x, rj);
                        mj.y = __shfl_sync(0xffffffffu, mc.y, rj);
                        mj.z = __shfl_sync(0xffffffffu, mc.z, rj);
                        mj.w = __shfl_sync(0xffffffffu, mc.w, rj);
                        uint32_t colw = (uint32_t)qcol | COL_CHECKED;
                        if (pass) {
                            float x1, x2, lead;
                            row_figures<METRIC>(mj, x1, x2);
                            const float b = pair_bound<METRIC>(qc, qs, x1, x1, x2, x2, mj.y, mj.z, mj.x, &lead);
                            if ((float)rd < b) pass = false;  // NaN bound (non-finite row or query): kept
                            if (!((float)rd < lead)) colw |= COL_LIKELY;
                        }
                        const unsigned mask = __ballot_sync(0xffffffffu, pass);
                        if (pass) {
                            const uint32_t slot = hold_n + (uint32_t)__popc(mask & lt_mask);
                            sh->hold_row[ew][slot] = row_first + (uint32_t)rj;
                            sh->hold_dot[ew][slot] = rd;
                            sh->hold_col[ew][slot] = colw;
                        }
                        hold_n += (uint32_t)__popc(mask);
                        if (hold_n > (uint32_t)(HOLD_CAP - 32)) flush_held(false);  // the next round may add 32 more
                    }
                } else {
                    // three or more rows of one query in one 32-row slice (rare in a live launch, whose thresholds are
                    // tight; the dense chunks of a learning prefix run the round-1 epilogue): pair by pair
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if ((int)v[j] >= bound) consider_img<METRIC>(a, im, qbase, (uint32_t)qcol, (int)v[j], row_first + j, sh);
                }
            }
            if (hold_n >= (uint32_t)HOLD_FLUSH) flush_held(false);
        }
        if (im.epi_exact) {
            if (hold_n) flush_held(true);
            park_complete(im, pp);
        } else {
            flush_img<METRIC>(a, im, qbase, sh, ew, lane, 1);
        }
#ifdef PKV_TILE_TIMING
        if (im.dbg && lane == 0) {
            unsigned long long *o = im.dbg + ((size_t)blockIdx.x * 20 + warp) * 8;
            o[0] = (unsigned long long)te_n; o[1] = (unsigned long long)te_wait; o[2] = (unsigned long long)te_ld;
            o[3] = (unsigned long long)te_proc; o[4] = (unsigned long long)te_max; o[5] = (unsigned long long)te_slow;
            o[6] = (unsigned long long)te_pre; o[7] = (unsigned long long)te_pre_slow;
            unsigned long long *o2 = im.dbg + (size_t)256 * 20 * 8 + ((size_t)blockIdx.x * 20 + warp) * 4;
            o2[0] = (unsigned long long)te_flush; o2[1] = (unsigned long long)te_nflush; o2[2] = (unsigned long long)te_fmax;
            o2[3] = (unsigned long long)tf_round;
            unsigned long long *o3 = im.dbg + (size_t)256 * 20 * 12 + ((size_t)blockIdx.x * 20 + warp) * 2;
            o3[0] = (unsigned long long)tf_park; o3[1] = (unsigned long long)tf_atom;
        }
#endif
        __syncwarp();
        if (lane == 0) {  // this warp has published its last survivor
            __threadfence_block();
            atomicAdd(&sh->epi_done, 1u);
        }
    } else if (warp >= RS_WARP0) {
        // ===================== exact re-scoring of this CTA's filter survivors (fused mode) =====================
        if (im.fused) {
            const int rw = warp - RS_WARP0;
            rescore_loop<METRIC>(a, im, sh, rs_slots + (size_t)rw * RS_SLOTS * (size_t)a.pitch_bytes, sh->rs_bar[rw], qbase,
                                 seq < nseq ? (uint32_t)EPI_USED : 0u, rw, lane);
        }
    }

    tc::fence_before_sync();
    __syncthreads();
    if (PAIR) tc::cluster_sync();
    if (warp == 1) {
        if (PAIR) tc::tmem_dealloc_cta2(tmem_base, TMEM_COLS);
        else tc::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// Behind a live launch: the pairs it parked (inside the filter's error band, or more than the in-kernel re-scorer
// could take) are checked once more - against the threshold the WHOLE scan arrived at - and only what still cannot be
// excluded is gathered and scored exactly.  One CTA per (query, slice); the check runs lane-parallel, the gathers one
// warp per pair with the summation order of rescore_kernel.
template <int METRIC, bool ROWS_F16>
__global__ void __launch_bounds__(256) rescore_deferred_kernel(const ScanArgs a, const ImgArgs im, SearchStatus *status) {
    extern __shared__ float4 s_q[];  // one query, dim_pad/4 float4 (f32, widened for an f16 index)
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31;
    const int nvec = a.dim_pad >> 2;
    const uint32_t raw = im.dlist.cnt[q];
    const uint32_t n = raw < im.dlist.cap ? raw : im.dlist.cap;
    if (blockIdx.y == 0 && threadIdx.x == 0 && raw > im.dlist.cap) atomicOr(&status->defer_overflow, 1u);
    if (n == 0) return;
    const float4 *gq = (const float4 *)a.queries + (size_t)q * nvec;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) s_q[i] = gq[i];
    __syncthreads();
    float4 qc;
    float qs;
    query_consts<METRIC>(__ldg(im.q_meta + q), a.topk.thr_f[q], qc, qs);
    const float qmag = __ldg(a.q_mag_f + q);
    const uint32_t nwarps = gridDim.y * (blockDim.x >> 5);
    const uint32_t wid = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t gathered = 0;
    for (uint32_t base = wid * 32u; base < n; base += nwarps * 32u) {
        const uint32_t e = base + lane;
        uint32_t row = 0;
        bool keep = false;
        if (e < n) {
            row = im.dlist.rows[(size_t)q * im.dlist.cap + e];
            const int d = im.dlist.dots[(size_t)q * im.dlist.cap + e];
            const float4 m = im.dlist.meta ? im.dlist.meta[(size_t)q * im.dlist.cap + e] : __ldg(im.row_meta + row);
            float x1, x2;
            row_figures<METRIC>(m, x1, x2);
            keep = !((float)d < pair_bound<METRIC>(qc, qs, x1, x1, x2, x2, m.y, m.z, m.x));
        }
        unsigned todo = __ballot_sync(0xffffffffu, keep);
        gathered += (uint32_t)__popc(todo);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const uint32_t r = __shfl_sync(0xffffffffu, row, src);
            const uint8_t *rb = (const uint8_t *)a.data + (size_t)r * (size_t)a.pitch_bytes;
            float acc = 0.f, nrm = 0.f;
            // all of the row's loads are issued before the first is consumed (one memory latency per gathered row instead
            // of one per 512 bytes); the element order of the sums is unchanged
            if (!ROWS_F16) {
                const float4 *rp = (const float4 *)rb;
                float4 rv[8];
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    if (lane + it * 32 < nvec) rv[it] = __ldg(rp + lane + it * 32);
#pragma unroll
                for (int it = 0; it < 8; ++it)
                    if (lane + it * 32 < nvec) acc_f4<METRIC>(rv[it], s_q[lane + it * 32], acc, nrm);
                for (int j = lane + 256; j < nvec; j += 32) acc_f4<METRIC>(__ldg(rp + j), s_q[j], acc, nrm);
            } else {
                const uint4 *rp = (const uint4 *)rb;
                const int nvec8 = a.dim_pad >> 3;
                uint4 rv[4];
#pragma unroll
                for (int it = 0; it < 4; ++it)
                    if (lane + it * 32 < nvec8) rv[it] = __ldg(rp + lane + it * 32);
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int j = lane + it * 32;
                    if (j < nvec8) acc_h8<METRIC>(rv[it], s_q[2 * j], s_q[2 * j + 1], acc, nrm);
                }
                for (int j = lane + 128; j < nvec8; j += 32) acc_h8<METRIC>(__ldg(rp + j), s_q[2 * j], s_q[2 * j + 1], acc, nrm);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (METRIC == PKV_COSINE) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
            }
            if (lane == 0) topk_push(a.topk, q, r, exact_distance<METRIC>(acc, nrm, qmag));
        }
    }
    if (lane == 0 && gathered) atomicAdd(&status->deferred, gathered);
}

// ---------------------------------------------------------------------------------------------
// Image of rows [row_begin, row_end): one warp per row.
template <bool ROWS_F16>
__global__ void __launch_bounds__(256) img8_build_kernel(const uint8_t *data, int64_t pitch, int dim, int dim_pad8,
                                                         int64_t row_begin, int64_t row_end, float U, int8_t *img,
                                                         float4 *meta) {
    const int lane = threadIdx.x & 31;
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_end) return;
    const uint8_t *src = data + (size_t)row * (size_t)pitch;
    auto load = [&](int i) -> float {
        return ROWS_F16 ? __half2float(reinterpret_cast<const __half *>(src)[i]) : reinterpret_cast<const float *>(src)[i];
    };
    float amax = 0.f, nrm2 = 0.f;
    bool bad = false;
    for (int i = lane; i < dim; i += 32) {
        const float x = load(i);
        const float ax = fabsf(x);
        if (!(ax <= 3.0e38f)) bad = true;  // NaN or inf
        amax = fmaxf(amax, ax);
        nrm2 = fmaf(x, x, nrm2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
    }
    bad = __any_sync(0xffffffffu, bad) || !(nrm2 <= 3.0e38f);
    float s = sqrtf(nrm2) / U;
    if (!(s > 0.f) || bad) s = 1.0f;  // zero row (codes 0, no error) or non-finite row (flagged below)
    const float r = 1.0f / s;
    if (!(r <= 3.0e38f)) bad = true;  // denormal scale: treat as unfilterable
    int8_t *dst = img + (size_t)row * dim_pad8;
    float err2 = 0.f;
    int cn2 = 0;
    for (int i0 = lane * 4; i0 < dim_pad8; i0 += 128) {
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            int c = 0;
            if (i < dim && !bad) {
                const float x = load(i);
                float t = rintf(__fdiv_rn(x, s));
                t = fminf(fmaxf(t, -127.f), 127.f);
                c = (int)t;
                const float ev = fmaf(-s, t, x);
                err2 = fmaf(ev, ev, err2);
                cn2 += c * c;
            }
            packed |= ((uint32_t)(uint8_t)(int8_t)c) << (8 * e);
        }
        *reinterpret_cast<uint32_t *>(dst + i0) = packed;
    }
    for (int o = 16; o > 0; o >>= 1) {
        err2 += __shfl_xor_sync(0xffffffffu, err2, o);
        cn2 += __shfl_xor_sync(0xffffffffu, cn2, o);
    }
    if (lane == 0) {
        float4 m;
        if (bad) {
            m = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 1.0f);  // w = +inf: the bound is -inf/NaN, always re-scored
        } else {
            m.x = sqrtf(nrm2) * r;
            m.y = sqrtf((float)cn2);
            m.z = sqrtf(err2) * 1.0001f * r;
            m.w = r;
            if (!(m.x <= 3.0e38f) || !(m.z <= 3.0e38f)) m = make_float4(0.f, 0.f, __int_as_float(0x7f800000), 1.0f);
        }
        meta[row] = m;
    }
}

// Peakiness statistics of rows [row_begin, row_end): sum, sum of squares and count of max|a_i| / |a| over the
// finite non-zero rows (one warp per row).
template <bool ROWS_F16>
__global__ void __launch_bounds__(256) img8_stats_kernel(const uint8_t *data, int64_t pitch, int dim, int64_t row_begin,
                                                         int64_t row_end, double *stats) {
    const int lane = threadIdx.x & 31;
    const int64_t row = row_begin + (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= row_end) return;
    const uint8_t *src = data + (size_t)row * (size_t)pitch;
    float amax = 0.f, nrm2 = 0.f;
    for (int i = lane; i < dim; i += 32) {
        const float x = ROWS_F16 ? __half2float(reinterpret_cast<const __half *>(src)[i]) : reinterpret_cast<const float *>(src)[i];
        amax = fmaxf(amax, fabsf(x));
        nrm2 = fmaf(x, x, nrm2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
    }
    if (lane == 0 && nrm2 > 0.f && nrm2 <= 3.0e38f && amax <= 3.0e38f) {
        const double p = (double)amax / sqrt((double)nrm2);
        atomicAdd(stats + 0, p);
        atomicAdd(stats + 1, p * p);
        atomicAdd(stats + 2, 1.0);
    }
}

// Query codes and figures: one warp per query, from the padded f32 queries of the workspace.
__global__ void __launch_bounds__(256) img8_prep_queries_kernel(const float *q, int nq, int dim, int dim_pad, int dim_pad8,
                                                                const float *q_mag_f, int8_t *q8, float4 *q_meta) {
    const int lane = threadIdx.x & 31;
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= nq) return;
    const float *src = q + (size_t)qi * dim_pad;
    float amax = 0.f, nrm2 = 0.f;
    bool bad = false;
    for (int i = lane; i < dim; i += 32) {
        const float x = src[i];
        const float ax = fabsf(x);
        if (!(ax <= 3.0e38f)) bad = true;
        amax = fmaxf(amax, ax);
        nrm2 = fmaf(x, x, nrm2);
    }
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
    }
    bad = __any_sync(0xffffffffu, bad) || !(nrm2 <= 3.0e38f);
    float s = amax / 127.0f;
    if (!(s > 0.f) || bad) s = 1.0f;
    const float r = 1.0f / s;
    if (!(r <= 3.0e38f)) bad = true;
    int8_t *dst = q8 + (size_t)qi * dim_pad8;
    float err2 = 0.f;
    for (int i0 = lane * 4; i0 < dim_pad8; i0 += 128) {
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            int c = 0;
            if (i < dim && !bad) {
                const float x = src[i];
                float t = rintf(__fdiv_rn(x, s));
                t = fminf(fmaxf(t, -127.f), 127.f);
                c = (int)t;
                const float ev = fmaf(-s, t, x);
                err2 = fmaf(ev, ev, err2);
            }
            packed |= ((uint32_t)(uint8_t)(int8_t)c) << (8 * e);
        }
        *reinterpret_cast<uint32_t *>(dst + i0) = packed;
    }
    for (int o = 16; o > 0; o >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, o);
    if (lane == 0) {
        const float INF = __int_as_float(0x7f800000);
        float4 m;
        m.x = r;
        m.y = bad ? INF : sqrtf(err2) * 1.0001f * r;
        m.z = bad ? INF : sqrtf(nrm2) * 1.00001f * r;
        m.w = q_mag_f[qi];  // the |q|^2 the thresholds and the re-scorer use
        q_meta[qi] = m;
    }
}

template <int METRIC, int CPS, bool PAIR, int TN, int NBUF>
int launch_img8(const Index &ix, const ScanArgs &a, const ImgArgs &im, int q0, int groups, int kchunks, cudaStream_t s) {
    constexpr int ROWS_CTA = PAIR ? TN / 2 : TN;
    constexpr int STAGE_BYTES = CPS * ROWS_CTA * CHUNK_BYTES;
    CUtensorMap mrows;
    PKV_TRY(make_tmap_bytes(&mrows, ix.d_img8, (uint64_t)ix.dim_pad8, (uint64_t)ix.sealed_rows, (uint64_t)ix.dim_pad8, ROWS_CTA));
    const size_t ctrl = sizeof(ImgShared);
    const size_t rs_bytes = im.fused ? (size_t)RS_WARPS * RS_SLOTS * (size_t)ix.pitch : 0;  // the re-scoring warps' row buffers
    int stages = (int)((227 * 1024 - 1024 - ctrl - rs_bytes) / STAGE_BYTES);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (ix.opt.ts_stages > 1 && ix.opt.ts_stages < stages) stages = ix.opt.ts_stages;
    if (stages < 2) return fail(PKV_ERR_UNSUPPORTED, "dim %d leaves no room for the row stages", ix.dim);
    const size_t smem = 1024 + (size_t)stages * STAGE_BYTES + rs_bytes + ctrl;
    auto kernel = scan_img8_kernel<METRIC, CPS, PAIR, TN, NBUF>;
    PKV_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t ntiles = (a.row_end - a.row_begin + TN - 1) / TN;
    uint32_t units = PAIR ? (uint32_t)ix.sm_count / 2 : (uint32_t)ix.sm_count;
    units = units / groups * groups;
    const uint32_t want = ntiles * (uint32_t)groups;
    if (units > want) units = want;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PAIR ? 2 * units : units);
    cfg.blockDim = dim3(IMG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PKV_CUDA(cudaLaunchKernelEx(&cfg, kernel, mrows, a, im, q0, groups, kchunks, stages, (int)rs_bytes));
    return PKV_OK;
}

// Tile shape by what TMEM has left behind the query columns (see pkv_scan_ts.cu)
template <int METRIC, int CPS, bool PAIR>
int launch_img8_shape(const Index &ix, const ScanArgs &a, const ImgArgs &im, int q0, int groups, int kchunks,
                      cudaStream_t s) {
    const int free_cols = TMEM_COLS - (ix.dim_pad8 / 4 + 31) / 32 * 32;
    const int nbuf_opt = ix.opt.ts_acc_buffers;
    if (nbuf_opt != 2 && free_cols >= 384) return launch_img8<METRIC, CPS, PAIR, 128, 3>(ix, a, im, q0, groups, kchunks, s);
    if (nbuf_opt == 3 && free_cols >= 288) return  // measured 1-3 % slower than two 128-row buffers at D=768
        launch_img8<METRIC, CPS, PAIR, 96, 3>(ix, a, im, q0, groups, kchunks, s);
    return launch_img8<METRIC, CPS, PAIR, 128, 2>(ix, a, im, q0, groups, kchunks, s);
}

template <int METRIC>
int launch_img8_metric(const Index &ix, const ScanArgs &a, const ImgArgs &im, int kchunks, cudaStream_t s, int *launches) {
    const int cps = (kchunks % 3 == 0) ? 3 : 2;
    const bool pairs_ok = (ix.sm_count % 2) == 0 && ix.opt.tc_cta2;
    int gmax = ix.opt.ts_groups;
    if (gmax < 1) gmax = 1;
    if (gmax > 4) gmax = 4;
    for (int q0 = 0; q0 < a.nq;) {
        const int left = a.nq - q0;
        *launches += 1;
        if (left > QM_CTA && pairs_ok) {
            int groups = (left + 2 * QM_CTA - 1) / (2 * QM_CTA);
            if (groups > gmax) groups = gmax;
            if (cps == 3) PKV_TRY((launch_img8_shape<METRIC, 3, true>(ix, a, im, q0, groups, kchunks, s)));
            else PKV_TRY((launch_img8_shape<METRIC, 2, true>(ix, a, im, q0, groups, kchunks, s)));
            q0 += groups * 2 * QM_CTA;
        } else {
            if (cps == 3) PKV_TRY((launch_img8_shape<METRIC, 3, false>(ix, a, im, q0, 1, kchunks, s)));
            else PKV_TRY((launch_img8_shape<METRIC, 2, false>(ix, a, im, q0, 1, kchunks, s)));
            q0 += QM_CTA;
        }
    }
    return PKV_OK;
}

}  // namespace

bool img8_usable(const Index &ix) {
    return ix.d_img8 && ix.d_img8_meta && ix.dim_pad8 <= 1024 && ix.image_rows >= ix.sealed_rows &&
           (ix.dtype == PKV_F32 || ix.dtype == PKV_F16);
}

int build_img8(Index &ix, int64_t row_begin, int64_t row_end, cudaStream_t s) {
    if (row_end <= row_begin || !ix.d_img8) return PKV_OK;
    const int warps = 8;
    if (ix.img8_U == 0.f) {
        // fix the code length U = |a|/s_a of the index from the peakiness of (a sample of) the first sealed batch
        int64_t sample_end = row_end - row_begin > 262144 ? row_begin + 262144 : row_end;
        double *d_stats = nullptr, h[3] = {0, 0, 0};
        PKV_CUDA(cudaMalloc((void **)&d_stats, 3 * sizeof(double)));
        PKV_CUDA(cudaMemsetAsync(d_stats, 0, 3 * sizeof(double), s));
        const int64_t sb = (sample_end - row_begin + warps - 1) / warps;
        if (ix.dtype == PKV_F16)
            img8_stats_kernel<true><<<(unsigned)sb, warps * 32, 0, s>>>(ix.d_data, ix.pitch, ix.dim, row_begin, sample_end, d_stats);
        else
            img8_stats_kernel<false><<<(unsigned)sb, warps * 32, 0, s>>>(ix.d_data, ix.pitch, ix.dim, row_begin, sample_end, d_stats);
        cudaError_t e = cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        cudaFree(d_stats);
        PKV_CUDA(e);
        double p_ref = 1.0;  // no usable row yet: one-hot rows would still fit
        if (h[2] > 0) {
            const double mean = h[0] / h[2];
            double var = h[1] / h[2] - mean * mean;
            if (var < 0) var = 0;
            p_ref = mean + 0.1 * ix.opt.img8_peak_sigma_x10 * sqrt(var);
            if (p_ref > 1.0) p_ref = 1.0;
            if (p_ref < 1e-3) p_ref = 1e-3;
        }
        ix.img8_U = (float)(127.0 / p_ref);
    }
    const int64_t blocks = (row_end - row_begin + warps - 1) / warps;
    if (ix.dtype == PKV_F16)
        img8_build_kernel<true><<<(unsigned)blocks, warps * 32, 0, s>>>(ix.d_data, ix.pitch, ix.dim, ix.dim_pad8, row_begin,
                                                                        row_end, ix.img8_U, ix.d_img8, ix.d_img8_meta);
    else
        img8_build_kernel<false><<<(unsigned)blocks, warps * 32, 0, s>>>(ix.d_data, ix.pitch, ix.dim, ix.dim_pad8, row_begin,
                                                                         row_end, ix.img8_U, ix.d_img8, ix.d_img8_meta);
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

static ImgArgs img_args(const Index &ix, const ScanArgs &a, Workspace &ws) {
    ImgArgs im;
    im.row_meta = ix.d_img8_meta;
    im.q8 = ws.d_q8;
    im.q_meta = ws.d_q8_meta;
    im.pend = PendDev{ws.d_pend_rows, ws.d_pend_cnt, (uint32_t)ws.pend_cap, nullptr, nullptr};
    im.dlist = PendDev{ws.d_defer_rows, ws.d_defer_cnt, (uint32_t)ws.pend_cap, ws.d_defer_dots, ws.d_defer_meta};
    im.dim_pad8 = ix.dim_pad8;
    // live launches always re-score in-kernel: the thresholds can only move while the scan runs if the exact
    // scores are produced while it runs.  The chunks before it keep the separate re-score kernel (every SM gathers
    // with all its warps: the early chunks pass many more rows than the scan's own few warps can take)
    im.fused = (ix.opt.img8_fused >= 2 || a.topk.live) ? 1 : 0;
    im.defer = a.topk.defer;
    im.rows_f16 = ix.dtype == PKV_F16 ? 1 : 0;
    // the round-2 epilogue serves live launches (tight thresholds, rare survivors); the chunks of a learning prefix pass
    // several rows per query and 32-row slice, which the round-1 hold lists handle better (measured: 500k rows chunked
    // 240 k queries/s with the round-1 epilogue, 183 k with the round-2 one)
    im.dbg = nullptr;
    im.epi_exact = (ix.opt.img8_epi && a.topk.live) ? 1 : 0;
    im.park_lean = ix.opt.img8_epi ? 1 : 0;
    if (im.park_lean) im.dlist.meta = nullptr;
    return im;
}

// The pairs parked by the chunks of a search, against the thresholds the scan ended with (before the last select).
int launch_rescore_deferred(const Index &ix, const ScanArgs &a, Workspace &ws, cudaStream_t s) {
    if (a.nq <= 0) return PKV_OK;
    const ImgArgs im = img_args(ix, a, ws);
    const size_t smem = (size_t)a.dim_pad * 4;
    int ry = (8 * ix.sm_count + a.nq - 1) / a.nq;  // ~8 CTAs (64 warps) per SM in total: the gathers are latency-bound
    if (ry < 1) ry = 1;
    if (ry > 64) ry = 64;
    const dim3 grid((unsigned)a.nq, (unsigned)ry);
#define PKV_DEFERRED(M)                                                                                  \
    do {                                                                                                 \
        if (ix.dtype == PKV_F16) {                                                                       \
            auto rk = rescore_deferred_kernel<M, true>;                                                  \
            PKV_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            rk<<<grid, 256, smem, s>>>(a, im, ws.d_status);                                              \
        } else {                                                                                         \
            auto rk = rescore_deferred_kernel<M, false>;                                                 \
            PKV_CUDA(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
            rk<<<grid, 256, smem, s>>>(a, im, ws.d_status);                                              \
        }                                                                                                \
    } while (0)
    if (a.metric == PKV_COSINE) PKV_DEFERRED(PKV_COSINE);
    else if (a.metric == PKV_L2) PKV_DEFERRED(PKV_L2);
    else PKV_DEFERRED(PKV_DOT);
#undef PKV_DEFERRED
    PKV_CUDA(cudaGetLastError());
    return PKV_OK;
}

#ifdef PKV_TILE_TIMING
static unsigned long long *g_tile_timing = nullptr;
}  // namespace pkv
// debug builds only (tools/gpu/tile_timing.py): the per-warp clock sums of the last live launch
extern "C" int pkv_debug_tile_timing(unsigned long long *out, size_t n) {
    if (!pkv::g_tile_timing) return -1;
    cudaDeviceSynchronize();
    const size_t have = (size_t)256 * 20 * 14;
    return (int)cudaMemcpy(out, pkv::g_tile_timing, (n < have ? n : have) * 8, cudaMemcpyDeviceToHost);
}
namespace pkv {
#endif

int launch_scan_img8(const Index &ix, const ScanArgs &a, Workspace &ws, cudaStream_t s, int *launches) {
    if (a.row_end <= a.row_begin || a.nq <= 0) return PKV_OK;
    ImgArgs im = img_args(ix, a, ws);
    // pend counters are zeroed by reset_state_kernel / select_kernel; the query codes are made once per search
    if (!ws.q8_ready) {
        img8_prep_queries_kernel<<<(a.nq + 7) / 8, 256, 0, s>>>((const float *)a.queries, a.nq, ix.dim, ix.dim_pad,
                                                                ix.dim_pad8, a.q_mag_f, ws.d_q8, ws.d_q8_meta);
        PKV_CUDA(cudaGetLastError());
        *launches += 1;
        ws.q8_ready = true;
    }
    const int kchunks = ix.dim_pad8 / CHUNK_BYTES;
#ifdef PKV_TILE_TIMING
    unsigned long long *&d_dbg = g_tile_timing;
    const size_t dbg_n = (size_t)256 * 20 * 14;
    if (!d_dbg) cudaMalloc((void **)&d_dbg, dbg_n * 8);
    if (a.topk.live) {
        cudaMemsetAsync(d_dbg, 0, dbg_n * 8, s);
        im.dbg = d_dbg;
    }
#endif
    switch (a.metric) {
        case PKV_COSINE: PKV_TRY(launch_img8_metric<PKV_COSINE>(ix, a, im, kchunks, s, launches)); break;
        case PKV_L2: PKV_TRY(launch_img8_metric<PKV_L2>(ix, a, im, kchunks, s, launches)); break;
        default: PKV_TRY(launch_img8_metric<PKV_DOT>(ix, a, im, kchunks, s, launches)); break;
    }
    if (!im.fused) {
        PKV_TRY(launch_rescore(ix, a, im.pend, ws.d_status, s));
        *launches += 1;
    }
    return PKV_OK;
}

}  // namespace pkv
