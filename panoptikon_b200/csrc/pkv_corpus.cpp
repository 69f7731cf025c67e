// pkv_corpus.cpp — lifecycle of the HBM replica of one stored payload.
//
// In the reference a quant "pair" (profile x setter) moves pending -> building -> ready in
// vector_quant_coverage; its codes are written by a chunked, resumable backfill
// (db/vector_quants.rs:1085-1163: 5000-row chunks keyed by item_data.id, rows already at the pair's
// artifact_rev are skipped) and, for rows that arrive later, by the inline hook (:1347-1438, which also
// downgrades the pair when a vector of the wrong dimensionality shows up); a query may use the pair only while
// it is ready at its current artifact_rev (:1829-1850), and every write to the index DB bumps an epoch that
// invalidates whatever was derived from it (db/epochs.rs:38-44).
//
// A pkv_corpus is the same state machine around a pkv_index: keyed by (index_db, space, profile_id), filled by
// idempotent chunks at one artifact_rev, searchable only while ready at the (rev, epoch) the host asks for.
// f32 source blobs of an int8 corpus are quantised on the GPU with the frozen scale (the quantize_int8 of
// backfill_chunk / write_inline_quants), so the 50 s CPU backfill of 1.45 M rows becomes a stream of uploads.
#include <stdint.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../include/pkv.h"

struct pkv_corpus {
    std::string index_db, space;
    int64_t profile_id = -1;
    int device = 0, dim = 0, dtype = PKV_F32;
    pkv_index *index = nullptr;
    int state = PKV_CORPUS_PENDING;
    int64_t rev = -1;        // artifact_rev of the rows held
    uint64_t epoch = 0;      // index-DB epoch the rows are in sync with
    float scale = 0.f;       // frozen scale (int8)
    bool has_scale = false;
    int64_t cursor = INT64_MIN;             // largest item_data.id uploaded by chunks (they arrive ascending)
    std::unordered_set<int64_t> inline_ids;  // rows added by the inline hook beyond the cursor
    int64_t rows = 0;
    std::mutex mu;
    std::string error;
};

namespace {

int fail_c(pkv_corpus *c, int code, const std::string &msg) {
    if (c) c->error = msg;
    return code;
}
size_t elem_bytes(int dtype) { return dtype == PKV_F32 ? 4 : (dtype == PKV_I8 ? 1 : 2); }

int recreate_index(pkv_corpus *c) {
    if (c->index) pkv_index_destroy(c->index);
    c->index = nullptr;
    int st = pkv_index_create(c->device, c->dim, c->dtype, &c->index);
    if (st != PKV_OK) return fail_c(c, st, pkv_last_error());
    c->rows = 0;
    c->cursor = INT64_MIN;
    c->inline_ids.clear();
    return PKV_OK;
}

// rows as the index stores them: source blobs of the index dtype are taken as they are; f32 blobs of an int8
// corpus go through the GPU codec with the frozen scale
int append_rows(pkv_corpus *c, const int64_t *ids, const void *blobs, int src_dtype, int64_t n) {
    if (n == 0) return PKV_OK;
    if (src_dtype == c->dtype) {
        int st = pkv_index_append(c->index, blobs, ids, n);
        return st == PKV_OK ? PKV_OK : fail_c(c, st, pkv_last_error());
    }
    if (c->dtype == PKV_I8 && src_dtype == PKV_F32) {
        if (!c->has_scale) return fail_c(c, PKV_ERR_NOT_READY, "the pair has no frozen scale artifact yet");
        std::vector<int8_t> codes((size_t)n * c->dim);
        int st = pkv_quantize_int8(c->device, static_cast<const float *>(blobs), n * (int64_t)c->dim, c->scale, codes.data());
        if (st == PKV_OK) st = pkv_index_append(c->index, codes.data(), ids, n);
        return st == PKV_OK ? PKV_OK : fail_c(c, st, pkv_last_error());
    }
    return fail_c(c, PKV_ERR_INVALID, "source blobs must be of the corpus dtype, or f32 for an int8 corpus");
}

}  // namespace

extern "C" {

int pkv_corpus_create(int device, int dim, int dtype, const char *index_db, const char *space, int64_t profile_id,
                      pkv_corpus **out) {
    if (!out || !index_db || !space) return PKV_ERR_INVALID;
    *out = nullptr;
    if (dim < 1 || (dtype != PKV_F32 && dtype != PKV_I8 && dtype != PKV_F16)) return PKV_ERR_INVALID;
    pkv_corpus *c = new (std::nothrow) pkv_corpus();
    if (!c) return PKV_ERR_OOM;
    c->device = device;
    c->dim = dim;
    c->dtype = dtype;
    c->index_db = index_db;
    c->space = space;
    c->profile_id = profile_id;
    *out = c;
    return PKV_OK;
}

int pkv_corpus_destroy(pkv_corpus *c) {
    if (!c) return PKV_OK;
    if (c->index) pkv_index_destroy(c->index);
    delete c;
    return PKV_OK;
}

const char *pkv_corpus_last_error(const pkv_corpus *c) { return c ? c->error.c_str() : ""; }

// pending/ready -> building at a NEW (artifact_rev, epoch): rows of another revision are dropped, as a rebuild
// rewrites every code under the new scale.  Beginning again at the SAME revision resumes (idempotent).
int pkv_corpus_begin(pkv_corpus *c, int64_t artifact_rev, const uint8_t *artifact, size_t artifact_len, uint64_t index_epoch) {
    if (!c) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    float scale = 0.f;
    if (c->dtype == PKV_I8) {
        // "artifact is not a scale; refusing to backfill" (db/vector_quants.rs:1136-1145)
        if (pkv_artifact_scale(artifact, artifact_len, &scale) != PKV_OK)
            return fail_c(c, PKV_ERR_INVALID, "Invalid vector quant scale artifact");
    }
    const bool resume = c->index && c->rev == artifact_rev && c->state != PKV_CORPUS_PENDING &&
                        (c->dtype != PKV_I8 || (c->has_scale && c->scale == scale));
    if (!resume) {
        int st = recreate_index(c);
        if (st != PKV_OK) return st;
        if (c->dtype == PKV_I8) {
            st = pkv_index_set_scale(c->index, artifact, artifact_len);
            if (st != PKV_OK) return fail_c(c, st, pkv_last_error());
            c->scale = scale;
            c->has_scale = true;
        }
    }
    c->rev = artifact_rev;
    c->epoch = index_epoch;
    c->state = PKV_CORPUS_BUILDING;
    return PKV_OK;
}

// One backfill chunk (db/vector_quants.rs:1085-1163): `ids` ascending.  Rows at or below the cursor (already
// uploaded at this revision) and rows the inline hook already added are skipped, so replaying a chunk after a crash
// or a retry writes nothing twice.  *written rows, *cursor = the resume point for the next chunk.
int pkv_corpus_upload_chunk(pkv_corpus *c, int64_t artifact_rev, const int64_t *ids, const void *blobs, int src_dtype,
                            int64_t n, int64_t *written, int64_t *cursor) {
    if (!c || n < 0 || (n > 0 && (!ids || !blobs))) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    if (written) *written = 0;
    if (cursor) *cursor = c->cursor;
    // "zero rows means ... the pair is no longer building (an explicit rebuild was marked mid-build ...; writing
    // codes at the frozen rev under a new scale would corrupt the pair)"
    if (c->state != PKV_CORPUS_BUILDING || c->rev != artifact_rev) return PKV_OK;
    const size_t row_bytes = (size_t)c->dim * elem_bytes(src_dtype);
    int64_t done = 0, run_begin = -1;
    auto flush = [&](int64_t end) -> int {
        if (run_begin < 0) return PKV_OK;
        int st = append_rows(c, ids + run_begin, static_cast<const uint8_t *>(blobs) + (size_t)run_begin * row_bytes, src_dtype,
                             end - run_begin);
        if (st == PKV_OK) done += end - run_begin;
        run_begin = -1;
        return st;
    };
    int64_t prev = INT64_MIN;
    for (int64_t i = 0; i < n; ++i) {
        if (ids[i] <= prev) return fail_c(c, PKV_ERR_INVALID, "chunk ids must ascend (ORDER BY d.id)");
        prev = ids[i];
        const bool skip = ids[i] <= c->cursor || c->inline_ids.count(ids[i]) != 0;
        if (skip) {
            int st = flush(i);
            if (st != PKV_OK) return st;
        } else if (run_begin < 0) {
            run_begin = i;
        }
    }
    int st = flush(n);
    if (st != PKV_OK) return st;
    if (n > 0 && ids[n - 1] > c->cursor) c->cursor = ids[n - 1];
    // inline rows at or below the cursor are covered by it from now on
    for (auto it = c->inline_ids.begin(); it != c->inline_ids.end();)
        it = (*it <= c->cursor) ? c->inline_ids.erase(it) : ++it;
    c->rows += done;
    if (written) *written = done;
    if (cursor) *cursor = c->cursor;
    return PKV_OK;
}

// The inline hook (write_inline_quants, db/vector_quants.rs:1347-1438): one freshly written embedding joins the replica
// while the pair is building or ready.  A blob of the wrong dimensionality DOWNGRADES the pair to pending ("search
// falls back to exact for that setter and the next reconcile repairs it") and reports PKV_ERR_DIM_MISMATCH.
int pkv_corpus_append_inline(pkv_corpus *c, int64_t data_id, const void *blob, size_t blob_bytes, int src_dtype,
                             uint64_t index_epoch) {
    if (!c || !blob) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    if (c->state == PKV_CORPUS_PENDING) return PKV_OK;  // nothing to keep in sync: the next build reads it from the DB
    if (blob_bytes != (size_t)c->dim * elem_bytes(src_dtype)) {
        c->state = PKV_CORPUS_PENDING;
        return fail_c(c, PKV_ERR_DIM_MISMATCH,
                      "embedding dimensionality does not match the quant coverage snapshot; coverage downgraded to pending");
    }
    if (data_id <= c->cursor || c->inline_ids.count(data_id)) {
        c->epoch = index_epoch;
        return PKV_OK;  // the upsert's ON CONFLICT: the row is already there at this revision
    }
    const bool was_ready = c->state == PKV_CORPUS_READY;
    int st = append_rows(c, &data_id, blob, src_dtype, 1);
    if (st != PKV_OK) return st;
    c->inline_ids.insert(data_id);
    c->rows += 1;
    if (was_ready) {
        st = pkv_index_seal(c->index);  // the new row becomes searchable at once
        if (st != PKV_OK) return fail_c(c, st, pkv_last_error());
    }
    c->epoch = index_epoch;
    return PKV_OK;
}

// finish_space_build: the pair's rows are complete at this revision -> ready.
int pkv_corpus_finish(pkv_corpus *c, int64_t artifact_rev, uint64_t index_epoch) {
    if (!c) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    if (c->state != PKV_CORPUS_BUILDING || c->rev != artifact_rev)
        return fail_c(c, PKV_ERR_NOT_READY, "the pair is not building at this artifact_rev");
    int st = pkv_index_seal(c->index);
    if (st != PKV_OK) return fail_c(c, st, pkv_last_error());
    c->state = PKV_CORPUS_READY;
    c->epoch = index_epoch;
    return PKV_OK;
}

// bump_index_epoch for writes that were NOT mirrored into the replica (a delete, a re-extraction, a rebuild mark):
// the replica is stale -> pending; `auto` queries fall back to exact until it is rebuilt.
int pkv_corpus_invalidate(pkv_corpus *c) {
    if (!c) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    c->state = PKV_CORPUS_PENDING;
    return PKV_OK;
}

// resolve_ready_pair's coverage check (db/vector_quants.rs:1829-1850): usable only while ready at the revision and
// the epoch the caller read from the DB.  PKV_OK (+ the pair, + the index) or PKV_ERR_NOT_READY.
int pkv_corpus_ready(pkv_corpus *c, int64_t artifact_rev, uint64_t index_epoch, pkv_ready_pair *pair, pkv_index **index) {
    if (!c) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    if (index) *index = nullptr;
    if (c->state != PKV_CORPUS_READY) return fail_c(c, PKV_ERR_NOT_READY, "the replica is not ready");
    if (c->rev != artifact_rev) return fail_c(c, PKV_ERR_NOT_READY, "the replica holds another artifact_rev");
    if (c->epoch != index_epoch) return fail_c(c, PKV_ERR_NOT_READY, "the index DB has moved on (epoch) since the replica was synced");
    if (pair) {
        pair->profile_id = c->profile_id;
        pair->scale = c->has_scale ? c->scale : 0.f;
        pair->dim = c->dim;
    }
    if (index) *index = c->index;
    return PKV_OK;
}

int pkv_corpus_get_info(pkv_corpus *c, pkv_corpus_info *info) {
    if (!c || !info) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(c->mu);
    memset(info, 0, sizeof(*info));
    info->state = c->state;
    info->dtype = c->dtype;
    info->dim = c->dim;
    info->profile_id = c->profile_id;
    info->artifact_rev = c->rev;
    info->index_epoch = c->epoch;
    info->rows = c->rows;
    info->cursor = c->cursor;
    return PKV_OK;
}

}  // extern "C"
