// pkv_policy.cpp — host-side mirror of the PQL vector-filter policy, in C++ because the
// reference's host side is compiled Rust and no Rust toolchain exists in this image.
// Same names, argument meaning and error text as:
//   pql/builder/filters/embedding_types.rs  (IndexMode, DistanceFunction, DistanceAggregation, default_k)
//   pql/preprocess.rs:314-465               (resolve_vector_quant, quant_requested, normalize_variant,
//                                            validate_quant_args)
//   db/vector_quants.rs:1784-1869           (ReadyPair / resolve_ready_pair)
// (paths relative to /root/reference/panoptikon/src)
#include <cctype>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>
#include <algorithm>
#include <unordered_map>

#include "../../include/pkv.h"

namespace pkv {
int fail(int code, const char *fmt, ...);
}
using pkv::fail;

namespace {

// normalize_variant (preprocess.rs:425-431): blank / whitespace-only names mean "unset".
bool normalize_variant(const char *variant, std::string *out) {
    if (!variant) return false;
    const char *b = variant;
    const char *e = variant + strlen(variant);
    while (b < e && isspace((unsigned char)*b)) ++b;
    while (e > b && isspace((unsigned char)e[-1])) --e;
    if (b == e) return false;
    if (out) out->assign(b, e);
    return true;
}

std::string lower(const char *s) {
    std::string r(s);
    for (auto &c : r) c = (char)tolower((unsigned char)c);
    return r;
}

struct QuantEntry {
    pkv_ready_pair pair;
    pkv_index *index;
};

}  // namespace

struct pkv_space {
    std::string model;
    pkv_index *exact = nullptr;
    std::map<std::string, QuantEntry> profiles;  // name -> ready pair (absent = not ready)
    std::string default_profile;                 // "" = no default configured
    // cross-modal space (db/vector_quants.rs:480-510): the rows of the image setter and of its "t"-prefixed text
    // sibling live in ONE index with one scale; image_only marks the image setter's rows (the membership of a filter
    // without clip_xmodal, image_embeddings.rs:140-199)
    std::vector<uint64_t> image_only;            // empty = single-setter space (every row is a member)
    int64_t modality_rows = 0;
    std::mutex mu;
};

extern "C" {

int pkv_parse_index_mode(const char *name, int *mode) {
    if (!name || !mode) return fail(PKV_ERR_INVALID, "NULL argument");
    // serde rename_all = "lowercase" (embedding_types.rs:50-58)
    if (!strcmp(name, "auto")) *mode = PKV_INDEX_AUTO;
    else if (!strcmp(name, "exact")) *mode = PKV_INDEX_EXACT;
    else if (!strcmp(name, "quant")) *mode = PKV_INDEX_QUANT;
    else if (!strcmp(name, "ann")) *mode = PKV_INDEX_ANN;
    else return fail(PKV_ERR_INVALID, "unknown variant `%s`, expected one of `auto`, `exact`, `quant`, `ann`", name);
    return PKV_OK;
}

int pkv_parse_distance_function(const char *name, int case_insensitive, int *metric) {
    if (!name || !metric) return fail(PKV_ERR_INVALID, "NULL argument");
    // serde names "L2"/"COSINE" (embedding_types.rs:20-26); from_override lowercases (:28-36)
    std::string n = case_insensitive ? lower(name) : std::string(name);
    if (n == (case_insensitive ? "l2" : "L2")) *metric = PKV_L2;
    else if (n == (case_insensitive ? "cosine" : "COSINE")) *metric = PKV_COSINE;
    else return fail(PKV_ERR_INVALID, "unknown variant `%s`, expected `L2` or `COSINE`", name);
    return PKV_OK;
}

int pkv_parse_distance_aggregation(const char *name, int *agg) {
    if (!name || !agg) return fail(PKV_ERR_INVALID, "NULL argument");
    if (!strcmp(name, "MIN")) *agg = PKV_AGG_MIN;
    else if (!strcmp(name, "MAX")) *agg = PKV_AGG_MAX;
    else if (!strcmp(name, "AVG")) *agg = PKV_AGG_AVG;
    else return fail(PKV_ERR_INVALID, "unknown variant `%s`, expected one of `MIN`, `MAX`, `AVG`", name);
    return PKV_OK;
}

// preprocess.rs:436-446
int pkv_validate_quant_args(int index_mode, int64_t k) {
    if (index_mode < PKV_INDEX_AUTO || index_mode > PKV_INDEX_ANN) return fail(PKV_ERR_INVALID, "unknown index mode");
    if (index_mode == PKV_INDEX_ANN) return fail(PKV_ERR_INVALID, "index \"ann\" is reserved and not yet available");
    if (k < 1) return fail(PKV_ERR_INVALID, "k must be a positive integer");
    return PKV_OK;
}

// preprocess.rs:413-421
int pkv_quant_requested(int index_mode, const char *variant_or_null) {
    (void)variant_or_null;
    switch (index_mode) {
        case PKV_INDEX_EXACT: return 0;
        case PKV_INDEX_ANN: return 0;
        case PKV_INDEX_AUTO:
        case PKV_INDEX_QUANT: return 1;
        default: return 0;
    }
}

// preprocess.rs:331-332
int pkv_quant_strict(int index_mode, const char *variant_or_null) {
    return (index_mode == PKV_INDEX_QUANT || normalize_variant(variant_or_null, nullptr)) ? 1 : 0;
}

int pkv_space_create(const char *model, pkv_index *exact_f32, pkv_space **out) {
    if (!model || !out) return fail(PKV_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (exact_f32) {
        pkv_index_info info;
        int st = pkv_index_get_info(exact_f32, &info);
        if (st != PKV_OK) return st;
        if (info.dtype != PKV_F32) return fail(PKV_ERR_INVALID, "the exact index of a space must be PKV_F32");
    }
    pkv_space *s = new (std::nothrow) pkv_space();
    if (!s) return fail(PKV_ERR_OOM, "out of host memory");
    s->model = model;
    s->exact = exact_f32;
    *out = s;
    return PKV_OK;
}

int pkv_space_destroy(pkv_space *s) {
    delete s;
    return PKV_OK;
}

int pkv_space_set_quant(pkv_space *s, const char *profile_name, int is_default, const pkv_ready_pair *pair,
                        pkv_index *quant_i8) {
    if (!s || !profile_name) return fail(PKV_ERR_INVALID, "NULL argument");
    std::string name;
    if (!normalize_variant(profile_name, &name)) return fail(PKV_ERR_INVALID, "profile names are never empty");
    std::lock_guard<std::mutex> g(s->mu);
    if (is_default) s->default_profile = name;
    if (!pair || !quant_i8) {  // withdraw: the pair is no longer ready
        s->profiles.erase(name);
        return PKV_OK;
    }
    pkv_index_info info;
    int st = pkv_index_get_info(quant_i8, &info);
    if (st != PKV_OK) return st;
    if (info.dtype != PKV_I8) return fail(PKV_ERR_INVALID, "a quant profile index must be PKV_I8");
    // a pair is only usable with a positive finite scale and a dimension (vector_quants.rs:1846-1850)
    if (!(pair->scale > 0.0f) || !(pair->scale <= 3.402823466e+38f) || pair->dim < 1)
        return fail(PKV_ERR_INVALID, "ready pair needs a positive finite scale and a dimension");
    if (info.dim != pair->dim)
        return fail(PKV_ERR_DIM_MISMATCH, "quant index dimension %d does not match the pair's %lld", info.dim,
                    (long long)pair->dim);
    s->profiles[name] = QuantEntry{*pair, quant_i8};
    return PKV_OK;
}

// xmodal_text_sibling_name (db/vector_quants.rs:51-53)
int pkv_xmodal_text_sibling_name(const char *model, char *out, size_t cap) {
    if (!model || !out) return fail(PKV_ERR_INVALID, "NULL argument");
    const size_t need = strlen(model) + 2;
    if (cap < need) return fail(PKV_ERR_INVALID, "buffer too small for the sibling name (%zu bytes needed)", need);
    out[0] = 't';
    memcpy(out + 1, model, need - 1);
    return PKV_OK;
}

// The loop of resolve_ready_pair (db/vector_quants.rs:1817-1867) over the setters a query involves (the model and,
// under clip_xmodal, its text sibling).  states[i]: 0 = no such setter (skipped: it contributes nothing), 1 = its
// coverage row is ready and pairs[i] holds (profile_id, scale, dim), 2 = the setter exists but its pair is not ready.
// PKV_OK and *out when every existing setter is ready AND all share one scale and dim ("xmodal siblings must share
// one artifact; a mismatch means a rebuild is pending"); PKV_ERR_NOT_READY otherwise (also when no setter exists).
int pkv_resolve_ready_pair(const pkv_ready_pair *pairs, const int32_t *states, int n, pkv_ready_pair *out) {
    if (n < 0 || (n > 0 && (!pairs || !states)) || !out) return fail(PKV_ERR_INVALID, "NULL argument");
    bool have = false;
    pkv_ready_pair r{};
    for (int i = 0; i < n; ++i) {
        if (states[i] == 0) continue;
        if (states[i] != 1) return fail(PKV_ERR_NOT_READY, "setter %d of the query has no ready pair", i);
        const pkv_ready_pair &p = pairs[i];
        if (!(p.scale > 0.0f) || !(p.scale <= 3.402823466e+38f) || p.dim < 1)
            return fail(PKV_ERR_NOT_READY, "setter %d has no dimension or no usable scale", i);
        if (!have) {
            r = p;
            have = true;
        } else if (r.scale != p.scale || r.dim != p.dim) {
            return fail(PKV_ERR_NOT_READY, "xmodal siblings do not share one artifact (a rebuild is pending)");
        }
    }
    if (!have) return fail(PKV_ERR_NOT_READY, "none of the query's setters exists");
    *out = r;
    return PKV_OK;
}

int pkv_space_set_modality(pkv_space *s, const uint8_t *row_modality, int64_t n_rows) {
    if (!s) return fail(PKV_ERR_INVALID, "space handle is NULL");
    std::lock_guard<std::mutex> g(s->mu);
    if (!row_modality || n_rows <= 0) {
        s->image_only.clear();
        s->modality_rows = 0;
        return PKV_OK;
    }
    s->image_only.assign((size_t)((n_rows + 63) / 64), 0ull);
    for (int64_t r = 0; r < n_rows; ++r) {
        if (row_modality[r] > 1) return fail(PKV_ERR_INVALID, "row %lld: modality must be 0 (image) or 1 (text)", (long long)r);
        if (row_modality[r] == 0) s->image_only[(size_t)(r >> 6)] |= 1ull << (r & 63);
    }
    s->modality_rows = n_rows;
    return PKV_OK;
}

static int space_search(pkv_space *s, const float *queries, int nq, int query_dim, int metric, int index_mode,
                        const char *variant_or_null, int64_t k_arg, int depth, int clip_xmodal, int64_t *out_ids,
                        float *out_dist, int32_t *out_counts, int64_t *used_profile_id);

int pkv_space_search(pkv_space *s, const float *queries, int nq, int query_dim, int metric, int index_mode,
                     const char *variant_or_null, int64_t k_arg, int depth, int64_t *out_ids, float *out_dist,
                     int32_t *out_counts, int64_t *used_profile_id) {
    return space_search(s, queries, nq, query_dim, metric, index_mode, variant_or_null, k_arg, depth, 0, out_ids, out_dist,
                        out_counts, used_profile_id);
}

int pkv_space_search_xmodal(pkv_space *s, const float *queries, int nq, int query_dim, int metric, int index_mode,
                            const char *variant_or_null, int64_t k_arg, int depth, int clip_xmodal, int64_t *out_ids,
                            float *out_dist, int32_t *out_counts, int64_t *used_profile_id) {
    return space_search(s, queries, nq, query_dim, metric, index_mode, variant_or_null, k_arg, depth, clip_xmodal ? 1 : 0,
                        out_ids, out_dist, out_counts, used_profile_id);
}

// resolve_vector_quant (preprocess.rs:314-393) followed by the scan the compiled SQL would run.
static int space_search(pkv_space *s, const float *queries, int nq, int query_dim, int metric, int index_mode,
                        const char *variant_or_null, int64_t k_arg, int depth, int clip_xmodal, int64_t *out_ids,
                        float *out_dist, int32_t *out_counts, int64_t *used_profile_id) {
    if (!s) return fail(PKV_ERR_INVALID, "space handle is NULL");
    if (used_profile_id) *used_profile_id = -1;
    int st = pkv_validate_quant_args(index_mode, k_arg);
    if (st != PKV_OK) return st;
    if (metric != PKV_L2 && metric != PKV_COSINE)
        return fail(PKV_ERR_INVALID, "distance function must be L2 or COSINE");
    if (depth < 1) return fail(PKV_ERR_INVALID, "k must be a positive integer");

    QuantEntry chosen{};
    bool use_quant = false;
    if (pkv_quant_requested(index_mode, variant_or_null)) {
        std::string variant;
        const bool named = normalize_variant(variant_or_null, &variant);
        const bool strict = index_mode == PKV_INDEX_QUANT || named;
        std::lock_guard<std::mutex> g(s->mu);
        std::string profile_name = named ? variant : s->default_profile;
        if (profile_name.empty()) {
            if (strict) return fail(PKV_ERR_INVALID, "no default vector quant profile is configured");
        } else {
            auto it = s->profiles.find(profile_name);
            if (it == s->profiles.end()) {
                if (strict)
                    return fail(PKV_ERR_NOT_READY,
                                "vector quant profile '%s' does not exist or is not ready for model '%s'",
                                profile_name.c_str(), s->model.c_str());
            } else if ((int64_t)query_dim != it->second.pair.dim) {
                if (strict)
                    return fail(PKV_ERR_DIM_MISMATCH,
                                "query embedding dimension mismatch for model '%s' (expected %lld, got %d)",
                                s->model.c_str(), (long long)it->second.pair.dim, query_dim);
            } else {
                chosen = it->second;
                use_quant = true;
            }
        }
    }

    pkv_search_params p;
    memset(&p, 0, sizeof(p));
    p.metric = metric;
    p.k = depth > PKV_MAX_K ? PKV_MAX_K : depth;
    p.query_dtype = PKV_F32;
    // membership (image_embeddings.rs:140-199): the model's own rows, plus its text sibling's under clip_xmodal
    std::vector<uint64_t> member;
    {
        std::lock_guard<std::mutex> g(s->mu);
        if (!clip_xmodal && !s->image_only.empty()) member = s->image_only;
    }
    if (!member.empty()) {
        p.bitmap = member.data();
        p.bitmap_stride_words = 0;
    }
    if (use_quant) {
        // compute_query_quant with the pair's frozen scale happens on the GPU inside pkv_search
        // (f32 queries against an int8 index); the index must carry that same scale.
        pkv_index_info info;
        st = pkv_index_get_info(chosen.index, &info);
        if (st != PKV_OK) return st;
        const bool strict = index_mode == PKV_INDEX_QUANT || normalize_variant(variant_or_null, nullptr);
        if (!info.has_scale || info.scale != chosen.pair.scale) {
            // resolve_vector_quant never fails a non-strict `auto`: it falls back to exact (preprocess.rs:356-362)
            if (strict) return fail(PKV_ERR_NOT_READY, "quant index scale does not match the ready pair's frozen scale");
            use_quant = false;
        } else {
            if (!member.empty() && (int64_t)member.size() * 64 < info.rows)
                return fail(PKV_ERR_INVALID, "the space's modality map covers fewer rows than its quant index");
            st = pkv_search(chosen.index, queries, nq, &p, out_ids, out_dist, out_counts);
            if (st == PKV_OK && used_profile_id) *used_profile_id = chosen.pair.profile_id;
            return st;
        }
    }
    if (!s->exact) return fail(PKV_ERR_NOT_READY, "space '%s' has no exact index", s->model.c_str());
    pkv_index_info info;
    st = pkv_index_get_info(s->exact, &info);
    if (st != PKV_OK) return st;
    if (info.dim != query_dim)
        return fail(PKV_ERR_DIM_MISMATCH, "query embedding dimension mismatch for model '%s' (expected %d, got %d)",
                    s->model.c_str(), info.dim, query_dim);
    return pkv_search(s->exact, queries, nq, &p, out_ids, out_dist, out_counts);
}

}  // extern "C"

// ---- rank fusion (builder.rs:1284-1317) -----------------------------------------------------
extern "C" int pkv_fuse_ranks(int mode, int n_lists, const int64_t *const *groups, const int64_t *const *ranks,
                              const int32_t *lens, const double *weights, const int32_t *ks, int64_t *out_groups,
                              double *out_scores, int32_t cap, int32_t *out_count) {
    if (mode < 0 || mode > 2 || n_lists < 1 || !groups || !ranks || !lens || !out_groups || !out_scores || !out_count)
        return fail(PKV_ERR_INVALID, "bad fuse arguments");
    if (mode == 0 && (!weights || !ks)) return fail(PKV_ERR_INVALID, "RRF needs weights and k per list");
    const double VERY_LARGE = 9223372036854775805.0, VERY_SMALL = -9223372036854775805.0;
    std::unordered_map<int64_t, std::vector<int64_t>> seen;  // group -> rank per list (INT64_MIN = missing)
    const int64_t MISSING = INT64_MIN;
    for (int l = 0; l < n_lists; ++l) {
        if (lens[l] < 0 || (lens[l] > 0 && (!groups[l] || !ranks[l]))) return fail(PKV_ERR_INVALID, "bad list %d", l);
        for (int i = 0; i < lens[l]; ++i) {
            auto &v = seen[groups[l][i]];
            if (v.empty()) v.assign((size_t)n_lists, MISSING);
            v[(size_t)l] = ranks[l][i];
        }
    }
    struct Row {
        int64_t g;
        double score;
    };
    std::vector<Row> rows;
    rows.reserve(seen.size());
    for (auto &kv : seen) {
        double score = mode == 1 ? VERY_LARGE : (mode == 2 ? VERY_SMALL : 0.0);
        for (int l = 0; l < n_lists; ++l) {
            const int64_t r = kv.second[(size_t)l];
            if (mode == 0) {
                const double rank = r == MISSING ? VERY_LARGE : (double)r;
                score += (1.0 / ((double)ks[l] + rank)) * weights[l];  // `1.0 / (k + rank) * weight`
            } else if (mode == 1) {
                score = std::min(score, r == MISSING ? VERY_LARGE : (double)r);
            } else {
                score = std::max(score, r == MISSING ? VERY_SMALL : (double)r);
            }
        }
        rows.push_back(Row{kv.first, score});
    }
    const bool desc = mode != 1;  // RRF score and max-rank order descending, min-rank ascending
    std::sort(rows.begin(), rows.end(), [desc](const Row &a, const Row &b) {
        if (a.score != b.score) return desc ? a.score > b.score : a.score < b.score;
        return a.g < b.g;
    });
    const int32_t n = (int32_t)std::min<size_t>(rows.size(), (size_t)(cap > 0 ? cap : 0));
    for (int32_t i = 0; i < n; ++i) {
        out_groups[i] = rows[(size_t)i].g;
        out_scores[i] = rows[(size_t)i].score;
    }
    *out_count = n;
    return PKV_OK;
}
