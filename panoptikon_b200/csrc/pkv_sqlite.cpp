// pkv_sqlite.cpp — the SQLite seam of the vector operator.
//
// The reference makes its vector arithmetic available to every connection with
//     sqlite3_auto_extension(sqlite3_vec_init)                      (db/sql_functions.rs:105-128)
// and its filter compilers render per-row `vec_distance_*(stored, ?)` calls inside a materialised CTE
// `dist_{cte}(item_id, file_id[, data_id], d)` (pql/builder/filters/exact.rs:106-165).  A per-row scalar cannot use a
// GPU; what can is a TABLE-VALUED function that returns the (item_data.id, d) rows of one whole query, joined exactly
// where `embeddings` / `embedding_quants` are joined today:
//
//     SELECT item_data.item_id, files.id AS file_id, t.d
//     FROM pkv_topk('<setter name>', ?query_blob, ?k, 'cosine') AS t
//          JOIN item_data ON item_data.id = t.id JOIN files ON files.item_id = item_data.item_id ...
//
// sqlite3_pkv_init has the signature of an extension entry point (int(sqlite3*, char**, const sqlite3_api_routines*))
// so it can be handed to sqlite3_auto_extension like sqlite3_vec_init, or loaded with load_extension().  It registers
//   pkv_topk(model, query, k [, metric])   eponymous virtual table: columns (id INTEGER, d REAL, rank INTEGER)
//   pkv_version()                          scalar, text
//   pkv_scale_from_absmax(x)               scalar (db/vector_quants.rs:1465-1471), pure host arithmetic
//   pkv_last_execute_ms()                  scalar: device milliseconds of this thread's last search - the part of
//                                          SearchMetrics.execute (api/search.rs:68-103) spent in the scan
// `model` names an index registered by the host with pkv_sqlite_register_index() (the setter's resident corpus).
//
// No sqlite3.h / sqlite3ext.h exists in this build image, so the handful of declarations used are restated here.
// sqlite3_api_routines is a struct of function pointers that has only ever been APPENDED to (sqlite3ext.h); it is
// addressed by slot number, and sqlite3_pkv_init refuses to register anything if slot 67 (libversion_number) does not
// answer with a 3.x version number.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <strings.h>

#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/pkv.h"

namespace {

// ---- the slice of sqlite3ext.h this file needs --------------------------------------------------
struct sqlite3;
struct sqlite3_context;
struct sqlite3_value;
typedef long long sqlite3_int64;

enum ApiSlot {
    S_create_function = 45,
    S_create_module = 47,
    S_declare_vtab = 50,
    S_free = 58,
    S_libversion_number = 67,
    S_malloc = 68,
    S_result_double = 79,
    S_result_error = 80,
    S_result_int64 = 83,
    S_result_null = 84,
    S_result_text = 85,
    S_value_blob = 102,
    S_value_bytes = 103,
    S_value_double = 105,
    S_value_int64 = 107,
    S_value_text = 109,
    S_value_type = 113,
};
constexpr int SQLITE_OK = 0, SQLITE_ERROR = 1, SQLITE_NOMEM = 7, SQLITE_CONSTRAINT = 19;
constexpr int SQLITE_INTEGER = 1, SQLITE_FLOAT = 2, SQLITE_TEXT = 3, SQLITE_BLOB = 4, SQLITE_NULL = 5;
constexpr int SQLITE_UTF8 = 1, SQLITE_DETERMINISTIC = 0x800;
constexpr int SQLITE_INDEX_CONSTRAINT_EQ = 2;
#define PKV_SQLITE_TRANSIENT ((void (*)(void *))(intptr_t)-1)

const void *const *g_api = nullptr;  // process-wide: every connection of a process shares one SQLite library

template <typename F>
F api(int slot) {
    return reinterpret_cast<F>(const_cast<void *>(g_api[slot]));
}

struct sqlite3_module;
struct sqlite3_vtab {
    const sqlite3_module *pModule;
    int nRef;
    char *zErrMsg;
};
struct sqlite3_vtab_cursor {
    sqlite3_vtab *pVtab;
};
struct sqlite3_index_constraint {
    int iColumn;
    unsigned char op;
    unsigned char usable;
    int iTermOffset;
};
struct sqlite3_index_orderby {
    int iColumn;
    unsigned char desc;
};
struct sqlite3_index_constraint_usage {
    int argvIndex;
    unsigned char omit;
};
struct sqlite3_index_info {
    int nConstraint;
    sqlite3_index_constraint *aConstraint;
    int nOrderBy;
    sqlite3_index_orderby *aOrderBy;
    sqlite3_index_constraint_usage *aConstraintUsage;
    int idxNum;
    char *idxStr;
    int needToFreeIdxStr;
    int orderByConsumed;
    double estimatedCost;
    sqlite3_int64 estimatedRows;
    int idxFlags;
    unsigned long long colUsed;
};
struct sqlite3_module {
    int iVersion;
    int (*xCreate)(sqlite3 *, void *, int, const char *const *, sqlite3_vtab **, char **);
    int (*xConnect)(sqlite3 *, void *, int, const char *const *, sqlite3_vtab **, char **);
    int (*xBestIndex)(sqlite3_vtab *, sqlite3_index_info *);
    int (*xDisconnect)(sqlite3_vtab *);
    int (*xDestroy)(sqlite3_vtab *);
    int (*xOpen)(sqlite3_vtab *, sqlite3_vtab_cursor **);
    int (*xClose)(sqlite3_vtab_cursor *);
    int (*xFilter)(sqlite3_vtab_cursor *, int, const char *, int, sqlite3_value **);
    int (*xNext)(sqlite3_vtab_cursor *);
    int (*xEof)(sqlite3_vtab_cursor *);
    int (*xColumn)(sqlite3_vtab_cursor *, sqlite3_context *, int);
    int (*xRowid)(sqlite3_vtab_cursor *, sqlite3_int64 *);
    int (*xUpdate)(sqlite3_vtab *, int, sqlite3_value **, sqlite3_int64 *);
    int (*xBegin)(sqlite3_vtab *);
    int (*xSync)(sqlite3_vtab *);
    int (*xCommit)(sqlite3_vtab *);
    int (*xRollback)(sqlite3_vtab *);
    int (*xFindFunction)(sqlite3_vtab *, int, const char *, void (**)(sqlite3_context *, int, sqlite3_value **), void **);
    int (*xRename)(sqlite3_vtab *, const char *);
};

// ---- registry: setter name -> resident index --------------------------------------------------------
std::mutex g_reg_mu;
std::map<std::string, pkv_index *> g_registry;
thread_local double g_last_execute_ms = 0.0;

char *sqlite_strdup(const std::string &s) {
    char *p = static_cast<char *>(api<void *(*)(int)>(S_malloc)((int)s.size() + 1));
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
}
void set_vtab_error(sqlite3_vtab *vt, const std::string &msg) {
    if (vt->zErrMsg) api<void (*)(void *)>(S_free)(vt->zErrMsg);
    vt->zErrMsg = sqlite_strdup(msg);
}

// ---- pkv_topk: columns 0 id, 1 d, 2 rank | hidden 3 model, 4 query, 5 k, 6 metric -----------------------
struct TopkCursor {
    sqlite3_vtab_cursor base;
    std::vector<int64_t> ids;
    std::vector<float> dist;
    int count = 0, pos = 0;
};

int topk_connect(sqlite3 *db, void *, int, const char *const *, sqlite3_vtab **out, char **) {
    const int rc = api<int (*)(sqlite3 *, const char *)>(S_declare_vtab)(
        db, "CREATE TABLE x(id INTEGER, d REAL, rank INTEGER, model HIDDEN, query HIDDEN, k HIDDEN, metric HIDDEN)");
    if (rc != SQLITE_OK) return rc;
    sqlite3_vtab *vt = new (std::nothrow) sqlite3_vtab();
    if (!vt) return SQLITE_NOMEM;
    memset(vt, 0, sizeof(*vt));
    *out = vt;
    return SQLITE_OK;
}
int topk_disconnect(sqlite3_vtab *vt) {
    if (vt->zErrMsg) api<void (*)(void *)>(S_free)(vt->zErrMsg);
    delete vt;
    return SQLITE_OK;
}
// The arguments arrive as equality constraints on the hidden columns; idxNum records which are present (bits 0..3 =
// model, query, k, metric) and they are handed to xFilter in that order.
int topk_best_index(sqlite3_vtab *, sqlite3_index_info *info) {
    int arg_of[4] = {-1, -1, -1, -1};
    for (int i = 0; i < info->nConstraint; ++i) {
        const sqlite3_index_constraint &c = info->aConstraint[i];
        if (c.iColumn < 3 || c.iColumn > 6 || c.op != SQLITE_INDEX_CONSTRAINT_EQ) continue;
        if (!c.usable) return SQLITE_CONSTRAINT;  // an argument that depends on a later table: not this plan
        arg_of[c.iColumn - 3] = i;
    }
    if (arg_of[0] < 0 || arg_of[1] < 0) return SQLITE_CONSTRAINT;  // model and query are mandatory
    int n = 0, mask = 0;
    for (int a = 0; a < 4; ++a) {
        if (arg_of[a] < 0) continue;
        info->aConstraintUsage[arg_of[a]].argvIndex = ++n;
        info->aConstraintUsage[arg_of[a]].omit = 1;
        mask |= 1 << a;
    }
    info->idxNum = mask;
    info->estimatedCost = 1000.0;
    info->estimatedRows = 100;
    // rows come out best first: ORDER BY d (or rank) ascending needs no sort
    if (info->nOrderBy == 1 && (info->aOrderBy[0].iColumn == 1 || info->aOrderBy[0].iColumn == 2) && !info->aOrderBy[0].desc)
        info->orderByConsumed = 1;
    return SQLITE_OK;
}
int topk_open(sqlite3_vtab *, sqlite3_vtab_cursor **out) {
    TopkCursor *c = new (std::nothrow) TopkCursor();
    if (!c) return SQLITE_NOMEM;
    c->base.pVtab = nullptr;
    *out = &c->base;
    return SQLITE_OK;
}
int topk_close(sqlite3_vtab_cursor *cur) {
    delete reinterpret_cast<TopkCursor *>(cur);
    return SQLITE_OK;
}
int topk_filter(sqlite3_vtab_cursor *cur, int idxNum, const char *, int argc, sqlite3_value **argv) {
    TopkCursor *c = reinterpret_cast<TopkCursor *>(cur);
    sqlite3_vtab *vt = cur->pVtab;
    c->count = c->pos = 0;
    auto vtype = api<int (*)(sqlite3_value *)>(S_value_type);
    auto vtext = api<const unsigned char *(*)(sqlite3_value *)>(S_value_text);
    int a = 0;
    sqlite3_value *v_model = (idxNum & 1) && a < argc ? argv[a++] : nullptr;
    sqlite3_value *v_query = (idxNum & 2) && a < argc ? argv[a++] : nullptr;
    sqlite3_value *v_k = (idxNum & 4) && a < argc ? argv[a++] : nullptr;
    sqlite3_value *v_metric = (idxNum & 8) && a < argc ? argv[a++] : nullptr;
    if (!v_model || !v_query || vtype(v_model) != SQLITE_TEXT || vtype(v_query) != SQLITE_BLOB) {
        set_vtab_error(vt, "pkv_topk(model TEXT, query BLOB, k INTEGER [, metric TEXT])");
        return SQLITE_ERROR;
    }
    const std::string model(reinterpret_cast<const char *>(vtext(v_model)));
    pkv_index *ix = nullptr;
    {
        std::lock_guard<std::mutex> g(g_reg_mu);
        auto it = g_registry.find(model);
        if (it != g_registry.end()) ix = it->second;
    }
    if (!ix) {
        set_vtab_error(vt, "pkv_topk: no resident index is registered for setter '" + model + "'");
        return SQLITE_ERROR;
    }
    long long k = PKV_DEFAULT_K;
    if (v_k && vtype(v_k) != SQLITE_NULL) k = api<sqlite3_int64 (*)(sqlite3_value *)>(S_value_int64)(v_k);
    if (k < 1) {
        set_vtab_error(vt, "k must be a positive integer");  // pql/preprocess.rs:442-444
        return SQLITE_ERROR;
    }
    if (k > PKV_MAX_K) k = PKV_MAX_K;  // the server clamps LIMIT the same way (api/search.rs:51)
    int metric = PKV_COSINE;
    if (v_metric && vtype(v_metric) == SQLITE_TEXT) {
        const char *m = reinterpret_cast<const char *>(vtext(v_metric));
        if (!strcasecmp(m, "dot")) metric = PKV_DOT;
        else if (pkv_parse_distance_function(m, 1, &metric) != PKV_OK) {
            set_vtab_error(vt, std::string("pkv_topk: unknown distance function '") + m + "'");
            return SQLITE_ERROR;
        }
    }
    pkv_index_info info;
    if (pkv_index_get_info(ix, &info) != PKV_OK) {
        set_vtab_error(vt, pkv_last_error());
        return SQLITE_ERROR;
    }
    const void *blob = api<const void *(*)(sqlite3_value *)>(S_value_blob)(v_query);
    const int bytes = api<int (*)(sqlite3_value *)>(S_value_bytes)(v_query);
    pkv_search_params p;
    memset(&p, 0, sizeof(p));
    p.metric = metric;
    p.k = (int)k;
    if (bytes == info.dim * 4) p.query_dtype = PKV_F32;  // serialize_f32 blob (pql/embedding_utils.rs:15-21)
    else if (bytes == info.dim && info.dtype == PKV_I8) p.query_dtype = PKV_I8;  // QuantResolved.query_quant
    else if (bytes == info.dim * 2 && info.dtype == PKV_F16) p.query_dtype = PKV_F16;
    else {
        char msg[256];
        // the wording of the reference's dimension check (pql/preprocess.rs:372-381)
        snprintf(msg, sizeof(msg), "query embedding has %d bytes but setter '%s' stores %d-dimensional vectors", bytes,
                 model.c_str(), info.dim);
        set_vtab_error(vt, msg);
        return SQLITE_ERROR;
    }
    c->ids.assign((size_t)k, -1);
    c->dist.assign((size_t)k, 0.f);
    int32_t count = 0;
    const int st = pkv_search(ix, blob, 1, &p, c->ids.data(), c->dist.data(), &count);
    if (st != PKV_OK) {
        set_vtab_error(vt, std::string("pkv_topk: ") + pkv_last_error());
        return SQLITE_ERROR;
    }
    pkv_counters ctr;
    if (pkv_index_counters(ix, &ctr) == PKV_OK) g_last_execute_ms = ctr.last_total_ms;
    c->count = count;
    return SQLITE_OK;
}
int topk_next(sqlite3_vtab_cursor *cur) {
    reinterpret_cast<TopkCursor *>(cur)->pos++;
    return SQLITE_OK;
}
int topk_eof(sqlite3_vtab_cursor *cur) {
    TopkCursor *c = reinterpret_cast<TopkCursor *>(cur);
    return c->pos >= c->count;
}
int topk_column(sqlite3_vtab_cursor *cur, sqlite3_context *ctx, int col) {
    TopkCursor *c = reinterpret_cast<TopkCursor *>(cur);
    if (col == 0) api<void (*)(sqlite3_context *, sqlite3_int64)>(S_result_int64)(ctx, c->ids[c->pos]);
    else if (col == 1) {
        const float d = c->dist[c->pos];
        if (isnan(d)) api<void (*)(sqlite3_context *)>(S_result_null)(ctx);  // zero-norm row: SQL NULL, as sqlite-vec
        else api<void (*)(sqlite3_context *, double)>(S_result_double)(ctx, (double)d);
    } else if (col == 2) api<void (*)(sqlite3_context *, sqlite3_int64)>(S_result_int64)(ctx, c->pos + 1);
    else api<void (*)(sqlite3_context *)>(S_result_null)(ctx);
    return SQLITE_OK;
}
int topk_rowid(sqlite3_vtab_cursor *cur, sqlite3_int64 *out) {
    *out = reinterpret_cast<TopkCursor *>(cur)->pos + 1;
    return SQLITE_OK;
}

const sqlite3_module g_topk_module = {
    /*iVersion*/ 1,    /*xCreate*/ nullptr, topk_connect, topk_best_index, topk_disconnect, /*xDestroy*/ nullptr,
    topk_open,         topk_close,          topk_filter,  topk_next,       topk_eof,        topk_column,
    topk_rowid,        nullptr,             nullptr,      nullptr,         nullptr,         nullptr,
    nullptr,           nullptr,
};

// ---- scalars ---------------------------------------------------------------------------------------
void fn_version(sqlite3_context *ctx, int, sqlite3_value **) {
    char buf[64];
    snprintf(buf, sizeof(buf), "libpkv abi %d (sm_100a)", pkv_abi_version());
    api<void (*)(sqlite3_context *, const char *, int, void (*)(void *))>(S_result_text)(ctx, buf, -1, PKV_SQLITE_TRANSIENT);
}
void fn_scale_from_absmax(sqlite3_context *ctx, int argc, sqlite3_value **argv) {
    if (argc != 1 || api<int (*)(sqlite3_value *)>(S_value_type)(argv[0]) == SQLITE_NULL) {
        api<void (*)(sqlite3_context *)>(S_result_null)(ctx);
        return;
    }
    const double x = api<double (*)(sqlite3_value *)>(S_value_double)(argv[0]);
    api<void (*)(sqlite3_context *, double)>(S_result_double)(ctx, (double)pkv_scale_from_absmax((float)x));
}
void fn_last_execute_ms(sqlite3_context *ctx, int, sqlite3_value **) {
    api<void (*)(sqlite3_context *, double)>(S_result_double)(ctx, g_last_execute_ms);
}

}  // namespace

extern "C" {

int pkv_sqlite_register_index(const char *model, pkv_index *h) {
    if (!model || !*model) return PKV_ERR_INVALID;
    std::lock_guard<std::mutex> g(g_reg_mu);
    if (h) g_registry[model] = h;
    else g_registry.erase(model);
    return PKV_OK;
}

int sqlite3_pkv_init(void *db_, char **pzErrMsg, const void *pApi) {
    (void)pzErrMsg;
    if (!db_ || !pApi) return SQLITE_ERROR;
    const void *const *slots = static_cast<const void *const *>(pApi);
    // slot check: sqlite3_libversion_number() must answer with a 3.x version
    const int ver = reinterpret_cast<int (*)(void)>(const_cast<void *>(slots[S_libversion_number]))();
    if (ver < 3008002 || ver >= 4000000) return SQLITE_ERROR;  // 3.8.2: the sqlite3_index_info fields used above
    g_api = slots;
    sqlite3 *db = static_cast<sqlite3 *>(db_);
    typedef void (*Fn)(sqlite3_context *, int, sqlite3_value **);
    auto create_function = api<int (*)(sqlite3 *, const char *, int, int, void *, Fn, Fn, void (*)(sqlite3_context *))>(
        S_create_function);
    int rc = create_function(db, "pkv_version", 0, SQLITE_UTF8 | SQLITE_DETERMINISTIC, nullptr, fn_version, nullptr, nullptr);
    if (rc == SQLITE_OK)
        rc = create_function(db, "pkv_scale_from_absmax", 1, SQLITE_UTF8 | SQLITE_DETERMINISTIC, nullptr, fn_scale_from_absmax,
                             nullptr, nullptr);
    if (rc == SQLITE_OK) rc = create_function(db, "pkv_last_execute_ms", 0, SQLITE_UTF8, nullptr, fn_last_execute_ms, nullptr, nullptr);
    if (rc == SQLITE_OK)
        rc = api<int (*)(sqlite3 *, const char *, const sqlite3_module *, void *)>(S_create_module)(db, "pkv_topk", &g_topk_module,
                                                                                                  nullptr);
    return rc;
}

}  // extern "C"
