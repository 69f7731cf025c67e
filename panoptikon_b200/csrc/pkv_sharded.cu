// pkv_sharded.cu — the corpus row-sharded over GPUs, inside the library (include/pkv.h "sharding").
//
// Two shapes, one total order (distance, then global row):
//   pkv_sharded  ONE process driving S shards (one per listed device; the same device may be listed several times).
//                pkv_sharded_search runs the per-shard scans concurrently (a host thread + stream per shard), gathers
//                the S x nq x k candidate lists on the first shard's device and merges them there: what a
//                single-process server (the reference: one Rust process) binds.  No collective library is involved:
//                the lists travel by cudaMemcpyAsync (peer copies over NVLink between devices).
//   pkv_comm     ONE process PER GPU (torchrun / MPI style).  The exchange is ONE ncclAllGather of the packed per-shard
//                lists over NVLink/NVSwitch, enqueued behind the shard's scan on the caller's stream, then the merge
//                kernel that reads the gathered buffer directly.  NCCL is resolved at run time (dlopen of
//                libnccl.so.2: inside a torch process that is the copy torch already loaded), so libpkv.so carries no
//                link-time dependency on it and loads on machines without it.
#include <dlfcn.h>

#include <thread>

#include "pkv_internal.cuh"

namespace pkv {

// ------------------------------------------------------------------ NCCL, resolved at run time
namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_UINT8 = 1, NCCL_UINT64 = 5 };
enum { NCCL_MIN = 3 };

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string error;
};

Nccl &nccl() {
    static Nccl n = [] {
        Nccl x;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            x.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (x.lib) break;
        }
        if (!x.lib) {
            x.error = "libnccl.so.2 not found (dlopen)";
            return x;
        }
        auto sym = [&](const char *s) {
            void *p = dlsym(x.lib, s);
            if (!p && x.error.empty()) x.error = std::string("NCCL symbol missing: ") + s;
            return p;
        };
        x.GetUniqueId = (decltype(x.GetUniqueId))sym("ncclGetUniqueId");
        x.CommInitRank = (decltype(x.CommInitRank))sym("ncclCommInitRank");
        x.CommDestroy = (decltype(x.CommDestroy))sym("ncclCommDestroy");
        x.AllGather = (decltype(x.AllGather))sym("ncclAllGather");
        x.AllReduce = (decltype(x.AllReduce))sym("ncclAllReduce");
        x.GetErrorString = (decltype(x.GetErrorString))sym("ncclGetErrorString");
        return x;
    }();
    return n;
}

#define PKV_NCCL(expr)                                                                                   \
    do {                                                                                                 \
        int _r = (expr);                                                                                 \
        if (_r != NCCL_SUCCESS)                                                                          \
            return fail(PKV_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(_r) : "?"); \
    } while (0)

}  // namespace

struct Comm {
    int device = 0, rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    std::mutex mu;  // one exchange at a time per communicator (NCCL calls of one communicator must not interleave)
    void *d_packed = nullptr, *d_gathered = nullptr;
    size_t packed_bytes = 0;
};

struct Sharded {
    std::vector<Index *> shards;
    std::vector<int> devices;
    int dim = 0, dtype = PKV_F32;
    int64_t per_shard = 0;  // rows per shard (multiple of 64: a global bitmap slices on word boundaries); 0 = not reserved
    int64_t rows = 0;
    // merge staging on the first shard's device
    void *d_gathered = nullptr;
    int64_t *d_ids = nullptr;
    float *d_dist = nullptr;
    int32_t *d_counts = nullptr;
    size_t stage_entries = 0, stage_queries = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
};

}  // namespace pkv

using namespace pkv;

extern "C" {

// ================================================================ one process, S shards
int pkv_sharded_create(const int *devices, int n_shards, int dim, int dtype, pkv_sharded **out) {
    if (!out) return fail(PKV_ERR_INVALID, "out must not be NULL");
    *out = nullptr;
    if (!devices || n_shards < 1 || n_shards > 64) return fail(PKV_ERR_INVALID, "need 1..64 shards");
    Sharded *sh = new (std::nothrow) Sharded();
    if (!sh) return fail(PKV_ERR_OOM, "out of host memory");
    sh->dim = dim;
    sh->dtype = dtype;
    for (int i = 0; i < n_shards; ++i) {
        pkv_index *ix = nullptr;
        const int st = pkv_index_create(devices[i], dim, dtype, &ix);
        if (st != PKV_OK) {
            for (Index *p : sh->shards) pkv_index_destroy(reinterpret_cast<pkv_index *>(p));
            delete sh;
            return st;
        }
        sh->shards.push_back(reinterpret_cast<Index *>(ix));
        sh->devices.push_back(devices[i]);
    }
    *out = reinterpret_cast<pkv_sharded *>(sh);
    return PKV_OK;
}

int pkv_sharded_destroy(pkv_sharded *h) {
    if (!h) return PKV_OK;
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    {
        DeviceGuard g;
        if (g.use(sh->devices[0]) == PKV_OK) {
            cudaFree(sh->d_gathered);
            cudaFree(sh->d_ids);
            cudaFree(sh->d_dist);
            cudaFree(sh->d_counts);
            if (sh->stream) cudaStreamDestroy(sh->stream);
        }
    }
    for (Index *p : sh->shards) pkv_index_destroy(reinterpret_cast<pkv_index *>(p));
    delete sh;
    return PKV_OK;
}

int pkv_sharded_shard_count(const pkv_sharded *h) {
    return h ? (int)reinterpret_cast<const Sharded *>(h)->shards.size() : 0;
}

int pkv_sharded_shard(pkv_sharded *h, int i, pkv_index **out) {
    if (!h || !out) return fail(PKV_ERR_INVALID, "NULL argument");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    if (i < 0 || i >= (int)sh->shards.size()) return fail(PKV_ERR_INVALID, "shard %d out of range", i);
    *out = reinterpret_cast<pkv_index *>(sh->shards[i]);
    return PKV_OK;
}

int pkv_sharded_reserve(pkv_sharded *h, int64_t total_rows) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    std::lock_guard<std::mutex> lock(sh->mu);
    if (sh->rows > 0) return fail(PKV_ERR_INVALID, "reserve must precede the first append");
    if (total_rows < 1) return fail(PKV_ERR_INVALID, "total_rows must be >= 1");
    const int64_t S = (int64_t)sh->shards.size();
    int64_t per = (total_rows + S - 1) / S;
    per = (per + 63) / 64 * 64;
    sh->per_shard = per;
    for (int64_t i = 0; i < S; ++i) {
        pkv_index *ix = reinterpret_cast<pkv_index *>(sh->shards[i]);
        PKV_TRY(pkv_index_set_row_base(ix, i * per));
        int64_t want = total_rows - i * per;
        if (want > per) want = per;
        if (want > 0) PKV_TRY(pkv_index_reserve(ix, want));
    }
    return PKV_OK;
}

int pkv_sharded_append(pkv_sharded *h, const void *rows, const int64_t *row_ids, int64_t n) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    std::lock_guard<std::mutex> lock(sh->mu);
    if (n < 0) return fail(PKV_ERR_INVALID, "n must be >= 0");
    if (sh->per_shard == 0) return fail(PKV_ERR_NOT_READY, "pkv_sharded_reserve(total_rows) fixes the shard boundaries first");
    const size_t row_bytes = (size_t)sh->dim * (sh->dtype == PKV_F32 ? 4 : (sh->dtype == PKV_I8 ? 1 : 2));
    int64_t done = 0;
    while (done < n) {
        const int64_t s = sh->rows / sh->per_shard;
        if (s >= (int64_t)sh->shards.size()) return fail(PKV_ERR_INVALID, "more rows than pkv_sharded_reserve announced");
        int64_t room = (s + 1) * sh->per_shard - sh->rows;
        if (room > n - done) room = n - done;
        PKV_TRY(pkv_index_append(reinterpret_cast<pkv_index *>(sh->shards[s]), (const uint8_t *)rows + (size_t)done * row_bytes,
                                 row_ids ? row_ids + done : nullptr, room));
        done += room;
        sh->rows += room;
    }
    return PKV_OK;
}

int pkv_sharded_set_scale(pkv_sharded *h, const uint8_t *artifact, size_t len) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    for (Index *p : sh->shards) PKV_TRY(pkv_index_set_scale(reinterpret_cast<pkv_index *>(p), artifact, len));
    return PKV_OK;
}

int pkv_sharded_set_option(pkv_sharded *h, const char *name, int64_t value) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    for (Index *p : sh->shards) PKV_TRY(pkv_index_set_option(reinterpret_cast<pkv_index *>(p), name, value));
    return PKV_OK;
}

int pkv_sharded_seal(pkv_sharded *h) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    std::lock_guard<std::mutex> lock(sh->mu);
    for (Index *p : sh->shards) PKV_TRY(pkv_index_seal(reinterpret_cast<pkv_index *>(p)));
    return PKV_OK;
}

int pkv_sharded_rows(const pkv_sharded *h, int64_t *rows) {
    if (!h || !rows) return fail(PKV_ERR_INVALID, "NULL argument");
    *rows = reinterpret_cast<const Sharded *>(h)->rows;
    return PKV_OK;
}

int pkv_sharded_search(pkv_sharded *h, const void *queries, int nq, const pkv_search_params *params, int64_t *out_ids,
                       float *out_dist, int32_t *out_counts) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    if (!params) return fail(PKV_ERR_INVALID, "search params are NULL");
    Sharded *sh = reinterpret_cast<Sharded *>(h);
    if (nq < 0) return fail(PKV_ERR_INVALID, "nq must be >= 0");
    if (nq == 0) return PKV_OK;
    if (!queries || !out_ids || !out_dist || !out_counts) return fail(PKV_ERR_INVALID, "queries/outputs must not be NULL");
    if (params->k < 1 || params->k > PKV_MAX_K) return fail(PKV_ERR_INVALID, "k must be in 1..%d", PKV_MAX_K);
    if (params->bitmap && params->bitmap_stride_words != 0 && sh->shards.size() > 1)
        return fail(PKV_ERR_UNSUPPORTED, "per-query bitmaps are not sliced across shards; pass one shared bitmap");
    const int S = (int)sh->shards.size(), k = params->k;
    if (S == 1) return pkv_search(reinterpret_cast<pkv_index *>(sh->shards[0]), queries, nq, params, out_ids, out_dist, out_counts);

    // per-shard scans, concurrently: one host thread each (pkv_search is synchronous and re-entrant)
    const size_t per = (size_t)nq * k;
    std::vector<int64_t> h_ids((size_t)S * per);
    std::vector<float> h_dist((size_t)S * per);
    std::vector<int32_t> h_cnt((size_t)S * nq);
    std::vector<int> status(S, PKV_OK);
    std::vector<std::string> errors(S);
    std::vector<std::thread> workers;
    for (int i = 0; i < S; ++i) {
        workers.emplace_back([&, i] {
            pkv_search_params p = *params;
            if (p.bitmap) p.bitmap = params->bitmap + (size_t)i * (size_t)(sh->per_shard / 64);  // shard i's words
            pkv_index_info info;
            pkv_index *ix = reinterpret_cast<pkv_index *>(sh->shards[i]);
            pkv_index_get_info(ix, &info);
            if (info.rows == 0) {  // an empty trailing shard contributes nothing
                for (size_t e = 0; e < per; ++e) {
                    h_ids[i * per + e] = -1;
                    h_dist[i * per + e] = __builtin_nanf("");
                }
                for (int q = 0; q < nq; ++q) h_cnt[(size_t)i * nq + q] = 0;
                return;
            }
            status[i] = pkv_search(ix, queries, nq, &p, h_ids.data() + i * per, h_dist.data() + i * per,
                                   h_cnt.data() + (size_t)i * nq);
            if (status[i] != PKV_OK) errors[i] = pkv_last_error();
        });
    }
    for (auto &w : workers) w.join();
    for (int i = 0; i < S; ++i)
        if (status[i] != PKV_OK) return fail(status[i], "shard %d: %s", i, errors[i].c_str());

    // merge on the first shard's device: [S][nq][k] lists under (distance, shard, position) = (distance, global row)
    std::lock_guard<std::mutex> lock(sh->mu);
    DeviceGuard guard;
    PKV_TRY(guard.use(sh->devices[0]));
    if (!sh->stream) PKV_CUDA(cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking));
    if (sh->stage_entries < (size_t)S * per || sh->stage_queries < (size_t)nq) {
        cudaFree(sh->d_gathered);
        cudaFree(sh->d_ids);
        cudaFree(sh->d_dist);
        cudaFree(sh->d_counts);
        sh->d_gathered = nullptr;
        sh->d_ids = nullptr;
        sh->d_dist = nullptr;
        sh->d_counts = nullptr;
        sh->stage_entries = sh->stage_queries = 0;
        PKV_CUDA(cudaMalloc(&sh->d_gathered, (size_t)S * per * 12));
        PKV_CUDA(cudaMalloc((void **)&sh->d_ids, sizeof(int64_t) * (size_t)S * per));
        PKV_CUDA(cudaMalloc((void **)&sh->d_dist, sizeof(float) * (size_t)S * per));
        PKV_CUDA(cudaMalloc((void **)&sh->d_counts, sizeof(int32_t) * (size_t)nq));
        sh->stage_entries = (size_t)S * per;
        sh->stage_queries = (size_t)nq;
    }
    cudaStream_t s = sh->stream;
    PKV_CUDA(cudaMemcpyAsync(sh->d_ids, h_ids.data(), sizeof(int64_t) * (size_t)S * per, cudaMemcpyHostToDevice, s));
    PKV_CUDA(cudaMemcpyAsync(sh->d_dist, h_dist.data(), sizeof(float) * (size_t)S * per, cudaMemcpyHostToDevice, s));
    // merged lists overwrite the first shard's slice of the staging buffers only after the kernel has read them all:
    // write to the gathered scratch instead
    int64_t *m_ids = reinterpret_cast<int64_t *>(sh->d_gathered);
    float *m_dist = reinterpret_cast<float *>(m_ids + per);
    PKV_TRY(launch_merge(sh->d_ids, sh->d_dist, S, nq, k, m_ids, m_dist, sh->d_counts, s));
    PKV_CUDA(cudaMemcpyAsync(out_ids, m_ids, sizeof(int64_t) * per, cudaMemcpyDeviceToHost, s));
    PKV_CUDA(cudaMemcpyAsync(out_dist, m_dist, sizeof(float) * per, cudaMemcpyDeviceToHost, s));
    PKV_CUDA(cudaMemcpyAsync(out_counts, sh->d_counts, sizeof(int32_t) * (size_t)nq, cudaMemcpyDeviceToHost, s));
    PKV_CUDA(cudaStreamSynchronize(s));
    return PKV_OK;
}

// ================================================================ one process per GPU: NCCL communicator
int pkv_comm_unique_id(uint8_t *out, size_t len) {
    if (!out || len != 128) return fail(PKV_ERR_INVALID, "the id buffer must be exactly 128 bytes");
    Nccl &n = nccl();
    if (!n.lib || !n.error.empty()) return fail(PKV_ERR_UNSUPPORTED, "NCCL is unavailable: %s", n.error.c_str());
    ncclUniqueId id;
    PKV_NCCL(n.GetUniqueId(&id));
    memcpy(out, id.internal, 128);
    return PKV_OK;
}

int pkv_comm_create(int device, int rank, int nranks, const uint8_t *unique_id, size_t len, pkv_comm **out) {
    if (!out) return fail(PKV_ERR_INVALID, "out must not be NULL");
    *out = nullptr;
    if (!unique_id || len != 128) return fail(PKV_ERR_INVALID, "the unique id must be exactly 128 bytes");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(PKV_ERR_INVALID, "rank %d of %d", rank, nranks);
    Nccl &n = nccl();
    if (!n.lib || !n.error.empty()) return fail(PKV_ERR_UNSUPPORTED, "NCCL is unavailable: %s", n.error.c_str());
    PKV_USE_DEVICE(device);
    Comm *c = new (std::nothrow) Comm();
    if (!c) return fail(PKV_ERR_OOM, "out of host memory");
    c->device = device;
    c->rank = rank;
    c->nranks = nranks;
    ncclUniqueId id;
    memcpy(id.internal, unique_id, 128);
    const int r = n.CommInitRank(&c->comm, nranks, id, rank);
    if (r != NCCL_SUCCESS) {
        delete c;
        return fail(PKV_ERR_CUDA, "ncclCommInitRank failed: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    }
    *out = reinterpret_cast<pkv_comm *>(c);
    return PKV_OK;
}

int pkv_comm_destroy(pkv_comm *h) {
    if (!h) return PKV_OK;
    Comm *c = reinterpret_cast<Comm *>(h);
    DeviceGuard g;
    if (g.use(c->device) == PKV_OK) {
        cudaFree(c->d_packed);
        cudaFree(c->d_gathered);
        if (c->comm && nccl().CommDestroy) nccl().CommDestroy(c->comm);
    }
    delete c;
    return PKV_OK;
}

int pkv_comm_info(const pkv_comm *h, int *rank, int *nranks) {
    if (!h) return fail(PKV_ERR_INVALID, "handle is NULL");
    const Comm *c = reinterpret_cast<const Comm *>(h);
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    return PKV_OK;
}

int pkv_search_sharded_device(pkv_index *shard, pkv_comm *comm, const void *d_queries, int nq,
                              const pkv_search_params *params, int64_t *d_out_ids, float *d_out_dist, int32_t *d_out_counts,
                              void *stream) {
    if (!shard || !comm) return fail(PKV_ERR_INVALID, "NULL handle");
    if (!params) return fail(PKV_ERR_INVALID, "search params are NULL");
    Comm *c = reinterpret_cast<Comm *>(comm);
    if (nq <= 0 || params->k < 1 || params->k > PKV_MAX_K) return fail(PKV_ERR_INVALID, "bad nq / k");
    Nccl &n = nccl();
    PKV_USE_DEVICE(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t entries = (size_t)nq * params->k;
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->packed_bytes < entries * 12) {
        cudaFree(c->d_packed);
        cudaFree(c->d_gathered);
        c->d_packed = c->d_gathered = nullptr;
        c->packed_bytes = 0;
        PKV_CUDA(cudaMalloc(&c->d_packed, (entries * 12 + 15) / 16 * 16 + entries * 12));  // packed list + this rank's (ids, dist) staging
        PKV_CUDA(cudaMalloc(&c->d_gathered, entries * 12 * (size_t)c->nranks));
        c->packed_bytes = entries * 12;
    }
    // this shard's top-k (staged: the caller's output buffers receive the MERGED lists)
    int64_t *l_ids = reinterpret_cast<int64_t *>((uint8_t *)c->d_packed + (entries * 12 + 15) / 16 * 16);
    float *l_dist = reinterpret_cast<float *>(l_ids + entries);
    PKV_TRY(pkv_search_device(shard, d_queries, nq, params, l_ids, l_dist, d_out_counts, stream));
    // pack (12-byte entries) -> ONE all-gather over NVLink -> merge of the gathered buffer; all enqueued, no host round trip
    PKV_TRY(launch_pack_topk(l_ids, l_dist, (int64_t)entries, c->d_packed, s));
    PKV_NCCL(n.AllGather(c->d_packed, c->d_gathered, entries * 12, NCCL_UINT8, c->comm, s));
    PKV_TRY(launch_merge_packed(c->d_gathered, c->nranks, nq, params->k, d_out_ids, d_out_dist, d_out_counts, s));
    // the exchange is only ENQUEUED: the outputs are complete once `stream` has been synchronised (the next search on
    // the same stream orders behind it by itself)
    return PKV_OK;
}

}  // extern "C"
