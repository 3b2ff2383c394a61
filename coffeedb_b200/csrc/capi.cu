// C ABI (include/coffeedb_b200.h) over the device index.  Every entry point converts C++ exceptions into a
// status code + thread-local message; the three conditions the reference throws on carry the reference's
// exact text (src/index.cpp:196,199,240).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "../host/micro_batcher.hpp"
#include "filter.cuh"
#include "index.cuh"
#include "locate.cuh"
#include "persist.cuh"
#include "sharded.cuh"
#include "verify.cuh"

namespace cdb {

static thread_local std::string g_last_error;
std::atomic<unsigned long long> g_launches{0};
thread_local LocateStats g_locate_stats;

Index::~Index() { free_device(); }

void Index::free_device() {
    if (d_sa) cudaFree(d_sa);
    if (own_text) cudaFree(own_text);
    if (own_off) cudaFree(own_off);
    if (own_ids) cudaFree(own_ids);
    if (d_ptab) cudaFree(d_ptab);
    if (d_rank_tab) cudaFree(d_rank_tab);
    if (d_ids_by_rank) cudaFree(d_ids_by_rank);
    if (d_sa_rank) cudaFree(d_sa_rank);
    d_sa_rank = nullptr;
    listing[0].reset();
    listing[1].reset();
    listing_state[0] = listing_state[1] = 0;
    listing_miss[0] = listing_miss[1] = 0;
    listing_skip[0] = listing_skip[1] = 0;
    d_ptab = nullptr;
    d_rank_tab = nullptr;
    d_ids_by_rank = nullptr;
    ids_order = -1;
    pt_k = pt_b = 0;
    d_sa = own_text = own_off = own_ids = nullptr;
    d_text = nullptr;
    d_off = nullptr;
    d_ids = nullptr;
    built = false;
}

// Pinned host buffers are expensive to create (~0.3 s/GB), so freed result buffers are kept for re-use.
class PinnedPool {
public:
    void* get(size_t bytes, size_t* cap_out) {
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = free_.lower_bound(bytes);
            if (it != free_.end() && it->first <= bytes * 2 + (1 << 20)) {
                void* p = it->second;
                *cap_out = it->first;
                held_ -= it->first;
                free_.erase(it);
                return p;
            }
        }
        void* p = nullptr;
        size_t cap = bytes < 4096 ? 4096 : bytes;
        CDB_CUDA(cudaHostAlloc(&p, cap, cudaHostAllocDefault));
        *cap_out = cap;
        return p;
    }
    // Keeps at most kMaxBytes / kMaxEntries of idle buffers (CDB_PINNED_POOL_MB overrides the byte cap): beyond that the
    // largest idle buffers go back to the driver, so a server with varying batch sizes does not pile up page-locked memory.
    void put(void* p, size_t cap) {
        std::vector<void*> drop;
        {
            std::lock_guard<std::mutex> g(mu_);
            free_.emplace(cap, p);
            held_ += cap;
            while (!free_.empty() && (held_ > max_bytes() || free_.size() > kMaxEntries)) {
                auto it = std::prev(free_.end());
                held_ -= it->first;
                drop.push_back(it->second);
                free_.erase(it);
            }
        }
        for (void* q : drop) cudaFreeHost(q);
    }
    void trim() {
        std::lock_guard<std::mutex> g(mu_);
        for (auto& kv : free_) cudaFreeHost(kv.second);
        free_.clear();
        held_ = 0;
    }
    ~PinnedPool() {
        for (auto& kv : free_) cudaFreeHost(kv.second);
    }

private:
    static constexpr size_t kMaxEntries = 64;
    static size_t max_bytes() {
        static const size_t v = [] {
            const char* e = getenv("CDB_PINNED_POOL_MB");
            return e ? (size_t)atoll(e) << 20 : (size_t)32 << 30;  // the 10 GB configuration's result is 12.9 GB
        }();
        return v;
    }
    std::mutex mu_;
    std::multimap<size_t, void*> free_;
    size_t held_ = 0;
};
static PinnedPool g_pinned;

struct HostResultOwner {
    void* row_off;
    size_t row_cap;
    void* pairs;
    size_t pairs_cap;
    // cdb_query: the result is a view of one row of a shared batch result (kept alive by `shared`)
    void* shared = nullptr;  // std::shared_ptr<QueryBatchResult>*
    int64_t view_off[2] = {0, 0};
    bool plain = false;  // row_off / pairs come from malloc (small batches: rows are copied by the CPU, nothing is DMA'd into them)
};

// One coalesced device batch of cdb_query callers: a host cdb_result, released when its last row view goes.
struct QueryBatchResult {
    cdb_result r{};
    const int64_t* row_off = nullptr;
    const int64_t* pairs = nullptr;
    ~QueryBatchResult() { cdb_result_free(&r); }
};
struct QueryBackend {
    const cdb_index* h;
    std::shared_ptr<QueryBatchResult> operator()(const std::string& bytes, const std::vector<int64_t>& off) const {
        auto res = std::make_shared<QueryBatchResult>();
        const cdb_status s = cdb_locate_batch(h, bytes.data(), off.data(), (int64_t)off.size() - 1, &res->r);
        if (s != CDB_OK) throw Error(s, cdb_last_error());  // rethrown in every member of the batch
        res->row_off = res->r.row_off;
        res->pairs = res->r.pairs;
        return res;
    }
};
using QueryBatcher = coffeedb_b200::micro_batcher<QueryBackend, QueryBatchResult>;

static int env_int(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    if (!e) return dflt;
    return std::max(lo, std::min(hi, atoi(e)));
}

static QueryBatcher& query_batcher(const Index* ix) {
    std::lock_guard<std::mutex> lk(ix->batcher_mu);
    if (!ix->batcher) {
        // CDB_QUERY_MAX_BATCH keywords per device batch, CDB_QUERY_IN_FLIGHT batches on the device at a time (the
        // second one overlaps its kernels with the first one's read-back), CDB_QUERY_LINGER_US extra wait of a
        // batch leader for company (0: an idle device serves a lone caller at once)
        ix->batcher = std::make_shared<QueryBatcher>(
            QueryBackend{reinterpret_cast<const cdb_index*>(ix)}, (size_t)env_int("CDB_QUERY_MAX_BATCH", 1 << 16, 1, 1 << 24),
            env_int("CDB_QUERY_IN_FLIGHT", 2, 1, 64), std::chrono::microseconds(env_int("CDB_QUERY_LINGER_US", 0, 0, 1000000)));
    }
    return *static_cast<QueryBatcher*>(ix->batcher.get());
}

struct DeviceSetter {
    int prev = -1;
    explicit DeviceSetter(int dev) {
        cudaGetDevice(&prev);
        if (dev >= 0 && dev != prev) CDB_CUDA(cudaSetDevice(dev));
    }
    ~DeviceSetter() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// The reference never updates a built index either: a new one is filled and built, then swapped in
// (src/database.cpp:170-172, 276-281).
static const char* const kAfterBuild =
    "the index has been built and its host staging copy released: add to a new index (or create this one with "
    "keep_host_copy = 1)";

ThreadCtx& thread_ctx(int device) {
    static thread_local std::map<int, ThreadCtx> ctxs;
    ThreadCtx& c = ctxs[device];
    if (!c.stream) {
        c.device = device;
        CDB_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        for (auto& e : c.ev) CDB_CUDA(cudaEventCreate(&e));
    }
    return c;
}

static void require_device() {
    int cnt = 0;
    cudaError_t e = cudaGetDeviceCount(&cnt);
    if (e != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        throw Error(CDB_ERR_CUDA, "no CUDA device available: coffeedb_b200 has no CPU fallback");
    }
}

static void keep_pool_memory(int dev) {
    // keep freed temporaries in the stream-ordered pool instead of returning them to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
}

}  // namespace cdb

using namespace cdb;

#define CDB_TRY try {
#define CDB_CATCH                                            \
    }                                                        \
    catch (const cdb::Error& e) {                            \
        g_last_error = e.what();                             \
        return e.code;                                       \
    }                                                        \
    catch (const std::bad_alloc&) {                          \
        g_last_error = "out of host memory";                 \
        return CDB_ERR_NOMEM;                                \
    }                                                        \
    catch (const std::exception& e) {                        \
        g_last_error = e.what();                             \
        return CDB_ERR_ARG;                                  \
    }

extern "C" {

const char* cdb_last_error(void) { return g_last_error.c_str(); }
const char* cdb_version(void) { return "coffeedb_b200 0.1 (sm_100a)"; }

uint64_t cdb_launch_count(void) { return g_launches.load(); }

void cdb_trim(void) { g_pinned.trim(); }

void cdb_last_locate_stats(double* ms6, int64_t* counts4) {
    const LocateStats& s = g_locate_stats;
    if (ms6) {
        ms6[0] = s.search_ms; ms6[1] = s.gather_ms; ms6[2] = s.large_ms;
        ms6[3] = s.tail_ms; ms6[4] = s.translate_ms; ms6[5] = s.total_ms;
    }
    if (counts4) {
        counts4[0] = s.npat; counts4[1] = s.total_pairs; counts4[2] = s.total_occ; counts4[3] = s.nlarge;
    }
}

void cdb_last_locate_stats_ex(double* ms8, int64_t* counts8) {
    const LocateStats& s = g_locate_stats;
    if (ms8) {
        cdb_last_locate_stats(ms8, nullptr);
        ms8[6] = s.listing_ms;
        ms8[7] = 0;
    }
    if (counts8) {
        cdb_last_locate_stats(nullptr, counts8);
        counts8[4] = s.nlisted; counts8[5] = s.listed_pairs; counts8[6] = counts8[7] = 0;
    }
}

int cdb_device_count(void) {
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return cnt;
}

cdb_status cdb_create(const cdb_options* opts, cdb_index** out) {
    CDB_TRY
    if (!out) throw Error(CDB_ERR_ARG, "cdb_create: out is NULL");
    Index* ix = new Index();
    if (opts) ix->opt = *opts;
    else {
        ix->opt.device = -1;
        ix->opt.compat_signed = 1;
        ix->opt.workspace_bytes = 0;
        ix->opt.keep_host_copy = 0;
    }
    ix->device = ix->opt.device;
    ix->h_text.set_device(ix->opt.device);
    *out = reinterpret_cast<cdb_index*>(ix);
    return CDB_OK;
    CDB_CATCH
}

void cdb_destroy(cdb_index* h) {
    if (!h) return;
    Index* ix = reinterpret_cast<Index*>(h);
    int prev = -1;
    cudaGetDevice(&prev);
    if (ix->device >= 0 && ix->built) cudaSetDevice(ix->device);
    delete ix;
    if (prev >= 0) cudaSetDevice(prev);
}

cdb_status cdb_add(cdb_index* h, int64_t id, const void* value, int64_t len) {
    CDB_TRY
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix || len < 0 || (len > 0 && !value)) throw Error(CDB_ERR_ARG, "cdb_add: bad argument");
    if (ix->host_dropped) throw Error(CDB_ERR_STATE, kAfterBuild);
    const size_t text0 = ix->h_text.size(), ids0 = ix->h_ids.size(), off0 = ix->h_off.size();
    try {
        ix->h_text.append(static_cast<const u8*>(value), (size_t)len);
        ix->h_ids.push_back(id);
        ix->h_off.push_back((i64)ix->h_text.size());
    } catch (...) {
        ix->h_text.truncate(text0);
        ix->h_ids.resize(ids0);
        ix->h_off.resize(off0);
        throw;
    }
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_add_many(cdb_index* h, const int64_t* ids, const void* text, const int64_t* doc_off, int64_t nd) {
    CDB_TRY
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix || nd < 0 || (nd > 0 && (!ids || !doc_off))) throw Error(CDB_ERR_ARG, "cdb_add_many: bad argument");
    if (ix->host_dropped) throw Error(CDB_ERR_STATE, kAfterBuild);
    if (nd == 0) return CDB_OK;
    // validate everything before the staging vectors are touched: a rejected call leaves the index unchanged
    if (doc_off[0] < 0) throw Error(CDB_ERR_ARG, "cdb_add_many: doc_off[0] must be >= 0");
    for (i64 d = 1; d <= nd; ++d)
        if (doc_off[d] < doc_off[d - 1]) throw Error(CDB_ERR_ARG, "cdb_add_many: doc_off must be non-decreasing");
    if (doc_off[nd] > doc_off[0] && !text) throw Error(CDB_ERR_ARG, "cdb_add_many: text is NULL");
    const u8* p = static_cast<const u8*>(text);
    const size_t text0 = ix->h_text.size(), ids0 = ix->h_ids.size(), off0 = ix->h_off.size();
    try {
        const i64 base = (i64)text0 - doc_off[0];
        if (doc_off[nd] > doc_off[0]) ix->h_text.append(p + doc_off[0], (size_t)(doc_off[nd] - doc_off[0]));
        ix->h_ids.insert(ix->h_ids.end(), ids, ids + nd);
        ix->h_off.reserve(off0 + (size_t)nd);
        for (i64 d = 1; d <= nd; ++d) ix->h_off.push_back(base + doc_off[d]);
    } catch (...) {  // out of host memory half way: roll back
        ix->h_text.truncate(text0);
        ix->h_ids.resize(ids0);
        ix->h_off.resize(off0);
        throw;
    }
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_staging_stats(const cdb_index* h, int64_t* staged, int64_t* on_device) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix) throw Error(CDB_ERR_ARG, "cdb_staging_stats: index is NULL");
    if (staged) *staged = ix->built ? ix->n : (i64)ix->h_text.size();
    if (on_device) *on_device = ix->built ? ix->staged_on_device : (i64)ix->h_text.uploaded_bytes();
    return CDB_OK;
    CDB_CATCH
}

static void build_from_staging(Index* ix, const SavedArraySource* saved) {
    if (ix->host_dropped) throw Error(CDB_ERR_STATE, kAfterBuild);
    require_device();
    if (ix->device < 0) CDB_CUDA(cudaGetDevice(&ix->device));
    DeviceSetter ds(ix->device);
    keep_pool_memory(ix->device);
    ix->free_device();
    const i64 nd = (i64)ix->h_ids.size();
    const i64 n = (i64)ix->h_text.size();
    cudaStream_t st;
    CDB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    try {
        CDB_CUDA(cudaMalloc(&ix->own_text, (size_t)(n + kTextPad)));
        CDB_CUDA(cudaMalloc(&ix->own_off, (size_t)(nd + 1) * 8));
        CDB_CUDA(cudaMalloc(&ix->own_ids, (size_t)(nd ? nd : 1) * 8));
        CDB_CUDA(cudaMemsetAsync((u8*)ix->own_text + n, 0, kTextPad, st));
        // the chunks that filled up during cdb_add are already on the device: stitched together device-to-device; the
        // last, partly filled one comes from the host
        ix->staged_on_device = (i64)ix->h_text.uploaded_bytes();
        ix->h_text.assemble(ix->own_text, st);
        CDB_CUDA(cudaMemcpyAsync(ix->own_off, ix->h_off.data(), (size_t)(nd + 1) * 8, cudaMemcpyHostToDevice, st));
        if (nd) CDB_CUDA(cudaMemcpyAsync(ix->own_ids, ix->h_ids.data(), (size_t)nd * 8, cudaMemcpyHostToDevice, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        ix->h_text.release_device_copies();  // before the build sizes its workspace from the free memory
        ix->d_text = (const u8*)ix->own_text;
        ix->d_off = (const i64*)ix->own_off;
        ix->d_ids = (const i64*)ix->own_ids;
        ix->nd = nd;
        build_index(*ix, st, saved);
    } catch (...) {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
        ix->free_device();
        throw;
    }
    cudaStreamDestroy(st);
    if (!ix->opt.keep_host_copy) {
        ix->h_text.clear();
        ix->host_dropped = true;
    }
}

cdb_status cdb_build(cdb_index* h) {
    CDB_TRY
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix) throw Error(CDB_ERR_ARG, "cdb_build: index is NULL");
    build_from_staging(ix, nullptr);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_build_or_load(cdb_index* h, const char* path, int32_t* loaded) {
    CDB_TRY
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix || !path) throw Error(CDB_ERR_ARG, "cdb_build_or_load: bad argument");
    SavedArrayFile src(path);
    build_from_staging(ix, &src);
    if (loaded) *loaded = ix->loaded_from_file ? 1 : 0;
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_save(const cdb_index* h, const char* path) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !path) throw Error(CDB_ERR_ARG, "cdb_save: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    DeviceSetter ds(ix->device);
    save_index(*ix, path, thread_ctx(ix->device).stream);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_build_device(cdb_index* h, const void* d_text, const int64_t* d_doc_off, const int64_t* d_ids, int64_t nd,
                            void* stream) {
    CDB_TRY
    Index* ix = reinterpret_cast<Index*>(h);
    if (!ix || nd < 0 || !d_doc_off) throw Error(CDB_ERR_ARG, "cdb_build_device: bad argument");
    if (((uintptr_t)d_text & 15) != 0) throw Error(CDB_ERR_ARG, "cdb_build_device: d_text must be 16-byte aligned");
    require_device();
    if (ix->device < 0) CDB_CUDA(cudaGetDevice(&ix->device));
    DeviceSetter ds(ix->device);
    keep_pool_memory(ix->device);
    ix->free_device();
    ix->d_text = (const u8*)d_text;
    ix->d_off = d_doc_off;
    ix->d_ids = d_ids;
    ix->nd = nd;
    try {
        build_index(*ix, (cudaStream_t)stream);
    } catch (...) {
        cudaStreamSynchronize((cudaStream_t)stream);
        ix->free_device();
        throw;
    }
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_info(const cdb_index* h, int64_t* n, int64_t* nd, int32_t* width, int32_t* bits, uint64_t* mask) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    if (n) *n = ix->n;
    if (nd) *nd = ix->nd;
    if (width) *width = ix->width;
    if (bits) *bits = ix->bits1;
    if (mask) *mask = ix->mask;
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_prefix_directory(const cdb_index* h, int32_t* symbols, int32_t* bits_per_symbol, int64_t* entries) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    const bool on = ix->d_ptab != nullptr;
    if (symbols) *symbols = on ? ix->pt_k : 0;
    if (bits_per_symbol) *bits_per_symbol = on ? ix->pt_b : 0;
    if (entries) *entries = on ? ((i64)1 << (ix->pt_b * ix->pt_k)) : 0;
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_listing_info(const cdb_index* h, int32_t order, int32_t* present, int32_t* hi_bytes, int64_t* bytes, double* build_ms) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    if (order != 0 && order != 1) throw Error(CDB_ERR_ARG, "cdb_listing_info: order must be 0 or 1");
    std::shared_ptr<Listing> L;
    {
        std::lock_guard<std::mutex> lk(ix->listing_mu);
        L = ix->listing[order];
        // ids that ascend with the doc index: the doc-order listing serves both orders
        if (!L && order == 1 && ix->ids_order == 1) L = ix->listing[0];
    }
    if (present) *present = L ? 1 : 0;
    if (hi_bytes) *hi_bytes = L ? L->hw : 0;
    if (bytes) *bytes = L ? (i64)L->bytes : 0;
    if (build_ms) *build_ms = L ? L->build_ms : 0.0;
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_export_sa(const cdb_index* h, void* buf, int64_t buf_bytes) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    const i64 need = ix->n * ix->width;
    if (buf_bytes < need || (need && !buf)) throw Error(CDB_ERR_ARG, "cdb_export_sa: buffer too small");
    DeviceSetter ds(ix->device);
    if (need) CDB_CUDA(cudaMemcpy(buf, ix->d_sa, (size_t)need, cudaMemcpyDeviceToHost));
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_sa_device_ptr(const cdb_index* h, const void** d_sa) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !ix->built || !d_sa) throw Error(CDB_ERR_STATE, "index has not been built");
    *d_sa = ix->d_sa;
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_verify_sa(const cdb_index* h, int64_t* out8) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out8) throw Error(CDB_ERR_ARG, "cdb_verify_sa: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    DeviceSetter ds(ix->device);
    verify_index(*ix, thread_ctx(ix->device).stream, out8);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_compare_sa(const cdb_index* h, const void* other_sa, int64_t other_bytes, int64_t* out3) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out3) throw Error(CDB_ERR_ARG, "cdb_compare_sa: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    if (other_bytes != ix->n * ix->width || (other_bytes && !other_sa))
        throw Error(CDB_ERR_ARG, "cdb_compare_sa: the other array must hold n elements of the index's width");
    DeviceSetter ds(ix->device);
    compare_index_sa(*ix, other_sa, thread_ctx(ix->device).stream, out3);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_build_stats(const cdb_index* h, double* total_ms, double* sort_ms, int64_t* rounds, int64_t* chunks) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    if (total_ms) *total_ms = ix->build_ms;
    if (sort_ms) *sort_ms = ix->sort_ms;
    if (rounds) *rounds = ix->rounds;
    if (chunks) *chunks = ix->chunks;
    return CDB_OK;
    CDB_CATCH
}

// ---- one index over several devices of this process (sharded.cu) -----------------------------------------------------------
cdb_status cdb_sharded_create(const int32_t* devices, int32_t ndev, const cdb_options* opts, cdb_sharded** out) {
    CDB_TRY
    if (!out || !devices || ndev < 1 || ndev > 64) throw Error(CDB_ERR_ARG, "cdb_sharded_create: bad argument");
    *out = reinterpret_cast<cdb_sharded*>(sharded_create(devices, ndev, opts));
    return CDB_OK;
    CDB_CATCH
}

void cdb_sharded_destroy(cdb_sharded* s) { sharded_destroy(reinterpret_cast<ShardedIndex*>(s)); }

cdb_status cdb_sharded_add_many(cdb_sharded* h, const int64_t* ids, const void* text, const int64_t* doc_off, int64_t nd) {
    CDB_TRY
    if (!h || nd < 0 || (nd > 0 && (!ids || !doc_off))) throw Error(CDB_ERR_ARG, "cdb_sharded_add_many: bad argument");
    if (nd == 0) return CDB_OK;
    if (doc_off[0] < 0) throw Error(CDB_ERR_ARG, "cdb_sharded_add_many: doc_off[0] must be >= 0");
    for (i64 d = 1; d <= nd; ++d)
        if (doc_off[d] < doc_off[d - 1]) throw Error(CDB_ERR_ARG, "cdb_sharded_add_many: doc_off must be non-decreasing");
    if (doc_off[nd] > doc_off[0] && !text) throw Error(CDB_ERR_ARG, "cdb_sharded_add_many: text is NULL");
    sharded_add_many(reinterpret_cast<ShardedIndex*>(h), ids, static_cast<const u8*>(text), doc_off, nd);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_sharded_add(cdb_sharded* h, int64_t id, const void* value, int64_t len) {
    const int64_t off[2] = {0, len};
    if (len < 0) {
        g_last_error = "cdb_sharded_add: bad argument";
        return CDB_ERR_ARG;
    }
    return cdb_sharded_add_many(h, &id, value, off, 1);
}

cdb_status cdb_sharded_build(cdb_sharded* h) {
    CDB_TRY
    if (!h) throw Error(CDB_ERR_ARG, "cdb_sharded_build: index is NULL");
    require_device();
    sharded_build(reinterpret_cast<ShardedIndex*>(h));
    return CDB_OK;
    CDB_CATCH
}

int32_t cdb_sharded_count(const cdb_sharded* h) { return h ? sharded_count(reinterpret_cast<const ShardedIndex*>(h)) : 0; }

cdb_status cdb_sharded_shard(const cdb_sharded* h, int32_t g, cdb_index** shard, int64_t* doc_begin, int64_t* doc_end) {
    CDB_TRY
    if (!h || !shard) throw Error(CDB_ERR_ARG, "cdb_sharded_shard: bad argument");
    *shard = sharded_shard(reinterpret_cast<const ShardedIndex*>(h), g, doc_begin, doc_end);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_sharded_locate_batch(const cdb_sharded* h, const void* pat, const int64_t* pat_off, int64_t npat, cdb_result* out) {
    CDB_TRY
    if (!h || !out || npat < 0 || (npat > 0 && !pat_off)) throw Error(CDB_ERR_ARG, "cdb_sharded_locate_batch: bad argument");
    std::memset(out, 0, sizeof(*out));
    for (i64 q = 0; q < npat; ++q)  // src/index.cpp:239-241
        if (pat_off[q + 1] <= pat_off[q]) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");
    HostResultOwner* own = new HostResultOwner{nullptr, 0, nullptr, 0};
    i64 total_pairs = 0, total_occ = 0;
    try {
        sharded_locate(const_cast<ShardedIndex*>(reinterpret_cast<const ShardedIndex*>(h)), static_cast<const u8*>(pat), pat_off, npat,
                       [&](i64 total, i64** row_off, i64** pairs) {
                           own->row_off = g_pinned.get((size_t)(npat + 1) * 8, &own->row_cap);
                           own->pairs = g_pinned.get((size_t)(total ? total : 1) * 16, &own->pairs_cap);
                           *row_off = (i64*)own->row_off;
                           *pairs = (i64*)own->pairs;
                       },
                       &total_pairs, &total_occ);
    } catch (...) {
        if (own->row_off) g_pinned.put(own->row_off, own->row_cap);
        if (own->pairs) g_pinned.put(own->pairs, own->pairs_cap);
        delete own;
        throw;
    }
    out->npat = npat;
    out->total_pairs = total_pairs;
    out->total_occurrences = total_occ;
    out->row_off = (const i64*)own->row_off;
    out->pairs = (const i64*)own->pairs;
    out->_owner = own;
    return CDB_OK;
    CDB_CATCH
}

// ---- filter() (filter.cu) ------------------------------------------------------------------------------------------------
cdb_status cdb_numeric_create(int32_t kind, const int64_t* ids, const void* values, int64_t n, int32_t device, cdb_numeric** out) {
    CDB_TRY
    if (!out || n < 0 || (kind != 0 && kind != 1) || (n > 0 && (!ids || !values))) throw Error(CDB_ERR_ARG, "cdb_numeric_create: bad argument");
    require_device();
    int dev = device;
    if (dev < 0) CDB_CUDA(cudaGetDevice(&dev));
    DeviceSetter ds(dev);
    keep_pool_memory(dev);
    *out = reinterpret_cast<cdb_numeric*>(numeric_create(kind, ids, values, n, dev, thread_ctx(dev).stream));
    return CDB_OK;
    CDB_CATCH
}

void cdb_numeric_destroy(cdb_numeric* c) {
    if (!c) return;
    NumericIndex* ni = reinterpret_cast<NumericIndex*>(c);
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(ni->device);
    delete ni;
    if (prev >= 0) cudaSetDevice(prev);
}

cdb_status cdb_numeric_query(const cdb_numeric* c, const int64_t lo[2], const int64_t hi[2], cdb_result* out) {
    CDB_TRY
    const NumericIndex* ni = reinterpret_cast<const NumericIndex*>(c);
    if (!ni || !lo || !hi || !out) throw Error(CDB_ERR_ARG, "cdb_numeric_query: bad argument");
    std::memset(out, 0, sizeof(*out));
    DeviceSetter ds(ni->device);
    cudaStream_t st = thread_ctx(ni->device).stream;
    i64 b = 0, e = 0;
    numeric_bounds(*ni, lo, hi, st, &b, &e);
    const i64 m = e - b;
    HostResultOwner* own = new HostResultOwner{nullptr, 0, nullptr, 0};
    try {
        own->row_off = g_pinned.get(16, &own->row_cap);
        own->pairs = g_pinned.get((size_t)(m ? m : 1) * 16, &own->pairs_cap);
        i64* pr = (i64*)own->pairs;
        // ids land in the upper half of the buffer, then spread into (id, 0) pairs front to back
        i64* tmp = pr + m;
        if (m) CDB_CUDA(cudaMemcpyAsync(tmp, ni->vid + b, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        for (i64 i = 0; i < m; ++i) {
            const i64 id = tmp[i];
            pr[2 * i] = id;
            pr[2 * i + 1] = 0;
        }
        ((i64*)own->row_off)[0] = 0;
        ((i64*)own->row_off)[1] = m;
    } catch (...) {
        if (own->row_off) g_pinned.put(own->row_off, own->row_cap);
        if (own->pairs) g_pinned.put(own->pairs, own->pairs_cap);
        delete own;
        throw;
    }
    out->npat = 1;
    out->total_pairs = m;
    out->row_off = (const i64*)own->row_off;
    out->pairs = (const i64*)own->pairs;
    out->_owner = own;
    return CDB_OK;
    CDB_CATCH
}

struct FilterOwner {
    void* p[3] = {nullptr, nullptr, nullptr};  // row_off, pairs, matched
    size_t cap[3] = {0, 0, 0};
    void release() {
        for (int i = 0; i < 3; ++i)
            if (p[i]) g_pinned.put(p[i], cap[i]);
    }
};

// fin_off of one part of a pipelined batch -> offsets inside the whole batch's result
__global__ void shift_offsets_kernel(u64* off, i64 n, u64 base) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] += base;
}

cdb_status cdb_filter(const cdb_filter_batch* b, cdb_filter_result* out) {
    CDB_TRY
    if (!b || !out || b->nreq < 0 || (b->nreq > 0 && !b->req_term_off) || (b->nkeys > 0 && !b->keys))
        throw Error(CDB_ERR_ARG, "cdb_filter: bad argument");
    std::memset(out, 0, sizeof(*out));
    require_device();
    int dev = filter_device_of(*b);
    if (dev < 0) CDB_CUDA(cudaGetDevice(&dev));
    DeviceSetter ds(dev);
    ThreadCtx& tc = thread_ctx(dev);
    cudaStream_t st = tc.stream;
    const i64 nreq = b->nreq;
    const bool dbg = getenv("CDB_DEBUG_TIMING") != nullptr;
    // Batches whose result size is bounded by their spans CAN be cut into parts that go through the device one after the
    // other, the copy of part k's slices to the host (its own stream) running under the kernels of part k+1
    // (CDB_FILTER_PARTS > 1).  Measured at cfg3 (10^6 requests, 0.5 GB of slices = 9.3 ms over PCIe): 4 parts take 27.3 ms
    // against 23.3 ms in one piece — every part pays its own upload, keyword packing, scans and synchronisations
    // (about 3 ms), more than the copy it hides — so one part is the default.
    int nparts = 1;
    u64 bound = 0;
    i64 part_min = (i64)1 << 17;
    if (const char* e = getenv("CDB_FILTER_PART_MIN")) part_min = std::max<i64>(2, atoll(e));  // tests: small batches in parts
    const char* env_parts = getenv("CDB_FILTER_PARTS");  // (the walk over the spans below costs a millisecond per 10^6 requests)
    if (b->span && nreq >= part_min && env_parts && atoi(env_parts) > 1) {
        bool ok = true;
        for (i64 r = 0; r < nreq && ok; ++r) {
            const i64 s0 = b->span[2 * r] < 0 ? 0 : b->span[2 * r], s1 = b->span[2 * r + 1];
            if (s1 > s0) bound += (u64)(s1 - s0);
            ok = bound <= ((u64)1 << 28);  // 4 GB of pairs
        }
        nparts = ok ? std::max(1, std::min(16, atoi(env_parts))) : 1;
        if ((i64)nparts > nreq) nparts = 1;
    }
    FilterOwner* own = new FilterOwner();
    std::vector<std::unique_ptr<FilterOut>> parts;
    std::vector<std::vector<i64>> rto((size_t)nparts);
    try {
        const auto t0 = std::chrono::steady_clock::now();
        own->p[0] = g_pinned.get((size_t)(nreq + 1) * 8, &own->cap[0]);
        own->p[2] = g_pinned.get((size_t)(nreq ? nreq : 1) * 8, &own->cap[2]);
        i64* ro = (i64*)own->p[0];
        i64* pr = nullptr;
        if (nparts > 1) {
            own->p[1] = g_pinned.get((size_t)(bound ? bound : 1) * 16, &own->cap[1]);
            pr = (i64*)own->p[1];
        }
        u64 total = 0;
        std::vector<i64> part_begin((size_t)nparts + 1, nreq);
        for (int k = 0; k < nparts; ++k) part_begin[k] = nreq / nparts * k;
        for (int k = 0; k < nparts; ++k) {
            const i64 r0 = part_begin[k], n = part_begin[k + 1] - r0;
            cdb_filter_batch sub = *b;
            if (nparts > 1) {  // requests [r0, r0 + n): their terms, rebased offsets, their rows of corr_range / span
                const i64 t0k = b->req_term_off[r0];
                rto[k].resize((size_t)n + 1);
                for (i64 r = 0; r <= n; ++r) rto[k][r] = b->req_term_off[r0 + r] - t0k;
                sub.terms = b->terms + t0k;
                sub.req_term_off = rto[k].data();
                sub.nreq = n;
                sub.corr_range = b->corr_range ? b->corr_range + 2 * r0 : nullptr;
                sub.span = b->span + 2 * r0;
            }
            parts.emplace_back(new FilterOut());
            FilterOut& fo = *parts.back();
            filter_batch_device(sub, st, fo);
            if (nparts == 1) {
                own->p[1] = g_pinned.get((size_t)(fo.total_fin ? fo.total_fin : 1) * 16, &own->cap[1]);
                pr = (i64*)own->p[1];
            }
            if (total && n) {
                shift_offsets_kernel<<<(unsigned)ceil_div(n + 1, 256), 256, 0, st>>>(fo.fin_off.p, n + 1, total);
                CDB_LAUNCH_CHECK();
            }
            cudaStream_t cs = st;
            if (nparts > 1) {
                if (!tc.copy_stream) CDB_CUDA(cudaStreamCreateWithFlags(&tc.copy_stream, cudaStreamNonBlocking));
                cs = tc.copy_stream;
                CDB_CUDA(cudaEventRecord(tc.ev[ThreadCtx::kEvents - 1], st));
                CDB_CUDA(cudaStreamWaitEvent(cs, tc.ev[ThreadCtx::kEvents - 1], 0));
            }
            if (n) CDB_CUDA(cudaMemcpyAsync(ro + r0, fo.fin_off.p, (size_t)n * 8, cudaMemcpyDeviceToHost, cs));
            if (fo.total_fin) CDB_CUDA(cudaMemcpyAsync(pr + 2 * total, fo.fin.p, (size_t)fo.total_fin * 16, cudaMemcpyDeviceToHost, cs));
            if (n) CDB_CUDA(cudaMemcpyAsync((i64*)own->p[2] + r0, fo.matched.p, (size_t)n * 8, cudaMemcpyDeviceToHost, cs));
            total += fo.total_fin;
        }
        const auto t1 = std::chrono::steady_clock::now();
        CDB_CUDA(cudaStreamSynchronize(st));
        if (nparts > 1) CDB_CUDA(cudaStreamSynchronize(tc.copy_stream));
        ro[nreq] = (i64)total;
        if (dbg)
            fprintf(stderr, "[cdb_filter] %d part(s): device %.3f ms, rest of the copy to the host %.3f ms (%lld pairs)\n", nparts,
                    std::chrono::duration<double, std::milli>(t1 - t0).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count(), (long long)total);
        // Requests that did not fit the warp path: their id-ascending survivors come back whole and get the reference's own
        // final step here — the std::sort of src/interface.cpp:143-146 and the span of :196-209.
        std::vector<std::pair<int64_t, int64_t>> v;
        for (int k = 0; k < nparts; ++k) {
            FilterOut& fo = *parts[k];
            for (i64 rl : fo.pending) {
                const i64 r = part_begin[k] + rl;
                u64 off = 0, len = 0;
                CDB_CUDA(cudaMemcpyAsync(&off, fo.raw_off.p + rl, 8, cudaMemcpyDeviceToHost, st));
                CDB_CUDA(cudaMemcpyAsync(&len, fo.raw_len.p + rl, 8, cudaMemcpyDeviceToHost, st));
                CDB_CUDA(cudaStreamSynchronize(st));
                v.resize((size_t)len);
                static_assert(sizeof(std::pair<int64_t, int64_t>) == 16, "pair layout");
                if (len) CDB_CUDA(cudaMemcpyAsync((void*)v.data(), fo.raw.p + 2 * off, (size_t)len * 16, cudaMemcpyDeviceToHost, st));
                CDB_CUDA(cudaStreamSynchronize(st));
                std::sort(v.begin(), v.end(), [](auto x, auto y) { return x.second > y.second; });
                const i64 take = ro[r + 1] - ro[r];
                i64 first = 0;
                if (b->span) first = b->span[2 * r] < 0 ? 0 : b->span[2 * r];
                if (take > 0) std::memcpy(pr + 2 * ro[r], (const void*)(v.data() + first), (size_t)take * 16);
            }
        }
        out->nreq = nreq;
        out->total_pairs = (i64)total;
        out->row_off = ro;
        out->pairs = pr;
        out->matched = (const i64*)own->p[2];
        out->_owner = own;
    } catch (...) {
        cudaStreamSynchronize(st);
        if (tc.copy_stream) cudaStreamSynchronize(tc.copy_stream);
        own->release();
        delete own;
        throw;
    }
    return CDB_OK;
    CDB_CATCH
}

void cdb_filter_result_free(cdb_filter_result* r) {
    if (!r) return;
    if (FilterOwner* own = reinterpret_cast<FilterOwner*>(r->_owner)) {
        own->release();
        delete own;
    }
    std::memset(r, 0, sizeof(*r));
}

cdb_status cdb_locate_batch_device(const cdb_index* h, const void* d_pat, const int64_t* d_pat_off, int64_t npat,
                                   void* stream, cdb_device_result* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || npat < 0) throw Error(CDB_ERR_ARG, "cdb_locate_batch_device: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    std::memset(out, 0, sizeof(*out));
    DeviceSetter ds(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (npat == 0) {
        DevBuf<i64> ro(1, st);
        CDB_CUDA(cudaMemsetAsync(ro.p, 0, 8, st));
        CDB_CUDA(cudaStreamSynchronize(st));
        out->row_off = ro.detach();
        out->_owner = (void*)st;
        return CDB_OK;
    }
    locate_device(*ix, (const u8*)d_pat, d_pat_off, npat, st, out);
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_locate_batch_device_ex(const cdb_index* h, const void* d_pat, const int64_t* d_pat_off, int64_t npat,
                                      void* stream, cdb_rows_ready_fn rows_ready, void* user, cdb_device_result* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || npat < 1) throw Error(CDB_ERR_ARG, "cdb_locate_batch_device_ex: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    std::memset(out, 0, sizeof(*out));
    DeviceSetter ds(ix->device);
    locate_device(*ix, (const u8*)d_pat, d_pat_off, npat, (cudaStream_t)stream, out, false, rows_ready, user);
    return CDB_OK;
    CDB_CATCH
}

void cdb_device_result_free(cdb_device_result* r) {
    if (!r) return;
    cudaStream_t st = (cudaStream_t)r->_owner;
    if (r->row_off) cudaFreeAsync(r->row_off, st);
    if (r->pairs) cudaFreeAsync(r->pairs, st);
    if (r->left) cudaFreeAsync(r->left, st);
    if (r->right) cudaFreeAsync(r->right, st);
    if (r->stats32) cudaFreeAsync(r->stats32, st);
    if (r->row_flags) cudaFreeAsync(r->row_flags, st);
    std::memset(r, 0, sizeof(*r));
}

cdb_status cdb_locate_batch(const cdb_index* h, const void* pat, const int64_t* pat_off, int64_t npat, cdb_result* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || npat < 0 || (npat > 0 && !pat_off)) throw Error(CDB_ERR_ARG, "cdb_locate_batch: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    std::memset(out, 0, sizeof(*out));
    DeviceSetter ds(ix->device);
    // the reference rejects an empty keyword before touching the index (src/index.cpp:239-241)
    for (i64 q = 0; q < npat; ++q)
        if (pat_off[q + 1] <= pat_off[q]) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");
    cudaStream_t st = thread_ctx(ix->device).stream;
    HostResultOwner* own = new HostResultOwner{nullptr, 0, nullptr, 0};
    // small batches (CDB_SMALL_BATCH, default 256 keywords): one upload, two launches, one synchronisation; rows arrive in the
    // thread's mapped pinned buffer and are copied into the caller's result
    if (const int small = small_batch_limit(); small > 0 && npat > 0 && npat <= small) {
        try {
            SmallResult sr;
            if (locate_small(*ix, (const u8*)pat, pat_off, npat, st, &sr)) {
                i64 total_pairs = 0;
                for (i64 q = 0; q < npat; ++q) total_pairs += (i64)sr.rowlen[q];
                own->plain = true;
                own->row_off = malloc((size_t)(npat + 1) * 8);
                own->pairs = malloc((size_t)(total_pairs ? total_pairs : 1) * 16);
                if (!own->row_off || !own->pairs) throw Error(CDB_ERR_NOMEM, "cdb_locate_batch: out of host memory");
                i64* ro = (i64*)own->row_off;
                i64* pr = (i64*)own->pairs;
                ro[0] = 0;
                u64 base = 0;  // rows sit at the scan of the occurrence counts
                for (i64 q = 0; q < npat; ++q) {
                    std::memcpy(pr + 2 * ro[q], sr.pairs + 2 * base, (size_t)sr.rowlen[q] * 16);
                    ro[q + 1] = ro[q] + (i64)sr.rowlen[q];
                    base += sr.rowocc[q];
                }
                g_locate_stats = LocateStats{};
                g_locate_stats.npat = npat;
                g_locate_stats.total_pairs = total_pairs;
                g_locate_stats.total_occ = sr.total_occ;
                out->npat = npat;
                out->total_pairs = total_pairs;
                out->total_occurrences = sr.total_occ;
                out->row_off = ro;
                out->pairs = pr;
                out->_owner = own;
                return CDB_OK;
            }
        } catch (...) {
            cudaStreamSynchronize(st);
            free(own->row_off);  // the small path only ever mallocs
            free(own->pairs);
            delete own;
            throw;
        }
    }
    cdb_device_result dr;
    std::memset(&dr, 0, sizeof(dr));
    try {
        const i64 pbytes = npat ? pat_off[npat] - pat_off[0] : 0;
        DevBuf<u8> d_pat((size_t)pbytes + 8, st);
        DevBuf<i64> d_poff((size_t)npat + 1, st);
        std::vector<i64> rel((size_t)npat + 1, 0);
        for (i64 q = 0; q <= npat && npat; ++q) rel[q] = pat_off[q] - pat_off[0];
        if (pbytes) CDB_CUDA(cudaMemcpyAsync(d_pat.p, (const u8*)pat + pat_off[0], (size_t)pbytes, cudaMemcpyHostToDevice, st));
        CDB_CUDA(cudaMemcpyAsync(d_poff.p, rel.data(), (size_t)(npat + 1) * 8, cudaMemcpyHostToDevice, st));
        if (npat == 0) {
            CDB_CUDA(cudaStreamSynchronize(st));
        } else {
            locate_device(*ix, d_pat.p, d_poff.p, npat, st, &dr);
        }
        own->row_off = g_pinned.get((size_t)(npat + 1) * 8, &own->row_cap);
        own->pairs = g_pinned.get((size_t)(dr.total_pairs ? dr.total_pairs : 1) * 16, &own->pairs_cap);
        if (npat) {
            CDB_CUDA(cudaMemcpyAsync(own->row_off, dr.row_off, (size_t)(npat + 1) * 8, cudaMemcpyDeviceToHost, st));
            if (dr.total_pairs)
                CDB_CUDA(cudaMemcpyAsync(own->pairs, dr.pairs, (size_t)dr.total_pairs * 16, cudaMemcpyDeviceToHost, st));
        } else {
            *(i64*)own->row_off = 0;
        }
        CDB_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cdb_device_result_free(&dr);
        cudaStreamSynchronize(st);
        if (own->row_off) g_pinned.put(own->row_off, own->row_cap);
        if (own->pairs) g_pinned.put(own->pairs, own->pairs_cap);
        delete own;
        throw;
    }
    out->npat = npat;
    out->total_pairs = dr.total_pairs;
    out->total_occurrences = dr.total_occurrences;
    out->row_off = (const i64*)own->row_off;
    out->pairs = (const i64*)own->pairs;
    out->_owner = own;
    cdb_device_result_free(&dr);  // stream-ordered frees on the thread's own stream: nothing to wait for
    return CDB_OK;
    CDB_CATCH
}

void cdb_result_free(cdb_result* r) {
    if (!r || !r->_owner) return;
    HostResultOwner* own = static_cast<HostResultOwner*>(r->_owner);
    if (own->shared) delete static_cast<std::shared_ptr<QueryBatchResult>*>(own->shared);
    if (own->plain) {
        free(own->row_off);
        free(own->pairs);
    } else {
        if (own->row_off) g_pinned.put(own->row_off, own->row_cap);
        if (own->pairs) g_pinned.put(own->pairs, own->pairs_cap);
    }
    delete own;
    std::memset(r, 0, sizeof(*r));
}

cdb_status cdb_query(const cdb_index* h, const void* keyword, int64_t len, cdb_result* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || len < 0 || (len > 0 && !keyword)) throw Error(CDB_ERR_ARG, "cdb_query: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    std::memset(out, 0, sizeof(*out));
    if (len == 0) throw Error(CDB_ERR_EMPTY_KEYWORD, "Empty keywords are not allowed");  // src/index.cpp:239-241
    QueryBatcher::row_view v = query_batcher(ix).query(std::string_view(static_cast<const char*>(keyword), (size_t)len));
    HostResultOwner* own = new HostResultOwner{nullptr, 0, nullptr, 0};
    own->shared = new std::shared_ptr<QueryBatchResult>(std::move(v.owner));
    own->view_off[1] = v.count;
    int64_t occ = 0;
    for (int64_t i = 0; i < v.count; ++i) occ += v.pairs[2 * i + 1];
    out->npat = 1;
    out->total_pairs = v.count;
    out->total_occurrences = occ;
    out->row_off = own->view_off;
    out->pairs = v.pairs;
    out->_owner = own;
    return CDB_OK;
    CDB_CATCH
}

cdb_status cdb_query_stats(const cdb_index* h, uint64_t* queries, uint64_t* batches, uint64_t* largest) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix) throw Error(CDB_ERR_ARG, "cdb_query_stats: bad argument");
    coffeedb_b200::micro_batcher_stats st;
    {
        std::lock_guard<std::mutex> lk(ix->batcher_mu);
        if (ix->batcher) st = static_cast<QueryBatcher*>(ix->batcher.get())->statistics();
    }
    if (queries) *queries = st.queries;
    if (batches) *batches = st.batches;
    if (largest) *largest = st.largest;
    return CDB_OK;
    CDB_CATCH
}

static cdb_status spans_host(const Index* ix, const void* kw, const int64_t* kw_off, int64_t nkw, const int64_t* req_kw_off,
                             int64_t nreq, const int64_t* text_req, const int64_t* text_doc, int64_t ntext, cdb_spans* out) {
    std::memset(out, 0, sizeof(*out));
    DeviceSetter ds(ix->device);
    cudaStream_t st = thread_ctx(ix->device).stream;
    std::vector<i64>* so = new std::vector<i64>();
    std::vector<i64>* sp = new std::vector<i64>();
    try {
        locate_spans_batch(*ix, (const u8*)kw, kw_off, nkw, req_kw_off, nreq, text_req, text_doc, ntext, st, *so, *sp);
    } catch (...) {
        cudaStreamSynchronize(st);
        delete so;
        delete sp;
        throw;
    }
    auto* pair = new std::pair<std::vector<i64>*, std::vector<i64>*>(so, sp);
    out->ntext = ntext;
    out->total_spans = (i64)sp->size() / 2;
    out->span_off = so->data();
    out->spans = sp->data();
    out->_owner = pair;
    return CDB_OK;
}

cdb_status cdb_locate_spans(const cdb_index* h, const void* kw, const int64_t* kw_off, int64_t nkw, const int64_t* docs,
                            int64_t ndocs, cdb_spans* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || nkw < 0 || ndocs < 0 || (nkw > 0 && !kw_off) || (ndocs > 0 && !docs))
        throw Error(CDB_ERR_ARG, "cdb_locate_spans: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    // one request that owns every keyword; every document is one of its texts
    const int64_t req_kw_off[2] = {0, nkw};
    std::vector<i64> text_req((size_t)ndocs, 0);
    return spans_host(ix, kw, kw_off, nkw, req_kw_off, 1, text_req.data(), docs, ndocs, out);
    CDB_CATCH
}

cdb_status cdb_locate_spans_batch(const cdb_index* h, const void* kw, const int64_t* kw_off, int64_t nkw,
                                  const int64_t* req_kw_off, int64_t nreq, const int64_t* text_req, const int64_t* text_doc,
                                  int64_t ntext, cdb_spans* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || nkw < 0 || nreq < 0 || ntext < 0 || (nkw > 0 && !kw_off) || !req_kw_off ||
        (ntext > 0 && (!text_req || !text_doc)))
        throw Error(CDB_ERR_ARG, "cdb_locate_spans_batch: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    if (req_kw_off[0] != 0 || req_kw_off[nreq] != nkw) throw Error(CDB_ERR_ARG, "cdb_locate_spans_batch: req_kw_off must span [0, nkw]");
    for (i64 r = 0; r < nreq; ++r)
        if (req_kw_off[r + 1] < req_kw_off[r]) throw Error(CDB_ERR_ARG, "cdb_locate_spans_batch: req_kw_off must be non-decreasing");
    for (i64 t = 0; t < ntext; ++t)
        if (text_req[t] < 0 || text_req[t] >= nreq) throw Error(CDB_ERR_ARG, "cdb_locate_spans_batch: request index out of range");
    return spans_host(ix, kw, kw_off, nkw, req_kw_off, nreq, text_req, text_doc, ntext, out);
    CDB_CATCH
}

cdb_status cdb_locate_spans_batch_device(const cdb_index* h, const void* d_kw, const int64_t* d_kw_off, int64_t nkw,
                                         const int64_t* d_req_kw_off, int64_t nreq, const int64_t* d_text_req,
                                         const int64_t* d_text_doc, int64_t ntext, void* stream, cdb_device_spans* out) {
    CDB_TRY
    const Index* ix = reinterpret_cast<const Index*>(h);
    if (!ix || !out || nkw < 0 || nreq < 0 || ntext < 0) throw Error(CDB_ERR_ARG, "cdb_locate_spans_batch_device: bad argument");
    if (!ix->built) throw Error(CDB_ERR_STATE, "index has not been built");
    std::memset(out, 0, sizeof(*out));
    DeviceSetter ds(ix->device);
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf<u64> soff;
    DevBuf<i64> sp;
    i64 total = 0;
    locate_spans_batch_device(*ix, (const u8*)d_kw, d_kw_off, nkw, d_req_kw_off, nreq, d_text_req, d_text_doc, ntext, st, soff, sp,
                              &total);
    out->ntext = ntext;
    out->total_spans = total;
    out->span_off = reinterpret_cast<int64_t*>(soff.detach());
    out->spans = sp.detach();
    out->_owner = (void*)st;
    return CDB_OK;
    CDB_CATCH
}

void cdb_device_spans_free(cdb_device_spans* s) {
    if (!s) return;
    cudaStream_t st = (cudaStream_t)s->_owner;
    if (s->span_off) cudaFreeAsync(s->span_off, st);
    if (s->spans) cudaFreeAsync(s->spans, st);
    std::memset(s, 0, sizeof(*s));
}

void cdb_spans_free(cdb_spans* s) {
    if (!s || !s->_owner) return;
    auto* pair = static_cast<std::pair<std::vector<i64>*, std::vector<i64>*>*>(s->_owner);
    delete pair->first;
    delete pair->second;
    delete pair;
    std::memset(s, 0, sizeof(*s));
}

// src/database.cpp:78-90
int64_t cdb_splice(const void* text, int64_t tlen, const int64_t* spans, int64_t nspans, const void* left, int64_t llen,
                   const void* right, int64_t rlen, void* out, int64_t out_cap) {
    const int64_t need = tlen + (llen + rlen) * nspans;
    if (!out || out_cap < need) return need;
    const char* t = static_cast<const char*>(text);
    char* o = static_cast<char*>(out);
    int64_t w = 0, prev = 0;
    for (int64_t s = 0; s < nspans; ++s) {
        const int64_t b = spans[2 * s], e = spans[2 * s + 1];
        std::memcpy(o + w, t + prev, (size_t)(b - prev));
        w += b - prev;
        std::memcpy(o + w, left, (size_t)llen);
        w += llen;
        std::memcpy(o + w, t + b, (size_t)(e - b + 1));
        w += e - b + 1;
        std::memcpy(o + w, right, (size_t)rlen);
        w += rlen;
        prev = e + 1;
    }
    std::memcpy(o + w, t + prev, (size_t)(tlen - prev));
    w += tlen - prev;
    return w;
}

}  // extern "C"
