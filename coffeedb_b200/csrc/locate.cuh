#pragma once
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>

#include "index.cuh"

namespace cdb {
// Batched locate with patterns and results in device memory (see locate.cu).
// id_order: rows in ascending id instead of ascending doc index (what filter() merges, src/interface.cpp:79-135): the
// interval's doc indices are mapped to id ranks before they are sorted, and translated through ids_by_rank.
// rows_ready: called on the host once the kernels that produce stats32 (per-pattern row length and occurrences) have been
// enqueued, before translate is — a sharded caller launches its exchange on another stream from there (cdb_locate_batch_device_ex).
// lazy: a caller that reads rows selectively (cdb_filter) passes a LazyListed — the rows answered from the document listing
// in which no document repeats are then NOT written into out->pairs (their CSR slots stay unwritten); the caller reads
// them from the listing (listed_row_value) and asks for the ones it needs in full with emit_listed_rows.
struct LazyListed {
    std::shared_ptr<Listing> lst;  // held while the rows are in use
    u64* pre = nullptr;            // device [npat]: kPreListed | kPreRepeat | row length per row (0: an ordinary row); owned
    cudaStream_t stream = nullptr;
    bool active = false;           // some rows were left unwritten
    ~LazyListed() {
        if (pre) cudaFreeAsync(pre, stream);
    }
};
void locate_device(const Index& ix, const u8* d_pat, const i64* d_pat_off, i64 npat, cudaStream_t st,
                   cdb_device_result* out, bool id_order = false, cdb_rows_ready_fn rows_ready = nullptr,
                   void* rows_ready_user = nullptr, LazyListed* lazy = nullptr);
// writes the unwritten listed rows q with need[q] != 0 into res.pairs
void emit_listed_rows(const cdb_device_result& res, const LazyListed& lazy, const u8* d_need, cudaStream_t st);

// per-pattern word of a row that is answered from the document listing: bit 63 set, bit 62 = some document repeats,
// low bits = distinct documents (the exact row length); 0 = the row takes the general path
constexpr u64 kPreListed = 1ull << 63;
constexpr u64 kPreRepeat = 1ull << 62;
constexpr u64 kPreCount = kPreRepeat - 1;
// A directory entry holds an SA rank in its low 48 bits.  Once a listing has been built, the high bits of ptab[c] describe
// bucket c (ranks [ptab[c], ptab[c+1])): bit 63 = listed (1 .. kWarpCap suffixes), bit 62 = some document occurs more than
// once, bits 48..58 = distinct documents.  The search finds them in the entry it reads anyway.
constexpr u64 kPtRank = (1ull << 48) - 1;
constexpr int kPtCountShift = 48;
constexpr u64 kPtCountMask = 0x7ffull;

constexpr int kMaxRanges = 64;   // doc-range partitions of ids[] used by translate_kernel (each <= ~32 MB of ids)
constexpr int kTrWarps = 8;      // warps per CTA of translate_kernel / listing_translate_kernel

// resident CTAs per SM of a kernel (its persistent grids are exactly one wave), cached per kernel and device
inline int resident_ctas(const void* kernel, int threads, size_t smem = 0) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, int> cache;
    int dev = 0;
    CDB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({kernel, dev});
    if (it != cache.end()) return it->second;
    int v = 0;
    CDB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, threads, smem));
    v = v > 0 ? v : 1;
    cache[{kernel, dev}] = v;
    return v;
}

// Doc-range partition of ids[] for translate_kernel: ranges of 2^rshift documents (default 2^22 = 32 MB of ids,
// CDB_RANGE_BITS overrides), at most kMaxRanges of them.
inline void ids_ranges(i64 nd, int* nranges, int* rshift) {
    const char* e = getenv("CDB_RANGE_BITS");  // read per call: the tests switch it to exercise many ranges
    const int env_bits = e ? atoi(e) : 22;
    int sh = env_bits < 8 ? 8 : (env_bits > 40 ? 40 : env_bits);
    while (ceil_div(nd > 0 ? nd : 1, (i64)1 << sh) > kMaxRanges) ++sh;
    *rshift = sh;
    *nranges = (int)ceil_div(nd > 0 ? nd : 1, (i64)1 << sh);
}

// listing.cu: the rows of a batch that are answered from the document listing.  mode 0: every listed row; 1: only the rows in
// which a document repeats (the others are read lazily by the caller); 2: the rows without repeats with need[q] != 0
// spare_ctas: CTAs the persistent grid leaves free (room for a collective that runs beside it)
void launch_listing_emit(const Listing& L, const u64* pre, const i64* left, const i64* right, const u64* row_off, i64 npat, i64* pairs,
                         int mode, const u8* need, cudaStream_t st, int spare_ctas);
// listing.cu: rowlen[q] = the listed rows' lengths (from pre[]); the other rows are zeroed when write_zero
void launch_listing_rowlen(const u64* pre, i64 npat, u64* rowlen, int write_zero, cudaStream_t st);

#ifdef __CUDACC__
// (id - base) of listing entry i; hw = bytes of the high plane (0, 1, 2, 4)
__device__ __forceinline__ u64 listed_row_value(const u32* __restrict__ lo, const void* __restrict__ hi, int hw, i64 i) {
    u64 v = (u64)__ldg(lo + i);
    if (hw == 1) v |= (u64)__ldg(reinterpret_cast<const u8*>(hi) + i) << 32;
    else if (hw == 2) v |= (u64)__ldg(reinterpret_cast<const u16*>(hi) + i) << 32;
    else if (hw == 4) v |= (u64)__ldg(reinterpret_cast<const u32*>(hi) + i) << 32;
    return v;
}
#endif
// Small-batch fast path (locate.cu): one upload, two launches, one synchronisation.  On success the rows are in mapped
// pinned memory checked out of a per-device pool — valid until the SmallResult is destroyed: row q has rowlen[q] pairs at
// pairs + 2 * (rowocc[0] + ... + rowocc[q-1]).  Returns false when the batch has to take the general path.
struct SmallResult {
    i64 total_occ = 0;
    const u64* rowlen = nullptr;
    const u64* rowocc = nullptr;
    const i64* pairs = nullptr;
    void* _ctx = nullptr;  // the buffer set the rows live in (returned to the pool by the destructor)
    int _device = 0;
    SmallResult() {}
    SmallResult(const SmallResult&) = delete;
    SmallResult& operator=(const SmallResult&) = delete;
    ~SmallResult();
};
int small_batch_limit();  // CDB_SMALL_BATCH (0 = path disabled)
bool locate_small(const Index& ix, const u8* pat, const i64* pat_off, i64 npat, cudaStream_t st, SmallResult* res);
// Highlight spans (spans.cu): merged occurrence spans of each request's keywords inside each of its texts.
// text t = document text_doc[t] highlighted for request text_req[t]; request r owns keywords [req_kw_off[r], req_kw_off[r+1]).
void locate_spans_batch(const Index& ix, const u8* kw, const i64* kw_off, i64 nkw, const i64* req_kw_off, i64 nreq,
                        const i64* text_req, const i64* text_doc, i64 ntext, cudaStream_t st, std::vector<i64>& span_off,
                        std::vector<i64>& spans);
void locate_spans_batch_device(const Index& ix, const u8* d_kw, const i64* d_kw_off, i64 nkw, const i64* d_req_kw_off, i64 nreq,
                               const i64* d_text_req, const i64* d_text_doc, i64 ntext, cudaStream_t st, DevBuf<u64>& span_off,
                               DevBuf<i64>& spans, i64* total_out);
}  // namespace cdb
