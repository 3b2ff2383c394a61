// The device-resident string index: the state string_index keeps in src/index.h:56-60, laid out for HBM.
//
//   text     u8  [n + pad]   all documents back to back (doc d = text[doc_off[d], doc_off[d+1]))
//   doc_off  i64 [nd + 1]    document-boundary array (replaces std::vector<std::string_view> data)
//   ids      i64 [nd]        external object id per doc index (src/index.h:59)
//   sa       u32|u64 [n]     packed suffix array, element = (offset_in_doc << bits1) | doc_index — the
//                            reference's own element format (src/index.cpp:209-215), so export is a memcpy
#pragma once
#include <memory>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "staging.cuh"

namespace cdb {

constexpr i64 kTextPad = 64;  // bytes of zero padding after the text (16-byte window loads never fault)

// Byte -> symbol re-coding of one corpus: 0 is reserved for end-of-document, the bytes that occur map to
// 1..sigma in unsigned byte order (so symbol order == memcmp order, with "end" first as in src/index.h:66-73).
struct SymTab {
    u16 sym[256];
};

// Document listing of the prefix directory's buckets (locate.cu: build_listing).  The suffixes of bucket c — those whose
// first pt_k symbols code to c, SA ranks [ptab[c], ptab[c+1]) — have their documents' ids stored at the same ranks,
// sorted by doc index (order 0, what string_index::query reports) or by id (order 1, what filter() merges), as
// (id - base) split into a 32-bit plane and an hw-byte plane.  A keyword of exactly pt_k symbols is then answered by
// streaming its bucket: no sort, no ids[] lookup.  Which buckets are listed (1 .. 1024 suffixes; longer ones take the
// general path), their number of distinct documents and whether one repeats sit in the high bits of ptab[c] (locate.cu).
struct Listing {
    u32* lo = nullptr;      // [n]
    void* hi = nullptr;     // [n] u8 / u16 / u32 (hw = 1 / 2 / 4), nullptr when hw == 0
    int* d_flag = nullptr;  // build-time: set when two documents share an id (the listing is then dropped)
    int hw = 0;
    i64 base = 0;
    size_t bytes = 0;
    double build_ms = 0;
    ~Listing() {
        if (lo) cudaFree(lo);
        if (hi) cudaFree(hi);
        if (d_flag) cudaFree(d_flag);
    }
};

struct Index {
    cdb_options opt{};
    int device = 0;
    // host staging filled by cdb_add (src/index.cpp:174-177): text in page-locked chunks that leave for the device as they
    // fill up (staging.cuh), offsets and ids in plain vectors (16 B per document)
    TextStaging h_text;
    std::vector<i64> h_off{0};
    std::vector<i64> h_ids;
    bool host_dropped = false;  // the staging copy was released after build (keep_host_copy == 0)
    // device corpus
    const u8* d_text = nullptr;
    const i64* d_off = nullptr;
    const i64* d_ids = nullptr;
    void* own_text = nullptr;
    void* own_off = nullptr;
    void* own_ids = nullptr;
    i64 n = 0, nd = 0;
    int width = 0, bits1 = 0, bits2 = 0;
    u64 mask = 0;
    void* d_sa = nullptr;
    bool built = false;
    bool mixed = false;  // text holds bytes on both sides of 0x80 (note N1 applies when n > 4096)
    i64 chuck_size = 0;  // max(4096, n/256), src/index.cpp:218
    // Prefix directory over the first pt_k symbols (pt_b bits each): ptab[c] = number of suffixes whose pt_k-symbol
    // code is < c, for c = 0 .. 2^(pt_b*pt_k).  Resolves the SA interval of short patterns with two lookups and
    // narrows the binary search of longer ones.  Only built when the array is sorted under one comparator
    // (i.e. not in the note-N1 layout), see build_prefix_table().
    SymTab symtab{};
    int sigma = 0, pt_b = 0, pt_k = 0;
    u64* d_ptab = nullptr;
    // build statistics
    double build_ms = 0, sort_ms = 0;
    i64 rounds = 0, chunks = 0;
    i64 staged_on_device = 0;       // bytes of text that were already in HBM when the last cdb_build started (staging.cuh)
    bool loaded_from_file = false;  // the last build read a saved array instead of sorting (cdb_build_or_load)
    // Id order of the documents, for filter() (src/interface.cpp:79-135 merges rows sorted by id; string_index::query
    // reports them in doc order).  ids_order: -1 not examined yet, 1 = ids ascend with the doc index (doc order IS id order),
    // 0 = they do not: rank_tab[doc] = rank of the document's id among all ids, ids_by_rank[rank] = that id.  Filled on
    // the first id-ordered locate (locate.cu: id_order_tables), under order_mu.
    mutable std::mutex order_mu;
    mutable int ids_order = -1;
    mutable u32* d_rank_tab = nullptr;
    mutable i64* d_ids_by_rank = nullptr;
    mutable u32* d_sa_rank = nullptr;  // rank companion of the suffix array (locate.cu: id_order_tables), when memory allows
    // document listings, [0] doc order, [1] id order (the same object when the ids ascend with the doc index); a locate
    // holds a reference for its duration, so one can be dropped (to make room for the other order) while calls are in flight.
    // listing_state: 0 not tried yet, 1 present, -1 not available (no directory, ids too wide, no memory, CDB_LISTING=0)
    mutable std::mutex listing_mu;
    mutable std::shared_ptr<Listing> listing[2];
    mutable int listing_state[2] = {0, 0};
    mutable int listing_miss[2] = {0, 0};  // calls of an order that found the other order's listing in its way (see get_listing)
    mutable int listing_skip[2] = {0, 0};  // calls of an order to let pass before its build is tried again (memory was short)
    // cdb_query's coalescing queue (capi.cu), created on first use
    mutable std::mutex batcher_mu;
    mutable std::shared_ptr<void> batcher;

    ~Index();
    void free_device();
};

// A saved suffix array that may stand in for the sort (persist.cu): try_load is called once the geometry of the staged
// corpus is known (n, widths, flags) and returns true after it has filled ix.d_sa from the file.
struct SavedArraySource {
    virtual bool try_load(Index& ix, cudaStream_t st) const = 0;
    virtual ~SavedArraySource() {}
};
void build_index(Index& ix, cudaStream_t st, const SavedArraySource* saved = nullptr);
// sa_build.cu: 64-bit hash of the corpus the index was built from (text, doc_off, ids)
u64 corpus_hash(const Index& ix, i64 n, cudaStream_t st);
// locate.cu: fills ix.d_ptab (after the suffix array is complete)
void build_prefix_table(Index& ix, cudaStream_t st);
// locate.cu: the listing of the given order (built on first use when memory allows), or an empty pointer
std::shared_ptr<Listing> get_listing(const Index& ix, int order, cudaStream_t st);

}  // namespace cdb
