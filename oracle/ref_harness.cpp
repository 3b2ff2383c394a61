// TEST INFRASTRUCTURE ONLY.  C-ABI harness around the UNMODIFIED reference string index.
//
// Linked against an object compiled straight from /root/reference/src/index.cpp (see Makefile);
// nothing from the reference is copied into this repository.  Only tests/, bench.py's CPU baseline
// legs and __graft_entry__.smoke() may load the resulting oracle/_ref/libcoffeeref.so.
//
// What it exposes:
//   ref_create/add/build/query/destroy   -> string_index::{add,build,query}   (src/index.cpp:174-326)
//   ref_export_sa                        -> the private packed suffix array   (src/index.h:56-60)
//   ref_query_batch / ref_query_batch_csr-> query() from T threads pulling 64-pattern blocks, the
//                                           concurrency model of the reference's HTTP pool
//                                           (package/httplib.h:97-101); used for CPU baseline timing
//
// Private state is read without editing the reference: string_index declares the public member
// template `parallel_sort<T>()` (src/index.h:85); an explicit specialisation for a harness-only
// tag type is a member function and may therefore read `sa`, `bits`, `mask`, `size`.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <stdexcept>
#include <string>
#include <string_view>
#include <thread>
#include <vector>
#include "index.h"

namespace {
struct peek_tag {};
struct adopt_tag {};
struct adopt_in {
    void* sa = nullptr;
    int width = 4;
    uint64_t bits = 1, size = 0;
};
thread_local adopt_in g_adopt;
struct peek_out {
    int width = 0;
    uint64_t bits = 0, mask = 0, size = 0;
    const void* sa = nullptr;
};
thread_local peek_out g_peek;

struct ref_handle {
    string_index idx;
    std::deque<std::string> owned;  // keeps the borrowed string_views alive (database.cpp:263-264)
    bool built = false;
    bool adopted = false;  // sa points into caller memory: must be detached before ~string_index deletes it
};

void copy_err(const char* what, char* err, int errlen) {
    if (err && errlen > 0) {
        std::strncpy(err, what, errlen - 1);
        err[errlen - 1] = 0;
    }
}
}  // namespace

template <>
void string_index::parallel_sort<peek_tag>() const {
    g_peek.bits = bits;
    g_peek.mask = mask;
    g_peek.size = size;
    std::visit(
        [](auto* p) {
            g_peek.width = (int)sizeof(*p);
            g_peek.sa = p;
        },
        sa);
}

// Test-only: makes the reference's query() run on a suffix array supplied by the caller (bench.py injects the
// array built on the GPU so that the CPU baseline can be timed on the full 10 GB configuration, whose
// reference build would take tens of minutes).  Sets exactly the members build() sets (src/index.cpp:201-208).
template <>
void string_index::parallel_sort<adopt_tag>() const {
    auto* self = const_cast<string_index*>(this);
    self->bits = g_adopt.bits;
    self->mask = (g_adopt.bits >= 64) ? ~0ull : ((1ull << g_adopt.bits) - 1);
    self->size = g_adopt.size;
    if (g_adopt.width == 4)
        self->sa = static_cast<uint32_t*>(g_adopt.sa);
    else
        self->sa = static_cast<uint64_t*>(g_adopt.sa);
}

extern "C" {

void* ref_create() { return new ref_handle(); }

void ref_destroy(void* h) {
    auto* r = static_cast<ref_handle*>(h);
    if (r && r->adopted) {
        g_adopt = adopt_in{};
        r->idx.parallel_sort<adopt_tag>();  // sa = nullptr: the destructor's delete[] becomes a no-op
    }
    delete r;
}

// add() without the owning copy: the views point into caller memory, exactly as the reference's own loader
// hands views into ::data (database.cpp:263-264).  The caller keeps `text` alive.
void ref_add_many_borrowed(void* h, const int64_t* ids, const char* text, const int64_t* off, int64_t nd) {
    auto* r = static_cast<ref_handle*>(h);
    for (int64_t d = 0; d < nd; ++d) r->idx.add(ids[d], std::string_view(text + off[d], (size_t)(off[d + 1] - off[d])));
}

void ref_adopt_sa(void* h, void* sa, int width, uint64_t bits, uint64_t size) {
    auto* r = static_cast<ref_handle*>(h);
    g_adopt.sa = sa;
    g_adopt.width = width;
    g_adopt.bits = bits;
    g_adopt.size = size;
    r->idx.parallel_sort<adopt_tag>();
    r->adopted = true;
    r->built = true;
}

void ref_add(void* h, int64_t id, const char* ptr, int64_t len) {
    auto* r = static_cast<ref_handle*>(h);
    r->owned.emplace_back(ptr, (size_t)len);
    r->idx.add(id, std::string_view(r->owned.back()));
}

// Adds `nd` documents laid out back to back in `text` (doc d = text[off[d], off[d+1])).
void ref_add_many(void* h, const int64_t* ids, const char* text, const int64_t* off, int64_t nd) {
    for (int64_t d = 0; d < nd; ++d) {
        ref_add(h, ids[d], text + off[d], off[d + 1] - off[d]);
    }
}

// 0 = ok, 1 = reference threw (message copied to err).
int ref_build(void* h, char* err, int errlen) {
    auto* r = static_cast<ref_handle*>(h);
    try {
        r->idx.build();
        r->built = true;
        return 0;
    } catch (const std::exception& e) {
        copy_err(e.what(), err, errlen);
        return 1;
    }
}

// Fills width (4|8), bits, mask, size.  If buf != NULL copies size*width bytes of raw SA into it.
void ref_export_sa(void* h, int* width, uint64_t* bits, uint64_t* mask, uint64_t* size, void* buf) {
    auto* r = static_cast<ref_handle*>(h);
    r->idx.parallel_sort<peek_tag>();
    if (width) *width = g_peek.width;
    if (bits) *bits = g_peek.bits;
    if (mask) *mask = g_peek.mask;
    if (size) *size = g_peek.size;
    if (buf && g_peek.sa) std::memcpy(buf, g_peek.sa, g_peek.size * g_peek.width);
}

// Returns number of (id,count) pairs, -1 if the reference threw.  *out is malloc'd (2 int64 per pair).
int64_t ref_query(void* h, const char* kw, int64_t len, int64_t** out, char* err, int errlen) {
    auto* r = static_cast<ref_handle*>(h);
    try {
        auto res = r->idx.query(std::string(kw, (size_t)len));
        int64_t* buf = (int64_t*)std::malloc(sizeof(int64_t) * 2 * (res.size() + 1));
        for (size_t i = 0; i < res.size(); ++i) {
            buf[2 * i] = res[i].first;
            buf[2 * i + 1] = res[i].second;
        }
        *out = buf;
        return (int64_t)res.size();
    } catch (const std::exception& e) {
        copy_err(e.what(), err, errlen);
        *out = nullptr;
        return -1;
    }
}

void ref_free(void* p) { std::free(p); }

// Runs all patterns through string_index::query on `nthreads` threads (64-pattern blocks from an
// atomic counter).  Returns wall seconds; totals let both sides be seen to do identical work.
double ref_query_batch(void* h, const char* pat, const int64_t* pat_off, int64_t npat, int nthreads,
                       int64_t* total_pairs, int64_t* total_occ) {
    auto* r = static_cast<ref_handle*>(h);
    std::atomic<int64_t> next{0}, pairs{0}, occ{0};
    auto worker = [&]() {
        int64_t lp = 0, lo = 0;
        for (;;) {
            int64_t b = next.fetch_add(64);
            if (b >= npat) break;
            int64_t e = std::min<int64_t>(npat, b + 64);
            for (int64_t q = b; q < e; ++q) {
                auto res = r->idx.query(std::string(pat + pat_off[q], (size_t)(pat_off[q + 1] - pat_off[q])));
                lp += (int64_t)res.size();
                for (auto& pr : res) lo += pr.second;
            }
        }
        pairs += lp;
        occ += lo;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (total_pairs) *total_pairs = pairs.load();
    if (total_occ) *total_occ = occ.load();
    return std::chrono::duration<double>(t1 - t0).count();
}

// Same, but keeps the answers: row_off[npat+1] (caller-allocated) and *pairs (malloc'd, 2 int64 per
// pair, rows in pattern order, each row in the reference's own order = ascending doc index).
int64_t ref_query_batch_csr(void* h, const char* pat, const int64_t* pat_off, int64_t npat, int nthreads,
                            int64_t* row_off, int64_t** pairs_out) {
    auto* r = static_cast<ref_handle*>(h);
    std::vector<std::vector<std::pair<int64_t, int64_t>>> rows((size_t)npat);
    std::atomic<int64_t> next{0};
    auto worker = [&]() {
        for (;;) {
            int64_t b = next.fetch_add(64);
            if (b >= npat) break;
            int64_t e = std::min<int64_t>(npat, b + 64);
            for (int64_t q = b; q < e; ++q) {
                rows[q] = r->idx.query(std::string(pat + pat_off[q], (size_t)(pat_off[q + 1] - pat_off[q])));
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    int64_t total = 0;
    for (int64_t q = 0; q < npat; ++q) {
        row_off[q] = total;
        total += (int64_t)rows[q].size();
    }
    row_off[npat] = total;
    int64_t* buf = (int64_t*)std::malloc(sizeof(int64_t) * 2 * (total + 1));
    int64_t k = 0;
    for (auto& row : rows)
        for (auto& pr : row) {
            buf[2 * k] = pr.first;
            buf[2 * k + 1] = pr.second;
            ++k;
        }
    *pairs_out = buf;
    return total;
}

// One-keyword requests through the rest of the reference's request path: what filter() and the `span` handling do with
// the result of query(key, keyword) when the constraint object holds that one key (src/interface.cpp:79-83: sort by
// (id, count); :143-146: std::sort by descending $correlation; :196-209: result[span0, span1)).  Those lines are
// restated here verbatim around the UNMODIFIED string_index::query, because filter() itself is reachable only through
// response() and the global object table (src/interface.cpp:149, src/database.cpp:26-33), which cannot hold 10^8
// objects.  Returns wall seconds; with row_off != NULL the sliced answers are kept (malloc'd *pairs_out).
double ref_filter_span_batch(void* h, const char* pat, const int64_t* pat_off, int64_t npat, int nthreads, int64_t span0,
                             int64_t span1, int64_t* row_off, int64_t** pairs_out, int64_t* matched) {
    auto* r = static_cast<ref_handle*>(h);
    std::vector<std::vector<std::pair<int64_t, int64_t>>> rows(row_off ? (size_t)npat : 0);
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> sink{0};
    auto worker = [&]() {
        int64_t acc = 0;
        for (;;) {
            int64_t b = next.fetch_add(64);
            if (b >= npat) break;
            int64_t e = std::min<int64_t>(npat, b + 64);
            for (int64_t q = b; q < e; ++q) {
                auto result = r->idx.query(std::string(pat + pat_off[q], (size_t)(pat_off[q + 1] - pat_off[q])));
                std::ranges::sort(result);                                          // src/interface.cpp:82
                std::sort(result.begin(), result.end(), [](auto x, auto y) {        // src/interface.cpp:143-146
                    return x.second > y.second;
                });
                if (matched) matched[q] = (int64_t)result.size();
                if (span0 >= (int64_t)std::ssize(result)) {                         // src/interface.cpp:198-207
                    result.clear();
                } else {
                    auto end = result.end();
                    if (span1 < (int64_t)std::ssize(result)) end = result.begin() + span1;
                    result = {result.begin() + span0, end};
                }
                acc += (int64_t)result.size();
                if (row_off) rows[q] = std::move(result);
            }
        }
        sink += acc;
    };
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int i = 1; i < nthreads; ++i) th.emplace_back(worker);
    worker();
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (row_off) {
        int64_t total = 0;
        for (int64_t q = 0; q < npat; ++q) {
            row_off[q] = total;
            total += (int64_t)rows[q].size();
        }
        row_off[npat] = total;
        int64_t* buf = (int64_t*)std::malloc(sizeof(int64_t) * 2 * (total + 1));
        int64_t k = 0;
        for (auto& row : rows)
            for (auto& pr : row) {
                buf[2 * k] = pr.first;
                buf[2 * k + 1] = pr.second;
                ++k;
            }
        *pairs_out = buf;
    }
    return std::chrono::duration<double>(t1 - t0).count();
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
