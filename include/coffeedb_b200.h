/* coffeedb_b200 — C ABI of the B200-native string index (suffix-array build + batched substring locate).
 *
 * This is the drop-in boundary for CoffeeDB's string-index hot path.  The reference has no
 * FFI layer; its seam is the abstract C++ class `index` (src/index.h:9-23) and the three call sites in
 * src/database.cpp (add :255-264, build :276-278, query :387-393).  Every entry point below names the
 * reference interface it replaces.  A `class string_index : public index` adaptor over this ABI is in
 * coffeedb_b200/host/string_index.hpp; INTEGRATION.md shows how a maintainer links it under database.cpp.
 *
 * Conventions
 *   - plain pointers and sizes only; all integers little-endian host types
 *   - every function returns a cdb_status; on failure cdb_last_error() (thread-local) holds the message.
 *     For the three conditions the reference throws on, the message is the reference's exact text, so the
 *     adaptor can rethrow std::runtime_error(cdb_last_error()) and the server's catch (src/server.cpp:58-62)
 *     produces the same HTTP 500 body.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with CDB_ERR_CUDA.
 *   - documents are identified by their "doc index" = order of cdb_add calls (src/index.cpp:174-177);
 *     results are reported as (ids[doc index], occurrences), rows in ascending doc index, exactly as
 *     string_index::query does (src/index.cpp:316-322).
 */
#ifndef COFFEEDB_B200_H
#define COFFEEDB_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum cdb_status {
    CDB_OK = 0,
    CDB_ERR_TOO_MUCH_DATA = 1,    /* src/index.cpp:195-197 "The amount of data exceeds the maximum range that CoffeeDB can handle" */
    CDB_ERR_TOO_MANY_OBJECTS = 2, /* src/index.cpp:198-200 "The number of objects exceeds the maximum range that CoffeeDB can handle" */
    CDB_ERR_EMPTY_KEYWORD = 3,    /* src/index.cpp:239-241 "Empty keywords are not allowed" */
    CDB_ERR_CUDA = 4,             /* CUDA runtime failure or no device */
    CDB_ERR_NOMEM = 5,            /* host or device allocation failed */
    CDB_ERR_STATE = 6,            /* call order violated (e.g. locate before build, add after build) */
    CDB_ERR_ARG = 7               /* bad argument */
} cdb_status;

typedef struct cdb_index cdb_index; /* replaces one `string_index` object (src/index.h:54-86) */

typedef struct cdb_options {
    int32_t device;            /* CUDA device ordinal; -1 = current device */
    int32_t compat_signed;     /* 1 (default) = reproduce the reference's signed-radix / unsigned-leaf order on
                                  corpora mixing bytes <0x80 and >=0x80 (SURVEY.md §8 note N1); 0 = plain unsigned */
    int64_t workspace_bytes;   /* cap on temporary device memory used by build (0 = 60 % of free memory) */
    int32_t keep_host_copy;    /* 1 = keep the host staging copy of the text after build: cdb_add + cdb_build may then be
                                  repeated on the same handle; 0 (default) = release it, later cdb_add / cdb_build
                                  fail with CDB_ERR_STATE (the reference fills and builds a NEW index every time,
                                  src/database.cpp:170-172, 276-281) */
    int32_t reserved;
} cdb_options;

/* Result of a batched locate, host memory.  Row q (pattern q) is pairs[2*row_off[q] .. 2*row_off[q+1]):
 * consecutive (id, count) int64 pairs — the memory layout of std::vector<std::pair<int64_t,int64_t>>,
 * which is what string_index::query returns (src/index.h:83). */
typedef struct cdb_result {
    int64_t npat;
    int64_t total_pairs;
    int64_t total_occurrences;
    const int64_t* row_off; /* [npat+1] */
    const int64_t* pairs;   /* [2*total_pairs] */
    void* _owner;
} cdb_result;

/* Same, device memory (valid until cdb_device_result_free). */
typedef struct cdb_device_result {
    int64_t npat;
    int64_t total_pairs;
    int64_t total_occurrences;
    int64_t* row_off; /* device [npat+1] */
    int64_t* pairs;   /* device [2*total_pairs] */
    int64_t* left;    /* device [npat]  SA interval [left,right) of every pattern (src/index.cpp:262-287) */
    int64_t* right;   /* device [npat] */
    int32_t* stats32; /* device [2*npat]: row length, then occurrences, of every pattern as 32-bit integers (saturating) —
                         what the shards of a split corpus exchange per batch */
    uint8_t* row_flags; /* device [npat]: bit 0 = some document occurs more than once in the row (otherwise every count is 1) */
    void* _owner;
} cdb_device_result;

/* Highlight spans (replaces the occurrence enumeration of ac_automaton::render, src/database.cpp:58-77):
 * for text t, spans[2*span_off[t] .. 2*span_off[t+1]) are inclusive [begin,end] byte ranges, ascending,
 * overlapping matches merged, touching ones not. */
typedef struct cdb_spans {
    int64_t ntext;
    int64_t total_spans;
    const int64_t* span_off; /* [ntext+1] */
    const int64_t* spans;    /* [2*total_spans] */
    void* _owner;
} cdb_spans;

/* Same, device memory (valid until cdb_device_spans_free). */
typedef struct cdb_device_spans {
    int64_t ntext;
    int64_t total_spans;
    int64_t* span_off; /* device [ntext+1] */
    int64_t* spans;    /* device [2*total_spans] */
    void* _owner;
} cdb_device_spans;

const char* cdb_last_error(void);
const char* cdb_version(void);
int cdb_device_count(void);

/* string_index::string_index() (src/index.h:54).  opts may be NULL. */
cdb_status cdb_create(const cdb_options* opts, cdb_index** out);
/* string_index::~string_index() (src/index.h:79-83) */
void cdb_destroy(cdb_index* idx);

/* string_index::add(int64_t id, std::string_view value) (src/index.cpp:174-177).  The bytes are copied;
 * the caller may release them on return (the reference borrows them, database.cpp:263-264). */
cdb_status cdb_add(cdb_index* idx, int64_t id, const void* value, int64_t len);
/* nd add() calls at once: document d = text[doc_off[d], doc_off[d+1]). */
cdb_status cdb_add_many(cdb_index* idx, const int64_t* ids, const void* text, const int64_t* doc_off, int64_t nd);
/* Loader staging (replaces the std::string-per-field loading of src/database.cpp:173-275): cdb_add /
 * cdb_add_many append to page-locked chunks, and every chunk that fills up is sent to the device while the loader goes
 * on, so cdb_build only stitches the chunks together.  staged = bytes of text added so far, on_device = how many of
 * them already have a device copy (after cdb_build: how many had one when the build started). */
cdb_status cdb_staging_stats(const cdb_index* idx, int64_t* staged, int64_t* on_device);

/* string_index::build() (src/index.cpp:178-236): uploads the staged text and constructs the packed suffix
 * array on the device.  Element = (offset_in_doc << bits) | doc_index, 4 bytes wide iff bits1+bits2 <= 32. */
cdb_status cdb_build(cdb_index* idx);
/* Suffix-array persistence (no reference counterpart: the reference rebuilds its index from the raw/ files at every
 * start and every `build`, src/server.cpp:43-44, src/database.cpp:170-282 — build time is restart latency).
 * cdb_save writes the finished packed array to `path`, keyed by a 64-bit hash of the corpus (text, doc_off, ids), the
 * geometry and the compat flag.  cdb_build_or_load is cdb_build for the staged corpus, except that a file at `path`
 * whose key matches is read back instead of sorting; a missing, stale or truncated file means a normal build.
 * *loaded (may be NULL) = 1 when the array came from the file. */
cdb_status cdb_save(const cdb_index* idx, const char* path);
cdb_status cdb_build_or_load(cdb_index* idx, const char* path, int32_t* loaded);
/* Same, from a corpus already resident in device memory (text[n], doc_off[nd+1], ids[nd] are BORROWED and
 * must outlive the index).  `stream` is a cudaStream_t (NULL = default stream).
 * Contract of the borrowed buffers: doc_off[0] == 0 and doc_off is non-decreasing; d_text is 16-byte aligned and has at
 * least 64 READABLE bytes after its last document, i.e. the allocation is >= doc_off[nd] + 64 bytes (their content
 * does not matter): the kernels fetch text in aligned 8- and 16-byte windows, and a window that starts inside the
 * last document may reach past its end.  cdb_build pads its own copy; a caller of this entry point must allocate
 * the padding itself. */
cdb_status cdb_build_device(cdb_index* idx, const void* d_text, const int64_t* d_doc_off, const int64_t* d_ids,
                            int64_t nd, void* stream);

/* Index geometry after build: SA length n, element width (4|8), doc-index field width and mask
 * (src/index.h:56-58). */
cdb_status cdb_info(const cdb_index* idx, int64_t* n, int64_t* nd, int32_t* width, int32_t* bits, uint64_t* mask);
/* The prefix directory built beside the suffix array (no reference counterpart): `symbols` leading symbols of
 * `bits_per_symbol` bits each index `entries` + 1 interval boundaries; symbols = 0 when the index has none (tiny
 * corpora, or the note-N1 layout where the reference's exact binary search must run). */
cdb_status cdb_prefix_directory(const cdb_index* idx, int32_t* symbols, int32_t* bits_per_symbol, int64_t* entries);
/* The document listing built beside the prefix directory (no reference counterpart): for every directory
 * bucket of at most 1024 suffixes, the ids of the suffixes' documents in the order query() reports them (order 0: ascending
 * doc index, src/index.cpp:288-322) or filter() merges them (order 1: ascending id, src/interface.cpp:82,88), 4 + hi_bytes
 * bytes per suffix.  A keyword of exactly `symbols` (cdb_prefix_directory) symbols is answered by streaming its bucket.
 * Order 0 is built by cdb_build when device memory allows, order 1 on the first id-ordered call (cdb_filter); when there
 * is no room for both, the one not in use is dropped.  present = 0: none (no directory, ids spanning more than 2^63,
 * two documents with one id, no memory, or CDB_LISTING=0) — every keyword then takes the suffix-array path. */
cdb_status cdb_listing_info(const cdb_index* idx, int32_t order, int32_t* present, int32_t* hi_bytes, int64_t* bytes, double* build_ms);
/* Copies the packed suffix array to host memory (n*width bytes), for parity checks. */
cdb_status cdb_export_sa(const cdb_index* idx, void* buf, int64_t buf_bytes);
/* Device pointer of the packed suffix array (borrowed). */
cdb_status cdb_sa_device_ptr(const cdb_index* idx, const void** d_sa);

/* Independent device-side check of the built suffix array (test / bench instrumentation; shares no code with the build).
 * Every adjacent pair is compared with the reference's comparator (std::string_view <, src/index.cpp:92-93; in the
 * note-N1 layout the signed rule of src/index.h:66-73 for groups larger than chuck_size, src/index.cpp:97-125), and the
 * array is checked to be a permutation of all (doc, offset) pairs.
 * out8 = {inversions, invalid elements, duplicate positions, ties (adjacent byte-identical suffixes), ties not in
 * ascending packed order, pairs queued for the signed-rule check, queued pairs left unchecked, pairs checked under the
 * signed rule}.  A correct array has out8[0] = out8[1] = out8[2] = out8[4] = out8[6] = 0. */
cdb_status cdb_verify_sa(const cdb_index* idx, int64_t* out8);
/* Element-wise comparison of the built array with another packed suffix array of the same corpus in host memory (n
 * elements of the index's width — e.g. the array the compiled reference built): out3 = {identical elements, different
 * elements whose two suffixes are byte-identical (SURVEY.md note N2 ties), elements naming different suffixes}.
 * "Identical up to ties" is out3[2] == 0. */
cdb_status cdb_compare_sa(const cdb_index* idx, const void* other_sa, int64_t other_bytes, int64_t* out3);

/* Batched string_index::query(keyword) (src/index.cpp:237-326): pattern q = pat[pat_off[q], pat_off[q+1]).
 * An empty pattern fails the whole batch with CDB_ERR_EMPTY_KEYWORD.  Re-entrant: may be called concurrently
 * from several host threads on the same index (the reference's query() is const and runs under a shared
 * lock, database.cpp:387-393). */
cdb_status cdb_locate_batch(const cdb_index* idx, const void* pat, const int64_t* pat_off, int64_t npat, cdb_result* out);
void cdb_result_free(cdb_result* r);
/* string_index::query(keyword) for ONE keyword, the call the reference's server makes (src/database.cpp:387-393, once
 * per keyword of a request — src/interface.cpp:79-86 — from up to max(8, hw-1) worker threads at a time).  Concurrent
 * callers on the same index are coalesced into one device batch ("group commit", SURVEY.md §8f-2,
 * coffeedb_b200/host/micro_batcher.hpp): whoever arrives while a batch is on the device joins the next one; a lone
 * caller on an idle device is served at once.  `out` is a one-row result (npat = 1) that views the caller's row of
 * the shared batch result; release it with cdb_result_free.  An empty keyword fails this caller only
 * (CDB_ERR_EMPTY_KEYWORD); a device failure fails every member of the affected batch.
 * Environment: CDB_QUERY_MAX_BATCH (65536), CDB_QUERY_IN_FLIGHT (2), CDB_QUERY_LINGER_US (0). */
cdb_status cdb_query(const cdb_index* idx, const void* keyword, int64_t len, cdb_result* out);
/* Counters of cdb_query on this index: keywords submitted, device batches issued, keywords in the largest batch. */
cdb_status cdb_query_stats(const cdb_index* idx, uint64_t* queries, uint64_t* batches, uint64_t* largest);
/* Same with patterns and results resident in device memory; work is enqueued on `stream` and the call
 * returns after the stream has drained (result sizes are data dependent). */
cdb_status cdb_locate_batch_device(const cdb_index* idx, const void* d_pat, const int64_t* d_pat_off, int64_t npat,
                                   void* stream, cdb_device_result* out);
/* Same, with a hook for sharded callers: `rows_ready(user, stats32, npat, stream)` runs on the calling host
 * thread as soon as the kernels that produce stats32 — per-pattern row length and occurrences, device [2*npat] — have been
 * enqueued.  The callback records an event on the stream it is HANDED — the caller's `stream`, or, when the rows are being
 * streamed from the document listing at that moment, a side stream of the library that only waits for stats32 — and starts
 * its exchange of stats32 (an NCCL all_gather) on another stream, where it runs under the rest of the locate. */
typedef void (*cdb_rows_ready_fn)(void* user, const int32_t* stats32, int64_t npat, void* stream);
cdb_status cdb_locate_batch_device_ex(const cdb_index* idx, const void* d_pat, const int64_t* d_pat_off, int64_t npat,
                                      void* stream, cdb_rows_ready_fn rows_ready, void* user, cdb_device_result* out);
void cdb_device_result_free(cdb_device_result* r);

/* Highlight: merged occurrence spans of the keyword set kw (nkw keywords, kw_off[nkw+1]) inside each of the
 * documents docs[0..ndocs) (doc indices).  Replaces ac_automaton::render's span loop (database.cpp:58-77);
 * marker splicing (database.cpp:78-90) stays on the host (cdb_splice). */
cdb_status cdb_locate_spans(const cdb_index* idx, const void* kw, const int64_t* kw_off, int64_t nkw,
                            const int64_t* docs, int64_t ndocs, cdb_spans* out);
/* Batched highlight (src/database.cpp:139-165, 401-432: renderer / select() call render() once per selected object and
 * constrained key).  nreq requests, request r owning the keywords [req_kw_off[r], req_kw_off[r+1]) of the keyword list
 * (kw, kw_off[nkw+1]); ntext texts, text t = document text_doc[t] (doc index) highlighted with the keywords of request
 * text_req[t].  One launch sequence and one synchronisation for the whole batch; spans of text t are
 * spans[2*span_off[t] .. 2*span_off[t+1]).  A document may appear in any number of texts. */
cdb_status cdb_locate_spans_batch(const cdb_index* idx, const void* kw, const int64_t* kw_off, int64_t nkw,
                                  const int64_t* req_kw_off, int64_t nreq, const int64_t* text_req, const int64_t* text_doc,
                                  int64_t ntext, cdb_spans* out);
/* Same with every array resident in device memory (the keyword bytes readable 8 bytes past their end); work is
 * enqueued on `stream` and the call returns after the stream has drained. */
cdb_status cdb_locate_spans_batch_device(const cdb_index* idx, const void* d_kw, const int64_t* d_kw_off, int64_t nkw,
                                         const int64_t* d_req_kw_off, int64_t nreq, const int64_t* d_text_req,
                                         const int64_t* d_text_doc, int64_t ntext, void* stream, cdb_device_spans* out);
void cdb_device_spans_free(cdb_device_spans* s);
void cdb_spans_free(cdb_spans* s);
/* database.cpp:78-90: writes text with left/right spliced around the spans into out (capacity out_cap);
 * returns the rendered length (also when out is too small or NULL). */
int64_t cdb_splice(const void* text, int64_t tlen, const int64_t* spans, int64_t nspans, const void* left, int64_t llen,
                   const void* right, int64_t rlen, void* out, int64_t out_cap);

/* ---- one string_index over several GPUs of this process (SURVEY.md 8e behind the boundary) ----------------------------
 * The reference is one process (src/database.cpp:170-282, 387-393); a sharded index therefore sits behind the same
 * add / build / query calls.  Documents are cut into contiguous doc-index ranges of about equal bytes, one shard per entry
 * of devices[] (an ordinal may repeat: several shards on one GPU), every shard is built by its own host thread, and a
 * batch is located on all shards at once: the packed patterns are uploaded once and copied device-to-device, the shards
 * exchange their per-pattern row offsets device-to-device and each writes its part of every row straight into the
 * result — row q = the shard rows in shard order = ascending doc index = string_index::query on the whole corpus
 * (src/index.cpp:316-322).  On corpora in the note-N1 layout (bytes on both sides of 0x80, SURVEY.md 8e "Exception") the
 * answer is the per-shard reference answers concatenated, as with any sharded index. */
typedef struct cdb_sharded cdb_sharded;
cdb_status cdb_sharded_create(const int32_t* devices, int32_t ndev, const cdb_options* opts, cdb_sharded** out);
void cdb_sharded_destroy(cdb_sharded* s);
cdb_status cdb_sharded_add(cdb_sharded* s, int64_t id, const void* value, int64_t len);
cdb_status cdb_sharded_add_many(cdb_sharded* s, const int64_t* ids, const void* text, const int64_t* doc_off, int64_t nd);
cdb_status cdb_sharded_build(cdb_sharded* s);
/* Batched string_index::query over all shards; the result is released with cdb_result_free. */
cdb_status cdb_sharded_locate_batch(const cdb_sharded* s, const void* pat, const int64_t* pat_off, int64_t npat, cdb_result* out);
/* Shard g (borrowed; valid until cdb_sharded_destroy) and the doc-index range [doc_begin, doc_end) it holds. */
cdb_status cdb_sharded_shard(const cdb_sharded* s, int32_t g, cdb_index** shard, int64_t* doc_begin, int64_t* doc_end);
int32_t cdb_sharded_count(const cdb_sharded* s);

/* ---- filter(): set algebra over constraint results, on the device -----------------------------------
 * Replaces the body of filter() (src/interface.cpp:46-147) and what it calls: query(key, range) of string keys
 * (src/database.cpp:387-393 -> string_index::query) and of integer / double keys (numeric_query, src/index.cpp:63-74).
 * Per request: the results of the ranges of ONE key are OR-merged by id with their counts added (interface.cpp:79-112),
 * the keys are AND-intersected with counts added (:113-134, numeric keys contribute 0, src/index.cpp:71), "$correlation"
 * keeps sums in [L, R) (:136-142), the answer is ordered by std::sort on descending $correlation (:143-146 — unstable:
 * the device reproduces libstdc++'s permutation of the id-ascending input exactly, coffeedb_b200/host/std_sort_order.hpp)
 * and `span` cuts result[span0, span1) (interface.cpp:196-209).  Only that slice crosses PCIe. */
typedef struct cdb_numeric cdb_numeric; /* replaces one integer_index / double_index (src/index.h:29-53) */
/* add() for every (id, value) + build() (src/index.cpp:152-173).  kind 0: values are int64_t, kind 1: double. */
cdb_status cdb_numeric_create(int32_t kind, const int64_t* ids, const void* values, int64_t n, int32_t device,
                              cdb_numeric** out);
void cdb_numeric_destroy(cdb_numeric* c);
/* integer_index::query / double_index::query (src/index.cpp:159-161, 170-172 -> numeric_query :63-74): a one-row result,
 * the (id, 0) pairs of data[lower_bound(lo), lower_bound(hi)) in (value, id) order.  lo / hi = {value bits, id} exactly
 * as parse_range builds its two pairs (src/utility.h:69-86): the id is INT64_MAX for "(" on the left / "]" on the
 * right and 0 otherwise; value bits are the int64_t, or the bit pattern of the double. */
cdb_status cdb_numeric_query(const cdb_numeric* c, const int64_t lo[2], const int64_t hi[2], cdb_result* out);

typedef struct cdb_filter_key {
    int32_t kind;      /* 0: index is a built cdb_index*; 1: a cdb_numeric*; -1: a key the database does not have — its
                          query() answers {} (src/database.cpp:389-391), so a request that names it matches nothing */
    int32_t reserved;
    const void* index;
} cdb_filter_key;
typedef struct cdb_filter_term {   /* one range string of one key of one request (interface.cpp:56-74) */
    int32_t key;       /* slot in keys[] */
    int32_t range;     /* numeric key: index into ranges[]; string key: -1 */
    int64_t kw_begin;  /* string key: the keyword is kw[kw_begin, kw_end) */
    int64_t kw_end;
} cdb_filter_term;
typedef struct cdb_filter_batch {
    const cdb_filter_key* keys;
    int32_t nkeys;               /* <= 32 */
    int32_t reserved;
    const void* kw;              /* keyword bytes of all string terms */
    int64_t kw_len;
    const int64_t* ranges;       /* [4 * nranges] {lo value bits, lo id, hi value bits, hi id}, see cdb_numeric_query */
    int64_t nranges;
    const cdb_filter_term* terms;     /* request r owns terms [req_term_off[r], req_term_off[r+1]), at most 64; terms with
                                         the same key are OR-ed, different keys AND-ed; a request without terms matches
                                         nothing (the reference answers it from its object table, interface.cpp:51-53) */
    const int64_t* req_term_off;      /* [nreq + 1] */
    int64_t nreq;
    const int64_t* corr_range;        /* NULL or [2 * nreq]: keep L <= $correlation < R (parse_uint_range, utility.h:87-104) */
    const int64_t* span;              /* NULL or [2 * nreq]: result[span0, span1) */
} cdb_filter_batch;
/* Request r: pairs[2*row_off[r] .. 2*row_off[r+1]) = (id, $correlation) in the reference's order; matched[r] = size of
 * its answer before the span (what `count` reports, interface.cpp:243-262). */
typedef struct cdb_filter_result {
    int64_t nreq;
    int64_t total_pairs;
    const int64_t* row_off; /* [nreq+1] */
    const int64_t* pairs;   /* [2*total_pairs] */
    const int64_t* matched; /* [nreq] */
    void* _owner;
} cdb_filter_result;
/* An empty keyword fails the batch with CDB_ERR_EMPTY_KEYWORD (src/index.cpp:239-241).  Re-entrant. */
cdb_status cdb_filter(const cdb_filter_batch* batch, cdb_filter_result* out);
void cdb_filter_result_free(cdb_filter_result* r);

/* Timing of the last build (milliseconds, CUDA events): total and the radix-sort share; refinement rounds
 * and number of key-range chunks. */
cdb_status cdb_build_stats(const cdb_index* idx, double* total_ms, double* sort_ms, int64_t* rounds, int64_t* chunks);
/* Phase timing of the calling thread's last locate (milliseconds, CUDA events on the launching stream):
 * ms6 = {search (incl. the occurrence scan), gather (phase A, incl. the row-length scan), large-interval path,
 * tail (large-path emit + total read-back), translate (phase B), total}; counts4 = {npat, pairs, occurrences,
 * patterns that took the large-interval path}. */
void cdb_last_locate_stats(double* ms6, int64_t* counts4);
/* Same with the listing phase: ms8 = {the six above, listing (rows streamed from the document listing), 0};
 * counts8 = {the four above, rows answered from the listing, their result pairs, 0, 0}. */
void cdb_last_locate_stats_ex(double* ms8, int64_t* counts8);
/* Number of CUDA kernels this library has launched in this process. */
uint64_t cdb_launch_count(void);
/* Returns the idle pinned host buffers the library keeps for result re-use to the driver (the pool is also capped:
 * CDB_PINNED_POOL_MB, default 32 GB, at most 64 buffers). */
void cdb_trim(void);

#ifdef __cplusplus
}
#endif
#endif
