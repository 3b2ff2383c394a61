#pragma once
#include "index.cuh"

namespace cdb {
// Batched locate with patterns and results in device memory (see locate.cu).
// id_order: rows in ascending id instead of ascending doc index (what filter() merges, src/interface.cpp:79-135): the
// interval's doc indices are mapped to id ranks before they are sorted, and translated through ids_by_rank.
// rows_ready: called on the host once the kernels that produce stats32 (per-pattern row length and occurrences) have been
// enqueued, before translate is — a sharded caller launches its exchange on another stream from there (cdb_locate_batch_device_ex).
void locate_device(const Index& ix, const u8* d_pat, const i64* d_pat_off, i64 npat, cudaStream_t st,
                   cdb_device_result* out, bool id_order = false, cdb_rows_ready_fn rows_ready = nullptr,
                   void* rows_ready_user = nullptr);
// Small-batch fast path (locate.cu, experimental): one upload, two launches, one synchronisation.  On success the
// rows are in mapped pinned memory of the calling thread — valid until its next call: row q has rowlen[q] pairs at
// pairs + 2 * (rowocc[0] + ... + rowocc[q-1]).  Returns false when the batch has to take the general path.
struct SmallResult {
    i64 total_occ = 0;
    const u64* rowlen = nullptr;
    const u64* rowocc = nullptr;
    const i64* pairs = nullptr;
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
