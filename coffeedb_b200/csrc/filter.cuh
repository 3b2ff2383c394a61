#pragma once
#include "index.cuh"

namespace cdb {

// One integer_index / double_index of the reference (src/index.h:29-53): (value, id) pairs sorted by value then id
// (src/index.cpp:155-158, 166-169), kept twice on the device — in that order (numeric_query's two lower bounds,
// src/index.cpp:63-74) and in id order (membership test of an id inside filter's intersections).  Values are stored as
// order-preserving u64 keys (int64: sign bit flipped; double: the usual sign fold, -0.0 folded onto +0.0).
struct NumericIndex {
    int device = 0;
    int kind = 0;  // 0 int64, 1 double
    i64 n = 0;
    u64* vkey = nullptr;  // (value, id) order
    i64* vid = nullptr;
    u64* ikey = nullptr;  // id order
    i64* iid = nullptr;
    ~NumericIndex();
};
NumericIndex* numeric_create(int kind, const i64* ids, const void* values, i64 n, int device, cudaStream_t st);
// positions [begin, end) of numeric_query's answer inside the (value, id) order; lo / hi = {value bits, id}
void numeric_bounds(const NumericIndex& c, const i64 lo[2], const i64 hi[2], cudaStream_t st, i64* begin, i64* end);

// Result of filter_batch_device, device memory.  Request r owns fin[2*fin_off[r] .. 2*fin_off[r+1]): (id, $correlation)
// pairs in the reference's final order with the request's span applied — except for the requests listed in `pending`,
// whose slots are left for the host: their id-ascending survivors sit at raw[2*raw_off[r] ..) (raw_len[r] pairs) and the
// caller finishes them with the one std::sort of src/interface.cpp:143-146 and the span.
struct FilterOut {
    i64 nreq = 0;
    u64 total_fin = 0;
    DevBuf<i64> fin;
    DevBuf<u64> fin_off;   // [nreq + 1]
    DevBuf<u64> matched;   // [nreq] size of the answer before the span
    DevBuf<i64> raw;
    DevBuf<u64> raw_off;   // [nreq + 1]
    DevBuf<u64> raw_len;   // [nreq]
    std::vector<i64> pending;
};
void filter_batch_device(const cdb_filter_batch& b, cudaStream_t st, FilterOut& out);
int filter_device_of(const cdb_filter_batch& b);  // the device every key of the batch lives on (throws when they differ)

}  // namespace cdb
