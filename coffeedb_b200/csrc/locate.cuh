#pragma once
#include "index.cuh"

namespace cdb {
// Batched locate with patterns and results in device memory (see locate.cu).
void locate_device(const Index& ix, const u8* d_pat, const i64* d_pat_off, i64 npat, cudaStream_t st,
                   cdb_device_result* out);
// Highlight spans (see spans.cu).
void locate_spans(const Index& ix, const u8* kw, const i64* kw_off, i64 nkw, const i64* docs, i64 ndocs,
                  cudaStream_t st, std::vector<i64>& span_off, std::vector<i64>& spans);
}  // namespace cdb
