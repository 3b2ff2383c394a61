#pragma once
#include "index.cuh"

namespace cdb {
// verify.cu: independent check of the finished suffix array (adjacent-pair order under the reference's comparator,
// permutation, ties).  out = {inversions, invalid elements, duplicate positions, ties, ties out of ascending packed
// order, pairs queued for the signed-rule (note N1) check, queued pairs left unchecked, pairs checked under the signed rule}
void verify_index(const Index& ix, cudaStream_t st, i64 out[8]);
// Element-wise comparison with another packed array of the same corpus held in host memory (same element width):
// out = {identical elements, different elements whose suffixes are byte-identical (ties), different suffixes}
void compare_index_sa(const Index& ix, const void* host_sa, cudaStream_t st, i64 out[3]);
}  // namespace cdb
