#pragma once
#include "index.cuh"

namespace cdb {
// verify.cu: independent check of the finished suffix array (adjacent-pair order under the reference's comparator,
// permutation, ties).  out = {inversions, invalid elements, duplicate positions, ties, ties out of ascending packed
// order, pairs queued for the signed-rule (note N1) check, queued pairs left unchecked, pairs checked under the signed rule}
void verify_index(const Index& ix, cudaStream_t st, i64 out[8]);
}  // namespace cdb
