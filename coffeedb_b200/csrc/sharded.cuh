#pragma once
#include <functional>

#include "index.cuh"

namespace cdb {
struct ShardedIndex;  // sharded.cu: one string_index over several devices of this process
ShardedIndex* sharded_create(const int32_t* devices, int32_t ndev, const cdb_options* opts);
void sharded_destroy(ShardedIndex* s);
void sharded_add_many(ShardedIndex* s, const i64* ids, const u8* text, const i64* doc_off, i64 nd);
void sharded_build(ShardedIndex* s);
void sharded_locate(ShardedIndex* s, const u8* pat, const i64* pat_off, i64 npat,
                    const std::function<void(i64 total_pairs, i64** row_off, i64** pairs)>& alloc, i64* total_pairs_out,
                    i64* total_occ_out);
int sharded_count(const ShardedIndex* s);
cdb_index* sharded_shard(const ShardedIndex* s, int g, i64* doc_begin, i64* doc_end);
}  // namespace cdb
