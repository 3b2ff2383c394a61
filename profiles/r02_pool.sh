#!/bin/bash
# coalescing queue of cdb_query: batches in flight / linger, C++ worker threads calling string_index::query (1 GB index, 5-byte keywords)
mkdir -p gpurun_out
{
for inflight in 1 2 4; do for linger in 0 20; do
  echo "== CDB_QUERY_IN_FLIGHT=$inflight CDB_QUERY_LINGER_US=$linger"
  CDB_QUERY_IN_FLIGHT=$inflight CDB_QUERY_LINGER_US=$linger timeout 300 tools/_build/query_pool_bench 10000000 100 5 400 | grep threads
done; done
} > gpurun_out/r02_pool.txt 2>&1
cat gpurun_out/r02_pool.txt
