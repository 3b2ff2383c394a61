#!/bin/bash
# cdb_query's coalescing queue after the wake-up rework (followers no longer wake through the queue's mutex), pooled small-batch
# buffer sets, malloc'ed small results: C++ worker threads calling string_index::query (1 GB index, 5-byte keywords, warm threads,
# 2000 calls per thread), batches in flight 1 / 2 / 4
mkdir -p gpurun_out
{
for inflight in 2 1 4; do
  echo "== CDB_QUERY_IN_FLIGHT=$inflight"
  CDB_QUERY_IN_FLIGHT=$inflight timeout 300 tools/_build/query_pool_bench 10000000 100 5 2000 | grep threads
done
} > gpurun_out/r02T_pool.txt 2>&1
cat gpurun_out/r02T_pool.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "small or concurrent or query or known" 2>&1 | tail -2
