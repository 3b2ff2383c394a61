#!/bin/bash
# where the cdb_filter e2e step spends its time at cfg3 (CDB_DEBUG_TIMING), + new tests, + N=1 sanity of the persistent gather
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_filter.py tests/test_sharded_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r02m_tests.txt
CDB_DEBUG_TIMING=1 timeout 900 python bench.py --workload cfg3 --steps 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
cat gpurun_out/r02m_tests.txt; grep "cdb_filter" gpurun_out/r02m_bench.err | tail -24; tail -c 1500 gpurun_out/r02m_bench.json
