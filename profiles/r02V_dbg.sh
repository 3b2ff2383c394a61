#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_listing.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r02V_dbg.txt
cat gpurun_out/r02V_dbg.txt
( CDB_DEBUG_TIMING=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-extras --no-spans --no-cpu-baseline --no-verify --no-rebuild ) > gpurun_out/r02V_1.json 2> gpurun_out/r02V_1.err
grep "document listing" gpurun_out/r02V_1.err | head -4
