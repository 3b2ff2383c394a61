#!/bin/bash
mkdir -p gpurun_out
tools/_build/port_microbench > gpurun_out/r02g_port_microbench.txt 2>&1
CDB_CFG5_BYTES=50000000 timeout 600 python -m pytest tests/test_gpu_cfg5.py -x -q -s 2>&1 | tail -12 > gpurun_out/r02h_cfg5_small.txt
timeout 1500 python -m pytest tests/test_gpu_cfg5.py -x -q -s 2>&1 | tail -15 > gpurun_out/r02h_cfg5_full.txt
timeout 900 python bench.py --steps 5 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
cat gpurun_out/r02g_port_microbench.txt gpurun_out/r02h_cfg5_small.txt gpurun_out/r02h_cfg5_full.txt
tail -c 6000 gpurun_out/r02h_bench.json; tail -5 gpurun_out/r02h_bench.err
