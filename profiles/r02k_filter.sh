#!/bin/bash
# device-side filter against the reference server, the sliced verifier at cfg5 shape, the segmented L2 micro-benchmark
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py -q -x 2>&1 | tail -30 > gpurun_out/r02k_filter.txt
tools/_build/l2_microbench quick > gpurun_out/r02k_l2seg.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_cfg5.py -q -k "verifier or geometry" 2>&1 | tail -8 > gpurun_out/r02k_cfg5.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -5 > gpurun_out/r02k_parity.txt
cat gpurun_out/r02k_filter.txt gpurun_out/r02k_l2seg.txt gpurun_out/r02k_cfg5.txt gpurun_out/r02k_parity.txt
