#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multidevice.py tests/test_host_adaptor.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02p_tests.txt
cat gpurun_out/r02p_tests.txt
