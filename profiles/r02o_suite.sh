#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -25 > gpurun_out/r02o_tests.txt
cat gpurun_out/r02o_tests.txt
