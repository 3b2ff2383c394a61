#!/bin/bash
# full GPU test suite with durations, then the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -40 > gpurun_out/r02i_tests.txt
timeout 900 python bench.py --steps 5 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
cat gpurun_out/r02i_tests.txt
tail -c 7000 gpurun_out/r02i_bench.json; tail -5 gpurun_out/r02i_bench.err
