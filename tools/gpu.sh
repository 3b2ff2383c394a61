#!/bin/bash
# builds everything that travels to the GPU box, then runs one gpurun call:  tools/gpu.sh <timeout-seconds> '<command>' [--gpus N]
set -e
cd /root/repo
make -C coffeedb_b200/csrc -j8 2>&1 | grep -E "error|warning|Error" && exit 1
T=$1; shift
CMD=$1; shift
exec /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
