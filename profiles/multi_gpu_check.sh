#!/bin/bash
# Multi-GPU validation (run under `gpurun --gpus N`): NCCL sharded tests + the bench at N ranks, launched exactly as
# the driver launches it.
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/pytest_sharded_gpu.log 2>&1; tail -2 gpurun_out/pytest_sharded_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1800 gpurun_out/bench_n$N.json; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
