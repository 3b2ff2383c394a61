#!/bin/bash
# BASELINE configs[4] at size: N ranks x 5 GB of UTF-8, substring AND numeric range through cdb_filter per shard, NCCL count merge
N=${1:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
    bench.py --cfg5 --gpus $N > gpurun_out/r02w_cfg5_n$N.json 2> gpurun_out/r02w_cfg5_n$N.err
tail -c 3000 gpurun_out/r02w_cfg5_n$N.json; grep -v "^\*\|OMP_NUM\|Warning\|warn" gpurun_out/r02w_cfg5_n$N.err | tail -8
