#!/bin/bash
# N GPUs with the document listing: NCCL row parity tests, the multi-device index, bench at N ranks (as the driver launches it)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_gpu_multidevice.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r02R_tests_n$N.txt
cat gpurun_out/r02R_tests_n$N.txt
CDB_SHARD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-rebuild --no-verify > gpurun_out/r02R_bench_n$N.json 2> gpurun_out/r02R_bench_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
j=json.loads(open(f'gpurun_out/r02R_bench_n{N}.json').read().strip().splitlines()[-1])
print('N', N, 'value %.4g ms/step %.3f' % (j['value'], j['ms_per_step']), j['roofline']['phases_ms'], 'e2e %.4g' % j['e2e']['value'], j.get('parity_sharded'), 'build', j['build']['ms'])
print('sa_path', j['sa_path']['value'], j['sa_path']['ms_per_step'], j['sa_path']['phases_ms'])
PY
grep "broadcast / locate" gpurun_out/r02R_bench_n$N.err | head -4
grep -v "^\*\|OMP_NUM\|broadcast / locate" gpurun_out/r02R_bench_n$N.err | tail -4
