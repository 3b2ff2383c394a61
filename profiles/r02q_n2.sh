#!/bin/bash
# two GPUs: NCCL world-2 row parity, the single-process multi-device index on two real devices, bench at N=1 and N=2 on the same box
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_gpu.py tests/test_gpu_multidevice.py -m gpu -q 2>&1 | tail -4 > gpurun_out/r02q_tests_n$N.txt
cat gpurun_out/r02q_tests_n$N.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify --no-filter > gpurun_out/r02q_bench_n1.json 2> gpurun_out/r02q_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-rebuild --no-verify > gpurun_out/r02q_bench_n$N.json 2> gpurun_out/r02q_bench_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for n in ('1', N):
    try:
        j=json.loads(open(f'gpurun_out/r02q_bench_n{n}.json').read().strip().splitlines()[-1])
        print('N', n, 'value %.4g ms/step %.3f' % (j['value'], j['ms_per_step']), j['roofline']['phases_ms'], 'e2e %.4g' % j['e2e']['value'], j.get('parity_sharded'), 'build', j['build']['ms'])
    except Exception as e: print(n, 'failed', e)
PY
grep -v "^\*\|OMP_NUM" gpurun_out/r02q_bench_n$N.err | tail -5
