#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q 2>&1 | tail -2
CDB_SHARD_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-rebuild --no-verify > gpurun_out/r02q_trace_n$N.json 2> gpurun_out/r02q_trace_n$N.err
grep "all_gather ms" gpurun_out/r02q_trace_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-rebuild --no-verify > gpurun_out/r02q_over_n$N.json 2> gpurun_out/r02q_over_n$N.err
python - $N <<'PY'
import json,sys
for f in ('trace','over'):
    j=json.loads(open(f'gpurun_out/r02q_{f}_n{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print(f, 'value %.4g ms/step %.3f' % (j['value'], j['ms_per_step']), j['roofline']['phases_ms'], j['parity_sharded'])
PY
grep -i "error\|Traceback" gpurun_out/r02q_over_n$N.err | head -5
