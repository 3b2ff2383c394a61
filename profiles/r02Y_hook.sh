#!/bin/bash
# sharded step on short rows (2 GPUs, cfg2 corpus = rows of ~42 entries per shard, 10^6 keywords): the all_gather hook before the emit
# kernel on the launching stream (0) against emit first + hook on a side stream that only waits for the statistics (1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q -x 2>&1 | tail -2
for late in 0 1; do
CDB_HOOK_AFTER_EMIT=$late CDB_SHARD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --workload cfg2 --npat 1000000 --steps 20 --warmup 5 --no-cpu-baseline --no-rebuild --no-verify --no-extras --no-spans --no-filter > gpurun_out/r02Y_$late.json 2> gpurun_out/r02Y_$late.err
python - $late <<'PY'
import json,sys
j=json.loads(open(f'gpurun_out/r02Y_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('hook after emit =', sys.argv[1], 'value %.4g ms/step %.3f' % (j['value'], j['ms_per_step']), {k: round(v,3) for k,v in j['roofline']['phases_ms'].items()}, j.get('parity_sharded',{}).get('status'))
PY
grep "broadcast / locate" gpurun_out/r02Y_$late.err | head -2
done
