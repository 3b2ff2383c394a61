#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py -q -x 2>&1 | tail -3 > gpurun_out/r02u_tests.txt
CDB_DEBUG_TIMING=1 timeout 900 python bench.py --workload cfg3 --steps 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify > gpurun_out/r02u_bench_dbg.json 2> gpurun_out/r02u_bench_dbg.err
timeout 900 python bench.py --steps 5 --no-rebuild > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
cat gpurun_out/r02u_tests.txt; grep "cdb_filter" gpurun_out/r02u_bench_dbg.err | tail -9
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'filter', j['filter']['ms_per_step'], j['filter']['ms_each_step'])
print('cpu', json.dumps(j['cpu_baseline'])[:600])
print('cfg5_shard', json.dumps(j['extras'].get('cfg5_shard'))[:1500])
print('traffic', j['roofline']['traffic'], j['roofline']['traffic_source'][:60])
PY
tail -3 gpurun_out/r02u_bench.err
