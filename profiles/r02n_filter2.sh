#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_filter.py tests/test_gpu_parity.py -q -x 2>&1 | tail -5 > gpurun_out/r02n_tests.txt
CDB_DEBUG_TIMING=1 timeout 900 python bench.py --workload cfg3 --steps 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify > gpurun_out/r02n_bench_dbg.json 2> gpurun_out/r02n_bench_dbg.err
timeout 900 python bench.py --workload cfg3 --steps 5 --no-rebuild --no-extras --no-spans --no-verify > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
cat gpurun_out/r02n_tests.txt; grep "cdb_filter" gpurun_out/r02n_bench_dbg.err | tail -9
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02n_bench.json').read().strip().splitlines()[-1])
print('e2e', j['e2e']['value'], 'filter', json.dumps(j['filter'])[:700])
print('cpu', json.dumps(j['cpu_baseline'].get('filter_span'))[:400])
PY
