#!/bin/bash
# first run of the document listing: its tests, the parity and filter suites (listing on by default), a short cfg3 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_listing.py tests/test_gpu_parity.py tests/test_gpu_filter.py -m gpu -q -x --durations=5 2>&1 | tail -25 > gpurun_out/r02L_tests.txt
cat gpurun_out/r02L_tests.txt
( time CDB_DEBUG_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras ) > gpurun_out/r02L_bench.json 2> gpurun_out/r02L_bench.err
tail -c 1500 gpurun_out/r02L_bench.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r02L_bench.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4g e2e %.4g ms/step %.3f' % (j['value'], j['e2e']['value'], j['ms_per_step']))
print('roofline', json.dumps(j['roofline'])[:1800])
print('sa_path', json.dumps(j.get('sa_path'))[:900])
print('filter', json.dumps(j.get('filter'))[:900])
print('build', json.dumps(j.get('build'))[:600])
PY
