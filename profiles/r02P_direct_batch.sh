#!/bin/bash
# cdb_filter: one-keyword requests 32 per warp (filter_direct_batch_kernel), rows with repeats through the list-driven kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_listing.py tests/test_gpu_filter.py tests/test_dropin_server.py -m gpu -q -x 2>&1 | tail -4
( CDB_DEBUG_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-spans --no-verify ) > gpurun_out/r02P.json 2> gpurun_out/r02P.err
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r02P.json').read().strip().splitlines() if l.startswith('{')][-1])
f=j['filter']
print('value %.4g ms/step %.3f' % (j['value'], j['ms_per_step']))
print('filter e2e %.4g ms/step %.2f each %s locate %s launches %s' % (f['value'], f['ms_per_step'], f['ms_each_step'], f['locate_phases_ms'], f['launches_per_step']))
print('cpu', json.dumps(j['cpu_baseline'])[:900])
PY
grep "cdb_filter" gpurun_out/r02P.err | tail -9
