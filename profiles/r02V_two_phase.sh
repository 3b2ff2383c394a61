#!/bin/bash
# two-phase listing build (sorted keys + split points, then ids by doc range): tests, build time at cfg3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_listing.py tests/test_gpu_filter.py -m gpu -q -x 2>&1 | tail -3
for tp in 1 0; do
( CDB_LISTING_TWO_PHASE=$tp CDB_DEBUG_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-spans --no-cpu-baseline --no-verify ) > gpurun_out/r02V_$tp.json 2> gpurun_out/r02V_$tp.err
grep "document listing" gpurun_out/r02V_$tp.err | head -4
python - $tp <<'PY'
import json,sys
j=json.loads([l for l in open(f'gpurun_out/r02V_{sys.argv[1]}.json').read().strip().splitlines() if l.startswith('{')][-1])
print('two_phase', sys.argv[1], 'value %.4g e2e %.4g build %.0f rebuild %.0f listing %s' % (j['value'], j['e2e']['value'], j['build']['ms'], j['build']['rebuild_ms'], j['roofline']['listing']))
PY
done
