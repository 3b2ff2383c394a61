#!/bin/bash
# lazy listed rows in cdb_filter (rows answered from the listing are not written unless a merge needs them) + the L2 fetch-granularity experiment
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_listing.py tests/test_gpu_filter.py -m gpu -q -x 2>&1 | tail -4
run() { name=$1; shift
  ( env "$@" CDB_DEBUG_TIMING=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-spans --no-cpu-baseline --no-verify ) > gpurun_out/r02O_$name.json 2> gpurun_out/r02O_$name.err
  grep "document listing\|build:" gpurun_out/r02O_$name.err | head -4
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
j=json.loads([l for l in open(f'gpurun_out/r02O_{n}.json').read().strip().splitlines() if l.startswith('{')][-1])
p=j['roofline']['phases_ms']; s=j['sa_path']['phases_ms']; f=j['filter']
print(n,'value %.4g ms/step %.3f | listing %.3f search %.3f | sa_path: gather %.3f translate %.3f total %.3f | build %.0f ms rebuild %.0f' % (j['value'], j['ms_per_step'], p['listing_ms'], p['search_ms'], s['gather_ms'], s['translate_ms'], s['total_ms'], j['build']['ms'], j['build']['rebuild_ms'] or 0))
print(n,'filter e2e %.4g ms/step %.2f each %s locate %s' % (f['value'], f['ms_per_step'], f['ms_each_step'], f['locate_phases_ms']))
PY
  grep "cdb_filter" gpurun_out/r02O_$name.err | tail -9
}
run base
run l2f32 CDB_L2_FETCH=32
run l2f128 CDB_L2_FETCH=128
