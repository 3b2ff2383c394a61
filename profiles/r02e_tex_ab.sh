#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02e_tests.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --workload cfg3 --steps 5 --no-cpu-baseline --no-rebuild 2>gpurun_out/r02e_$name.err | tail -1 > gpurun_out/r02e_$name.json
  python - "$name" <<'PY'
import sys,json
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r02e_{n}.json")); p=j['roofline']['phases_ms']
    print(n, "ms/step %.3f search %.3f gather %.3f translate %.3f total %.3f e2e %.3g 1q %.1f us conc %s" % (j['ms_per_step'],p['search_ms'],p['gather_ms'],p['translate_ms'],p['total_ms'],j['e2e']['value'],j['single_query_latency_us'],j['concurrent_single_queries']))
except Exception as e: print(n,"failed",e)
PY
}
{
run tex
run notex CDB_IDS_TEX=0
run texrb21 CDB_RANGE_BITS=21
run texrb23 CDB_RANGE_BITS=23
run texrb20 CDB_RANGE_BITS=20
} > gpurun_out/r02e_ab.txt 2>&1
cat gpurun_out/r02e_tests.txt gpurun_out/r02e_ab.txt
