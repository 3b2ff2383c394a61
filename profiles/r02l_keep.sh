#!/bin/bash
# translate: priority of the ids[] slice in L2 (evict_last keeps the slices of EARLIER doc ranges alive too) + the full bench line with the filter leg
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload cfg3 --steps 5 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-filter --no-verify 2>gpurun_out/r02l_$name.err | tail -1 > gpurun_out/r02l_$name.json
  python - "$name" <<'PY'
import sys,json
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r02l_{n}.json")); p=j['roofline']['phases_ms']
    print(n, "ms/step %.3f search %.3f gather %.3f translate %.3f total %.3f" % (j['ms_per_step'],p['search_ms'],p['gather_ms'],p['translate_ms'],p['total_ms']))
except Exception as e: print(n,"failed",e)
PY
}
{
run last CDB_TRANSLATE_KEEP=0
run normal CDB_TRANSLATE_KEEP=1
run unchanged CDB_TRANSLATE_KEEP=2
run normal_rb23 CDB_TRANSLATE_KEEP=1 CDB_RANGE_BITS=23
run normal_rb21 CDB_TRANSLATE_KEEP=1 CDB_RANGE_BITS=21
} > gpurun_out/r02l_keep.txt 2>&1
timeout 900 python bench.py --steps 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
cat gpurun_out/r02l_keep.txt; tail -c 5000 gpurun_out/r02l_bench.json; tail -3 gpurun_out/r02l_bench.err
