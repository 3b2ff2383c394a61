#!/bin/bash
# translate A/B at cfg3: first design vs sub-warp-group design, occupancy, group width, slice size
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02c_tests.txt
run() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --workload cfg3 --steps 5 --no-cpu-baseline --no-rebuild 2>gpurun_out/r02c_$name.err | tail -1 > gpurun_out/r02c_$name.json
  python - "$name" <<'PY'
import sys,json
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r02c_{n}.json")); p=j['roofline']['phases_ms']
    print(n, "ms/step %.3f search %.3f gather %.3f translate %.3f total %.3f e2e %.3g" % (j['ms_per_step'],p['search_ms'],p['gather_ms'],p['translate_ms'],p['total_ms'],j['e2e']['value']))
except Exception as e: print(n,"failed",e)
PY
}
{
run old CDB_TRANSLATE=1
run g8o6 CDB_TRANSLATE_G=8
run g8o5 CDB_TRANSLATE_G=8 CDB_TRANSLATE_OCC=5
run g4o6 CDB_TRANSLATE_G=4
run g16o6 CDB_TRANSLATE_G=16
run g8rb21 CDB_TRANSLATE_G=8 CDB_RANGE_BITS=21
run g4rb21 CDB_TRANSLATE_G=4 CDB_RANGE_BITS=21
run g4rb20 CDB_TRANSLATE_G=4 CDB_RANGE_BITS=20
run g8rb23 CDB_TRANSLATE_G=8 CDB_RANGE_BITS=23
run g16rb23 CDB_TRANSLATE_G=16 CDB_RANGE_BITS=23
} > gpurun_out/r02c_ab.txt 2>&1
cat gpurun_out/r02c_tests.txt gpurun_out/r02c_ab.txt
