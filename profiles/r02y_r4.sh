#!/bin/bash
# A/B: distribution sort for intervals of 65..128 occurrences too (R = 4) instead of the sorting network — shard-sized workload
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 600 python bench.py --workload cfg2 --npat 1000000 --steps 10 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify --no-filter 2>gpurun_out/r02y_$name.err | tail -1 > gpurun_out/r02y_$name.json
  python - "$name" <<'PY'
import sys,json
n=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/r02y_{n}.json")); p=j['roofline']['phases_ms']
    print(n, "ms/step %.3f search %.3f gather %.3f translate %.3f total %.3f" % (j['ms_per_step'],p['search_ms'],p['gather_ms'],p['translate_ms'],p['total_ms']))
except Exception as e: print(n,"failed",e)
PY
}
{
run base
run r4 CDB_LIB=$PWD/coffeedb_b200/libcoffeedb_b200_r4.so
run base_rb24 CDB_RANGE_BITS=24
CDB_LIB=$PWD/coffeedb_b200/libcoffeedb_b200_r4.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -2
} > gpurun_out/r02y_ab.txt 2>&1
cat gpurun_out/r02y_ab.txt
