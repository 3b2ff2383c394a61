#!/bin/bash
# A/B of listing_emit_kernel variants (unroll, minimum CTAs per SM, store flavour) at cfg3 and on shard-sized rows (cfg2 corpus, 10^6 keywords)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_listing.py -m gpu -q -x 2>&1 | tail -3
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --npat 1000000 --steps 10 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify --no-filter 2>gpurun_out/r02M_${name}_$wl.err | tail -1 > gpurun_out/r02M_${name}_$wl.json
  python - "$name" "$wl" <<'PY'
import sys,json
n,wl=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r02M_{n}_{wl}.json")); p=j['roofline']['phases_ms']
    print(n, wl, "ms/step %.3f search %.3f gather(scan etc) %.3f listing %.3f total %.3f frac %.3f" % (j['ms_per_step'],p['search_ms'],p['gather_ms'],p['listing_ms'],p['total_ms'], j['roofline']['frac']))
except Exception as e: print(n,wl,"failed",e)
PY
}
{
for wl in cfg3 cfg2; do
run base $wl
for v in u4b8 u8b1 u8b6 u4b8st1 u4b8st2 u2b8; do
run $v $wl CDB_LIB=$PWD/coffeedb_b200/libcoffeedb_b200_$v.so
done
done
} > gpurun_out/r02M_ab.txt 2>&1
cat gpurun_out/r02M_ab.txt
