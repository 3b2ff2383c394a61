#!/bin/bash
# listing_emit_kernel: a CTA per row (8 warps x 128 entries) against a warp per row, at cfg3 and on shard-sized rows.
# Result (r02W_coop_ab.txt): the CTA-per-row variant is 34 % slower at cfg3 (4.16 against 3.11 ms) and 4.6 x slower on short rows — a warp
# then has one 128-entry burst per row and idles until the next row; the variant was removed again (CDB_EMIT_COOP no longer exists).
mkdir -p gpurun_out
CDB_EMIT_COOP=1 timeout 600 python -m pytest tests/test_gpu_listing.py -m gpu -q -x 2>&1 | tail -2
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --npat 1000000 --steps 10 --warmup 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify --no-filter 2>gpurun_out/r02W_${name}_$wl.err | tail -1 > gpurun_out/r02W_${name}_$wl.json
  python - "$name" "$wl" <<'PY'
import sys,json
n,wl=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r02W_{n}_{wl}.json")); p=j['roofline']['phases_ms']
    print(n, wl, "ms/step %.3f search %.3f listing %.3f total %.3f frac %.3f" % (j['ms_per_step'],p['search_ms'],p['listing_ms'],p['total_ms'], j['roofline']['frac']))
except Exception as e: print(n,wl,"failed",e)
PY
}
{
run warp cfg3 CDB_EMIT_COOP=0
run coop cfg3 CDB_EMIT_COOP=1
run coop3 cfg3 CDB_EMIT_COOP=1 CDB_EMIT_CTAS=3
run coop2 cfg3 CDB_EMIT_COOP=1 CDB_EMIT_CTAS=2
run auto cfg3
run warp cfg2 CDB_EMIT_COOP=0
run coop cfg2 CDB_EMIT_COOP=1
run auto cfg2
} > gpurun_out/r02W_ab.txt 2>&1
cat gpurun_out/r02W_ab.txt
