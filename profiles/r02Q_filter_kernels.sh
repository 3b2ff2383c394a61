#!/bin/bash
# per-kernel times of the cdb_filter leg at cfg3 (ncu launch list of the filter kernels only) + listing build with no-allocate lookups
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_listing.py -m gpu -q -x 2>&1 | tail -2
( CDB_DEBUG_TIMING=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'filter_|size_kernel|compact_kernel|key_|mark_need|listing_build|listing_emit' --csv --log-file gpurun_out/r02Q_filter_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-extras --no-spans --no-verify --no-cpu-baseline --no-rebuild ) > gpurun_out/r02Q.json 2> gpurun_out/r02Q.err
grep "document listing" gpurun_out/r02Q.err | head -3
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02Q_filter_launches.csv')) if len(r)>10]
hdr=rows[0]; ki,vi=hdr.index("Kernel Name"),hdr.index("Metric Value")
seen=collections.OrderedDict()
for r in rows[1:]:
    seen.setdefault(r[ki][:60],[]).append(r[vi])
for k,v in seen.items(): print(k, v[-3:], len(v))
PY
