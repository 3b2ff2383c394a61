#!/bin/bash
# cdb_filter in parts (copies of part k under the kernels of part k+1) now that the device phase is 2.5 ms of a 12 ms call
mkdir -p gpurun_out
for parts in 1 2 4; do
( CDB_FILTER_PARTS=$parts timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-spans --no-cpu-baseline --no-verify --no-rebuild ) 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); f=j['filter']
print('parts $parts e2e %.4g ms/step %.2f each %s' % (f['value'], f['ms_per_step'], f['ms_each_step']))"
done
