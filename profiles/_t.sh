#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bt_b1.json 2> gpurun_out/bt_b1.err
CDB_GATHER_BUCKETS=0 python bench.py --no-cpu-baseline --no-rebuild > gpurun_out/bt_b0.json 2> gpurun_out/bt_b0.err
python - <<'P'
import json
for t in ('b1','b0'):
    try:
        d=json.load(open(f'gpurun_out/bt_{t}.json'))
        ph=d['roofline']['phases_ms']
        print(t, round(d['value']/1e6,2),'Mq/s', {k:round(v,3) for k,v in ph.items()}, 'e2e', round(d['e2e']['value']/1e6,2), (d.get('cpu_baseline') or {}).get('parity_with_gpu'))
    except Exception as e: print(t,'failed',e)
P
