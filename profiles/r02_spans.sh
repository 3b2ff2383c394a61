#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "spans or highlight" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_dropin_server.py -q -x -m gpu 2>&1 | tail -2
timeout 900 python bench.py --workload cfg3 --steps 3 --no-cpu-baseline --no-rebuild --no-extras --no-verify --no-filter > gpurun_out/r02_spans_bench.json 2> gpurun_out/r02_spans_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02_spans_bench.json').read().strip().splitlines()[-1])
print(json.dumps(j['spans'])[:700])
PY
