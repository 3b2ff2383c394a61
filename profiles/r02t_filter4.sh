#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg3 --steps 8 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify > gpurun_out/r02t_bench.json 2> gpurun_out/r02t_bench.err
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02t_bench.json').read().strip().splitlines()[-1])
print('e2e', j['e2e']['value'], 'filter ms', j['filter']['ms_per_step'], j['filter']['ms_each_step'], j['filter']['locate_phases_ms'])
print('full rows', j['e2e_full_rows'])
PY
nvidia-smi topo -m 2>/dev/null | head -8; numactl -H 2>/dev/null | head -5; nproc
