#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the reference arm and our arm of the bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x --durations=6 2>&1 | tail -14 > gpurun_out/r02Z_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02Z_smoke.txt 2>&1
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02Z_bench_ref.json 2> gpurun_out/r02Z_bench_ref.err
( time timeout 1200 python bench.py ) > gpurun_out/r02Z_bench.json 2> gpurun_out/r02Z_bench.err
cat gpurun_out/r02Z_tests.txt; tail -2 gpurun_out/r02Z_smoke.txt | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/r02Z_bench_ref.json','gpurun_out/r02Z_bench.json'):
    try:
        j=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith('{')][-1])
        print(f, 'value %.4g' % j['value'], 'e2e %.4g' % j['e2e']['value'], 'ms/step', j.get('ms_per_step'))
        if 'roofline' in j: print('  roofline', j['roofline']['frac'], j['roofline']['path']['frac'], 'launches', j['gpu_launches'], 'clocks', j['clocks'])
        if j.get('extras'): print('  extras keys', list(j['extras'].keys()), json.dumps(j['extras'].get('query_pool_cpp'))[:600])
        if j.get('filter_span'): print('  ref filter_span', j['filter_span']['value'])
    except Exception as e: print(f, 'failed', e)
PY
grep real gpurun_out/r02Z_bench_ref.err gpurun_out/r02Z_bench.err
