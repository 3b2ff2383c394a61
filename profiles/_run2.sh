python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_b_default.json 2> gpurun_out/bench_b_default.err
CDB_RANGE_BITS=21 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_b_rb21.json 2>&1
CDB_RANGE_BITS=23 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_b_rb23.json 2>&1
CDB_RANGE_BITS=30 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_b_rb30.json 2>&1
CDB_L2_FETCH=32 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_b_l2f32.json 2>&1
for f in default rb21 rb23 rb30 l2f32; do python - <<P
import json
try:
    j=json.loads(open("gpurun_out/bench_b_$f.json").read().strip().splitlines()[-1])
    print("$f", round(j["value"]/1e6,2),"Mq/s", {k:round(v,3) for k,v in j["roofline"]["phases_ms"].items()}, "e2e", round(j["e2e"]["value"]/1e6,2))
except Exception as e: print("$f", "ERR", e)
P
done
bash profiles/run_profile.sh r01c > gpurun_out/profile.log 2>&1; tail -3 gpurun_out/profile.log
