python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --steps 3 --rebuild > gpurun_out/bench_l_$tag.json 2>gpurun_out/bench_l_$tag.err; python - <<P
import json
try:
    j=json.loads(open("gpurun_out/bench_l_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(j["value"]/1e6,2),"Mq/s", {k:round(v,3) for k,v in j["roofline"]["phases_ms"].items()}, "e2e", round(j["e2e"]["value"]/1e6,2), "build", {k:(round(v,1) if isinstance(v,float) else v) for k,v in j["build"].items() if k in ("ms","sort_ms","rounds","chunks","rebuild_ms")})
except Exception as e: print("$tag", "ERR", e); print(open("gpurun_out/bench_l_$tag.err").read()[-1500:])
P
}
run cfg3
run cfg2 CDB_BENCH_WORKLOAD=cfg2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01i_launches_build_cfg3.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r01i.log 2>&1
