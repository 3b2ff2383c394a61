python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_c_$tag.json 2>gpurun_out/bench_c_$tag.err; python - <<P
import json
try:
    j=json.loads(open("gpurun_out/bench_c_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(j["value"]/1e6,2),"Mq/s", {k:round(v,3) for k,v in j["roofline"]["phases_ms"].items()}, "e2e", round(j["e2e"]["value"]/1e6,2), "build_ms", round(j["build"]["ms"]))
except Exception as e: print("$tag", "ERR", e)
P
}
run m0 CDB_TR_MODE=0
run m1 CDB_TR_MODE=1
run m2 CDB_TR_MODE=2
run m5 CDB_TR_MODE=5
run m6 CDB_TR_MODE=6
run m0p64 CDB_TR_MODE=0 CDB_L2_PERSIST_MB=64
run m1rb21 CDB_TR_MODE=1 CDB_RANGE_BITS=21
run m5rb21 CDB_TR_MODE=5 CDB_RANGE_BITS=21
run m1rb20 CDB_TR_MODE=1 CDB_RANGE_BITS=20
run m1noptab CDB_TR_MODE=1 CDB_PTAB_BITS=0
