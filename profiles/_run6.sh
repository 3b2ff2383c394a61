run() { tag=$1; shift; env "$@" python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_f_$tag.json 2>gpurun_out/bench_f_$tag.err; python - <<P
import json
try:
    j=json.loads(open("gpurun_out/bench_f_$tag.json").read().strip().splitlines()[-1])
    print("$tag", round(j["value"]/1e6,2),"Mq/s", {k:round(v,3) for k,v in j["roofline"]["phases_ms"].items()}, "e2e", round(j["e2e"]["value"]/1e6,2), "build", {k:(round(v,1) if isinstance(v,float) else v) for k,v in j["build"].items() if k in ("ms","sort_ms","rounds","chunks","rebuild_ms")})
except Exception as e: print("$tag", "ERR", e); print(open("gpurun_out/bench_f_$tag.err").read()[-1500:])
P
grep "cdb\]" gpurun_out/bench_f_$tag.err | tail -4
}
run c1 CDB_LOCATE_CHUNKS=1 CDB_DEBUG_TIMING=1
run c1pool CDB_LOCATE_CHUNKS=1 CDB_DEBUG_TIMING=1 CDB_BIGBUF_POOL=1
run c4 CDB_LOCATE_CHUNKS=4 CDB_DEBUG_TIMING=1
run c4pool CDB_LOCATE_CHUNKS=4 CDB_BIGBUF_POOL=1
run c8pool CDB_LOCATE_CHUNKS=8 CDB_BIGBUF_POOL=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gather_kernel|translate_kernel|search_kernel' -c 30 --csv --log-file gpurun_out/r01e_launches_locate.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r01e.log 2>&1
grep -c gather gpurun_out/r01e_launches_locate.csv
