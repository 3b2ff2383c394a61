#!/bin/bash
# Profile of the listing path (recipe of B200_PROFILING.md) + a grid-size A/B of listing_emit_kernel, run under gpurun:
#   r02N_launches_cfg3.csv          launch list of the bench command at the 10 GB configuration
#   r02zz_locate_cfg3_metrics.csv   DRAM bytes, request port, L2 hit rate, issue / occupancy / stall metrics of the kernels of a
#                                   step AT cfg3 (one launch = 10^6 keywords): listing path, then the suffix-array path
#   r02N_listing_cfg2.ncu-rep       --set full capture with source (1 GB corpus, 10^6 keywords)
TAG=r02N
mkdir -p gpurun_out
COMMON="--warmup 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify --no-filter"
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 600 python bench.py --workload $wl --npat 1000000 --steps 10 $COMMON 2>gpurun_out/${TAG}_${name}_$wl.err | tail -1 > gpurun_out/${TAG}_${name}_$wl.json
  python - "$name" "$wl" <<'PY'
import sys,json
n,wl=sys.argv[1:3]
try:
    j=json.load(open(f"gpurun_out/r02N_{n}_{wl}.json")); p=j['roofline']['phases_ms']
    print(n, wl, "ms/step %.3f search %.3f gather(scan etc) %.3f listing %.3f total %.3f frac %.3f" % (j['ms_per_step'],p['search_ms'],p['gather_ms'],p['listing_ms'],p['total_ms'], j['roofline']['frac']))
except Exception as e: print(n,wl,"failed",e)
PY
}
{
for wl in cfg3 cfg2; do
for c in 1 2 3 4 8; do run ctas$c $wl CDB_EMIT_CTAS=$c; done
done
} > gpurun_out/${TAG}_ab.txt 2>&1
cat gpurun_out/${TAG}_ab.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_registers
M=$M,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct
M=$M,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv \
    python bench.py --workload cfg3 $COMMON --steps 2 > gpurun_out/${TAG}_launches_cfg3.bench.log 2>&1
# 3 warm-up + 1 timed step on the listing path (search, listing_rowlen, scans, stats32, listing_emit), then the same on the
# suffix-array path (search, gather, translate): -s skips the build's launches of the kernels named in the filter (none)
timeout 900 ncu --metrics $M --clock-control none -k regex:'search_kernel|gather_kernel|translate_kernel|listing_emit_kernel|listing_rowlen_kernel|stats32_kernel|listing_build_kernel' -c 40 --csv \
    --log-file gpurun_out/r02zz_locate_cfg3_metrics.csv python bench.py --workload cfg3 $COMMON --steps 1 \
    > gpurun_out/r02zz_locate_cfg3_metrics.bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'search_kernel|listing_emit_kernel' -s 6 -c 2 \
    -f -o gpurun_out/${TAG}_listing_cfg2 python bench.py --workload cfg2 --npat 1000000 --steps 1 $COMMON \
    > gpurun_out/${TAG}_listing_cfg2.bench.log 2>&1
ls -la gpurun_out/${TAG}_* gpurun_out/r02zz_*
