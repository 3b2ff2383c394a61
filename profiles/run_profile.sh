#!/bin/bash
# Profiling recipe (B200_PROFILING.md) — run under gpurun from the repo root:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash profiles/run_profile.sh r01b'
# Writes into gpurun_out/:
#   <tag>_launches_cfg3.csv     every launch with its device time, the bench command at the 10 GB configuration
#   <tag>_locate_cfg3_metrics.csv  selected DRAM / issue / cache metrics of the locate kernels AT cfg3 (few replay
#                               passes; a --set full capture would have to save/restore >100 GB per pass)
#   <tag>_locate_cfg2.ncu-rep   --set full capture (with source) of the locate kernels at the 1 GB configuration
TAG=${1:-r01}
KERNELS=${2:-'search_kernel|gather_kernel|translate_kernel'}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,smsp__inst_executed.sum,lts__t_bytes.sum
M=$M,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,launch__occupancy_limit_registers
M=$M,launch__occupancy_limit_shared_mem,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_cfg3.bench.log 2>&1
timeout 900 ncu --metrics $M --clock-control none -k regex:"$KERNELS" -s 4 -c 4 --csv \
    --log-file gpurun_out/${TAG}_locate_cfg3_metrics.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_locate_cfg3_metrics.bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KERNELS" -s 4 -c 4 \
    -f -o gpurun_out/${TAG}_locate_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_locate_cfg2.bench.log 2>&1
ls -la gpurun_out/
# build kernels (--set full, with source) at the 1 GB configuration, chunked like the 10 GB one (workspace cap)
CDB_BUILD_WORKSPACE_MB=8000 timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'extract_kernel|onesweep_kernel|ties_|hist_kernel' -c 14 \
    -f -o gpurun_out/${TAG}_build_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_build_cfg2.bench.log 2>&1
ls -la gpurun_out/
