#!/bin/bash
# DRAM / pipe / stall metrics of the locate kernels at cfg3 (one launch = 10^6 patterns), a few replay passes only.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash profiles/locate_metrics.sh r01q'
TAG=${1:-r01q}
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,launch__registers_per_thread
M=$M,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
M=$M,smsp__warp_issue_stalled_barrier_per_warp_active.pct,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct
M=$M,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct
M=$M,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct
M=$M,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct
timeout 600 ncu --metrics $M --clock-control none -k regex:'search_kernel|gather_kernel|translate_kernel' -s 4 -c 3 --csv \
    --log-file gpurun_out/${TAG}_locate_cfg3_metrics.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-rebuild \
    > gpurun_out/${TAG}_locate_cfg3_metrics.bench.log 2>&1
tail -2 gpurun_out/${TAG}_locate_cfg3_metrics.bench.log | cut -c1-300
