#!/bin/bash
# Round-2 profile refresh (recipe of B200_PROFILING.md), run under gpurun from the repo root:
#   r02z_launches_cfg3.csv          launch list of the bench command at the 10 GB configuration (locate steps + cdb_filter leg)
#   r02z_locate_cfg3_metrics.csv    DRAM bytes, L1->crossbar request port, L2 hit rate, issue / occupancy / stall metrics of
#                                   the locate kernels AT cfg3 (one launch = 10^6 patterns)
#   r02z_filter_cfg3_metrics.csv    the same for the kernels of the cdb_filter leg
#   r02z_locate_cfg2.ncu-rep        --set full capture with source at the 1 GB configuration
TAG=${1:-r02z}
mkdir -p gpurun_out
COMMON="--workload cfg3 --warmup 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct
M=$M,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,smsp__inst_executed.sum,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem
M=$M,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
M=$M,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct
M=$M,smsp__warp_issue_stalled_wait_per_warp_active.pct,smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct
M=$M,smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_cfg3.csv \
    python bench.py $COMMON --steps 2 > gpurun_out/${TAG}_launches_cfg3.bench.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'search_kernel|gather_kernel|translate_kernel' -s 4 -c 3 --csv \
    --log-file gpurun_out/${TAG}_locate_cfg3_metrics.csv python bench.py $COMMON --steps 1 --no-filter \
    > gpurun_out/${TAG}_locate_cfg3_metrics.bench.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:'filter_direct_kernel|filter_warp_kernel|size_kernel|compact_kernel|sa_rank_kernel|key_pack_kernel' -c 8 --csv \
    --log-file gpurun_out/${TAG}_filter_cfg3_metrics.csv python bench.py $COMMON --steps 1 \
    > gpurun_out/${TAG}_filter_cfg3_metrics.bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'search_kernel|gather_kernel|translate_kernel' -s 4 -c 3 \
    -f -o gpurun_out/${TAG}_locate_cfg2 python bench.py --workload cfg2 --npat 1000000 --steps 1 --warmup 3 --no-cpu-baseline --no-rebuild --no-extras --no-spans --no-verify --no-filter \
    > gpurun_out/${TAG}_locate_cfg2.bench.log 2>&1
ls -la gpurun_out/${TAG}_*
