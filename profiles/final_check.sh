#!/bin/bash
# Round-end check under gpurun: GPU parity suite, locate metrics at cfg3 (tag as $1), the bench line.
TAG=${1:-r01p}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
M=$M,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
M=$M,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum,launch__registers_per_thread,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
timeout 600 ncu --metrics $M --clock-control none -k regex:'search_kernel|gather_kernel|translate_kernel' -s 4 -c 4 --csv \
    --log-file gpurun_out/${TAG}_locate_cfg3_metrics.csv python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_locate_cfg3_metrics.bench.log 2>&1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench.json
