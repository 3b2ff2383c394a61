#!/bin/bash
# ncu launch list (device time of every launch) of the bench command at cfg3:
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash profiles/launch_list.sh r01q'
TAG=${1:-r01q}
mkdir -p gpurun_out
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline --no-rebuild \
    > gpurun_out/${TAG}_launches_cfg3.bench.log 2>&1
wc -l gpurun_out/${TAG}_launches_cfg3.csv
