#!/bin/bash
# Profiling recipe (B200_PROFILING.md) — run under gpurun from the repo root:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/run_profile.sh r01'
# Writes gpurun_out/<tag>_launches_cfg3.csv (every launch with its device time, the bench command at the 10 GB
# configuration) and gpurun_out/<tag>_locate_cfg2.ncu-rep (--set full capture of the locate kernels at the 1 GB
# configuration: kernel replay has to save/restore device memory, which is impractical with 100 GB resident).
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches_cfg3.csv python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_cfg3.bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'search_kernel|small_path_kernel' -s 6 -c 3 \
    -f -o gpurun_out/${TAG}_locate_cfg2 python bench.py --workload cfg2 --steps 2 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_locate_cfg2.bench.log 2>&1
ls -la gpurun_out/
