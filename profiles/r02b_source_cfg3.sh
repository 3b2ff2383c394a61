#!/bin/bash
# source-level counters (per-instruction stalls) of the locate kernels AT cfg3 — a handful of replay passes only
mkdir -p gpurun_out
timeout 1200 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis \
    --section LaunchStats --section Occupancy --section SpeedOfLight --clock-control none --import-source on \
    -k regex:'gather_kernel|translate_kernel' -s 6 -c 2 -f -o gpurun_out/r02b_locate_cfg3 \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-rebuild > gpurun_out/r02b_locate_cfg3.bench.log 2>&1
tail -3 gpurun_out/r02b_locate_cfg3.bench.log | cut -c1-400
ls -la gpurun_out/
