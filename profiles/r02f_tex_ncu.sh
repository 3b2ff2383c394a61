#!/bin/bash
mkdir -p gpurun_out
for v in tex notex; do
  if [ $v = notex ]; then export CDB_IDS_TEX=0; else unset CDB_IDS_TEX; fi
  timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section MemoryWorkloadAnalysis_Chart \
    --section WarpStateStats --section SchedulerStats --section Occupancy --clock-control none \
    -k regex:'translate_kernel' -s 3 -c 1 -f -o gpurun_out/r02f_translate_$v \
    python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-rebuild > gpurun_out/r02f_$v.bench.log 2>&1
done
ls -la gpurun_out | tail -5
