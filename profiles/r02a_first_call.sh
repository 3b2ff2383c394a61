#!/bin/bash
# round 2, first GPU call: the experimental paths that never ran in round 1 + a source-level capture of the round-1 kernels
mkdir -p gpurun_out
{
echo "== small batch check"; timeout 300 python tools/small_batch_check.py
echo "== experimental tests"; CDB_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -k "small_batch or gather_mid" -x -q 2>&1 | tail -5
echo "== pool bench general"; timeout 300 tools/_build/query_pool_bench 100000 100 4 1000
echo "== pool bench small"; CDB_SMALL_BATCH=256 timeout 300 tools/_build/query_pool_bench 100000 100 4 1000
echo "== bench cfg3 L2 persist 48"; CDB_L2_PERSIST_MB=48 timeout 600 python bench.py --workload cfg3 --steps 5 --no-cpu-baseline --no-rebuild 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['phases_ms'])"
echo "== bench cfg3 base"; timeout 600 python bench.py --workload cfg3 --steps 5 --no-cpu-baseline --no-rebuild 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['roofline']['phases_ms'], j['single_query_latency_us'], j['concurrent_single_queries'])"
} > gpurun_out/r02a_log.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|translate_kernel' -s 6 -c 2 \
    -f -o gpurun_out/r02a_locate_cfg2 python bench.py --workload cfg2 --steps 1 --warmup 3 --no-cpu-baseline --no-rebuild \
    > gpurun_out/r02a_locate_cfg2.bench.log 2>&1
tail -40 gpurun_out/r02a_log.txt
