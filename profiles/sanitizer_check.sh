#!/bin/bash
# compute-sanitizer over the small GPU parity tests (memcheck / racecheck on every kernel at config-1
# size).  Run under gpurun; writes gpurun_out/sanitizer_*.log and prints the summaries.
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_locate_matches_reference tests/test_gpu_parity.py::test_highlight_spans_both_enumerations tests/test_gpu_parity.py::test_chunked_build_equals_single_chunk tests/test_gpu_parity.py::test_large_interval_path tests/test_gpu_parity.py::test_translate_many_doc_ranges"
for tool in memcheck racecheck; do
    timeout 1500 compute-sanitizer --tool $tool --print-limit 3 python -m pytest $T -x -q > gpurun_out/sanitizer_$tool.log 2>&1
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -3
done
