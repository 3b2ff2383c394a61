#!/bin/bash
# compute-sanitizer over this round's new kernels at small sizes: cdb_filter (direct / warp / CTA / numeric-scan paths, id order with
# and without the rank companion), rank doubling, the persistent short-row gather, the small-batch path, the multi-device index,
# loader staging, batched spans, the verifier.
mkdir -p gpurun_out
T="tests/test_gpu_filter.py::test_filter_errors tests/test_gpu_filter.py::test_numeric_query_is_the_reference_numeric_query tests/test_gpu_filter.py::test_filter_many_single_keyword_requests_match_locate_rows tests/test_gpu_parity.py::test_long_repeats_take_rank_doubling tests/test_gpu_parity.py::test_small_batch_path_matches_general_path tests/test_gpu_parity.py::test_gather_mid_size_variants tests/test_gpu_parity.py::test_highlight_spans_batch_matches_reference tests/test_gpu_parity.py::test_verifier_detects_corruption tests/test_gpu_parity.py::test_loader_staging_uploads_full_chunks_during_add tests/test_gpu_multidevice.py::test_errors_and_empty_cases"
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest $T -x -q > gpurun_out/r02_sanitizer_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_memcheck.log | tail -3
T2="tests/test_gpu_filter.py::test_filter_many_single_keyword_requests_match_locate_rows tests/test_gpu_parity.py::test_small_batch_path_matches_general_path tests/test_gpu_parity.py::test_gather_mid_size_variants"
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest $T2 -x -q > gpurun_out/r02_sanitizer_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02_sanitizer_racecheck.log | tail -3
# the reference-server parity of cdb_filter under memcheck (the HTTP server runs outside the sanitizer)
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "tests/test_gpu_filter.py::test_filter_equals_reference_server" -x -q > gpurun_out/r02_sanitizer_memcheck_filter.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02_sanitizer_memcheck_filter.log | tail -3
