#!/bin/bash
# compute-sanitizer over the document-listing kernels (listing_build_kernel, listing_emit_kernel incl. the run-length path, the tagged
# directory, lazy rows in cdb_filter: filter_direct_batch_kernel, mark_need_kernel, emit_listed_rows) at small sizes
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_listing.py tests/test_gpu_filter.py::test_filter_equals_reference_server tests/test_gpu_filter.py::test_filter_many_single_keyword_requests_match_locate_rows -x -q > gpurun_out/r02S_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02S_memcheck.log | tail -3
timeout 1500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest "tests/test_gpu_listing.py::test_listed_rows_equal_reference_rows" "tests/test_gpu_listing.py::test_id_order_listing_through_filter" -x -q > gpurun_out/r02S_racecheck.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r02S_racecheck.log | tail -3
timeout 900 compute-sanitizer --tool initcheck --print-limit 5 python -m pytest "tests/test_gpu_listing.py::test_listed_rows_equal_reference_rows" -x -q > gpurun_out/r02S_initcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r02S_initcheck.log | tail -3
