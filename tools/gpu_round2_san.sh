#!/bin/bash
# GPU call: compute-sanitizer over the round-2 kernel set (streamed waves on two kernel streams, tiled quality pass, -n kernels).
set -u
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
(timeout 300 $S --tool memcheck python __graft_entry__.py smoke) > gpurun_out/r2_san_memcheck_smoke.log 2>&1; grep -E "ERROR SUMMARY|smoke" gpurun_out/r2_san_memcheck_smoke.log | tail -2
(timeout 300 $S --tool racecheck python __graft_entry__.py smoke) > gpurun_out/r2_san_racecheck_smoke.log 2>&1; grep -E "RACECHECK SUMMARY|smoke" gpurun_out/r2_san_racecheck_smoke.log | tail -2
# long reads (tile pass, cooperative CIGAR walk), several waves
(timeout 300 $S --tool memcheck python tools/prof_run.py 3000 6 1 2) > gpurun_out/r2_san_memcheck_long.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r2_san_memcheck_long.log | tail -1
(timeout 300 $S --tool racecheck python tools/prof_run.py 1500 6 1 2) > gpurun_out/r2_san_racecheck_long.log 2>&1; grep -E "RACECHECK SUMMARY" gpurun_out/r2_san_racecheck_long.log | tail -1
# the edge-case and multi-wave tests (tiny waves: carry, -n, ring recycling, quality table edges) under memcheck
(timeout 600 $S --tool memcheck --target-processes all python -m pytest tests/test_gpu_edge.py tests/test_gpu_stream.py -m gpu -x -q -k "not ring and not progress") > gpurun_out/r2_san_memcheck_tests.log 2>&1; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_san_memcheck_tests.log | tail -3
