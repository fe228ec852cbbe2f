#!/bin/bash
set -u
mkdir -p gpurun_out
cp ngs_b200/libngs_cuda.so gpurun_out/r2g_libngs_cuda.so
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:crc32|facets' -s 2 -c 2 -o gpurun_out/r2g_step python tools/prof_run.py 12000000 1 2 > gpurun_out/r2g_ncu_step.log 2>&1; tail -1 gpurun_out/r2g_ncu_step.log
