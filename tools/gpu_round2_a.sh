#!/bin/bash
# GPU call: first run of the streamed engine — box facts, smoke, the GPU test-suite (per-test timeout), a 30 M-record bench.
set -u
mkdir -p gpurun_out
(nproc; free -g | head -2; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; df -h /tmp | tail -1; lscpu | grep -i "numa\|model name" ) > gpurun_out/r2a_box.log 2>&1
cat gpurun_out/r2a_box.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/r2a_smoke.log 2>&1; tail -2 gpurun_out/r2a_smoke.log
(timeout 1500 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider) > gpurun_out/r2a_gpu_tests.log 2>&1; tail -25 gpurun_out/r2a_gpu_tests.log
(timeout 600 python bench.py --records 30000000 --steps 3 --warmup 3) > gpurun_out/r2a_bench30.json 2> gpurun_out/r2a_bench30.err; tail -3 gpurun_out/r2a_bench30.err; cat gpurun_out/r2a_bench30.json | head -c 3000
