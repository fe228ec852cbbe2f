#!/bin/bash
# First GPU call of the next session: everything that was prepared without a GPU, in the order of its value.
#   gpurun --timeout 2400 -- 'bash tools/next_gpu_session.sh'
# 1. the shipped path is still green (tests + smoke)                      -> gpurun_out/next_gpu_tests.log
# 2. the parked tests of wip/ (decoy, Edits, Genomic Features)            -> gpurun_out/next_wip_*.log
# 3. A/B of the inflate kernel variants (tools/ab_decode.sh)              -> gpurun_out/ab_v*.json
# About 2 + 2 + 6 x 2.5 GPU-minutes.  Nothing here changes the repository; decide from the logs.
set -u
mkdir -p gpurun_out
(timeout 700 python -m pytest tests -m gpu -x -q) > gpurun_out/next_gpu_tests.log 2>&1; tail -2 gpurun_out/next_gpu_tests.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()") > gpurun_out/next_smoke.log 2>&1; tail -1 gpurun_out/next_smoke.log
for t in decoy edits features driver_next; do
  (timeout 300 python -m pytest wip/test_gpu_$t.py -x -q) > gpurun_out/next_wip_$t.log 2>&1
  echo "wip $t: $(tail -1 gpurun_out/next_wip_$t.log)"
done
bash tools/ab_decode.sh 0:0 7:0 15:0 31:0 0:1 31:1
