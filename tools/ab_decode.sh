#!/bin/bash
# A/B of the decode kernel's symbol-loop variants (inflate_lane.cuh, NGSQ_DEC_VARIANT) on one GPU box:
#   gpurun --timeout 1700 -- 'bash tools/ab_decode.sh 0 7 15 31'
# For every variant: rebuild libngs_cuda.so in place, run the inflate / facet parity tests, then a 30 M-record
# resident-only bench (stage_ms carries the decode kernel's time).  Results: gpurun_out/ab_v<variant>.{json,log}.
# The default library (variant 0) is rebuilt at the end.  About 1 GPU-minute of build + 1.5 of run per variant.
set -u
mkdir -p gpurun_out
for v in "$@"; do
  make -C ngs_b200/csrc -B VARIANT=$v ../../ngs_b200/libngs_cuda.so > gpurun_out/ab_v${v}.log 2>&1 || { echo "variant $v: build failed"; continue; }
  (timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q) >> gpurun_out/ab_v${v}.log 2>&1
  tests=$(tail -1 gpurun_out/ab_v${v}.log)
  timeout 400 python bench.py --records 30000000 --no-e2e --no-cpu > gpurun_out/ab_v${v}.json 2>> gpurun_out/ab_v${v}.log
  python - "$v" "$tests" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/ab_v{sys.argv[1]}.json").read().splitlines()[-1])
    print("variant %s: %.1f ms/step, %s, launch_ms %s | %s" % (sys.argv[1], d["ms_per_step"], {k: round(x, 1) for k, x in d["config"]["stage_ms"].items()},
          d["roofline"].get("launch_ms"), sys.argv[2]))
except Exception as e:
    print("variant %s: no bench line (%s) | %s" % (sys.argv[1], e, sys.argv[2]))
PY
done
make -C ngs_b200/csrc -B VARIANT=0 ../../ngs_b200/libngs_cuda.so > /dev/null 2>&1
