#!/bin/bash
# A/B of the inflate kernels' compile-time variants (inflate_lane.cuh NGSQ_DEC_VARIANT, inflate2.cuh NGSQ_RES_VARIANT)
# on one GPU box; arguments are <decode mask>:<resolve mask> pairs:
#   gpurun --timeout 1700 -- 'bash tools/ab_decode.sh 0:0 7:0 15:0 31:0 0:1 31:1'
# For every pair: rebuild libngs_cuda.so in place, run the inflate / facet parity tests, then a 30 M-record
# resident-only bench (stage_ms carries the kernels' times).  Results: gpurun_out/ab_v<dec>_<res>.{json,log}.
# The default library (Makefile defaults) is rebuilt at the end.  About 1 GPU-minute of build + 1.5 of run per pair.
set -u
mkdir -p gpurun_out
for pair in "$@"; do
  v=${pair%%:*}; r=${pair##*:}; tag=ab_v${v}_${r}
  make -C ngs_b200/csrc -B cuda VARIANT=$v RES_VARIANT=$r > gpurun_out/$tag.log 2>&1 || { echo "$pair: build failed"; continue; }
  (timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q) >> gpurun_out/$tag.log 2>&1
  tests=$(tail -1 gpurun_out/$tag.log)
  timeout 400 python bench.py --records 30000000 --no-e2e --no-cpu > gpurun_out/$tag.json 2>> gpurun_out/$tag.log
  python - "$pair" "$tag" "$tests" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[2]}.json").read().splitlines()[-1])
    print("variant %s: %.1f ms/step, %s, launch_ms %s | %s" % (sys.argv[1], d["ms_per_step"], {k: round(x, 1) for k, x in d["config"]["stage_ms"].items()},
          d["roofline"].get("launch_ms"), sys.argv[3]))
except Exception as e:
    print("variant %s: no bench line (%s) | %s" % (sys.argv[1], e, sys.argv[3]))
PY
done
make -C ngs_b200/csrc -B cuda > /dev/null 2>&1
