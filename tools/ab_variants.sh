#!/bin/bash
# A/B of compile-time kernel variants on one GPU box; arguments are "<make variables>" strings, e.g.
#   gpurun --timeout 1500 -- 'bash tools/ab_variants.sh "RES_VARIANT=1" "RES_VARIANT=3"'
# For every argument: rebuild libngs_cuda.so in place, run the inflate / facet parity tests, then a 30 M-record resident-only
# bench (stage_ms carries the kernels' times).  Results: gpurun_out/ab_<tag>.{json,log}.  The default library is rebuilt at the end.
set -u
mkdir -p gpurun_out
for vars in "$@"; do
  tag=ab_$(echo "$vars" | tr ' =' '__')
  make -C ngs_b200/csrc -B cuda $vars > gpurun_out/$tag.log 2>&1 || { echo "$vars: build failed"; continue; }
  (timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_stream.py -m gpu -x -q --timeout 240) >> gpurun_out/$tag.log 2>&1
  tests=$(tail -1 gpurun_out/$tag.log)
  timeout 400 python bench.py --records 30000000 --no-e2e --no-cpu --steps 5 > gpurun_out/$tag.json 2>> gpurun_out/$tag.log
  python - "$vars" "$tag" "$tests" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/{sys.argv[2]}.json").read().splitlines()[-1])
    print("%s: %.1f ms/step, %s | %s" % (sys.argv[1], d["ms_per_step"], {k: round(x, 1) for k, x in d["config"]["stage_ms"].items()}, sys.argv[3]))
except Exception as e:
    print("%s: no bench line (%s) | %s" % (sys.argv[1], e, sys.argv[3]))
PY
done
make -C ngs_b200/csrc -B cuda > /dev/null 2>&1
