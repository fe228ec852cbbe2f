#!/bin/bash
# GPU call: decode tail spreading — inflate parity tests, then configs[0] and configs[1] bench lines.
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -x -q 2>&1 | tail -4)
for shape in c1 wgs; do
  (timeout 900 python bench.py --shape $shape --steps 3) > gpurun_out/r2i_bench_$shape.json 2> gpurun_out/r2i_bench_$shape.err
  python - $shape <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2i_bench_{sys.argv[1]}.json").read().splitlines()[-1])
    print(sys.argv[1], "resident %.1f ms  e2e %.1f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, "|", d["parity"], "|", d["roofline"]["frac"])
except Exception as e:
    print("no bench line", e)
PY
done
