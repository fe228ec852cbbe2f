#!/bin/bash
# GPU call: scan/facets on their own stream beside the next wave's inflate — multi-wave tests, then the configs[1] bench line
# with the per-wave timeline.
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_decoy.py -m gpu -x -q 2>&1 | tail -4)
(NGSQ_TRACE=1 timeout 900 python bench.py --steps 3) > gpurun_out/r2j_bench_wgs.json 2> gpurun_out/r2j_bench_wgs.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2j_bench_wgs.json").read().splitlines()[-1])
    print("wgs resident %.1f ms  e2e %.1f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, "|", d["parity"], "|", d["roofline"]["frac"])
except Exception as e:
    print("no bench line", e)
PY
grep -E "^\[ngsq trace\]|wave " gpurun_out/r2j_bench_wgs.err | tail -24
