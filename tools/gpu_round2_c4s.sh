#!/bin/bash
# GPU call: long-read paths after a kernel change — the long-read tests, then configs[3]'s shape at a quarter of its size.
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_edge.py tests/test_gpu_stream.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -15)
(timeout 600 python bench.py --shape c4 --records 500000 --steps 2 --warmup 3 --cpu-sample 20000) > gpurun_out/r2_bench_c4s.json 2> gpurun_out/r2_bench_c4s.err; echo "rc=$?"
grep -E "generated|warm-up|resident steps|e2e steps|Error|error" gpurun_out/r2_bench_c4s.err | tail -8
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c4s.json").read().splitlines()[-1])
    print("c4/4 resident %.1f ms (%.2f M rec/s, %.1f GB/s inflated)  e2e %.1f ms" % (d["ms_per_step"], d["value"] / 1e6, d["decompressed_gbs"], d["e2e"]["ms_per_step"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}); print(d["parity"]); print(d["sample_parity"])
except Exception as e:
    print("no bench line", e)
PY
