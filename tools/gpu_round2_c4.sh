#!/bin/bash
# GPU call: configs[3] (long reads, 2 M records of 10-50 kb) at its named size on one GPU; the long-read tests first.
set -u
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_edge.py tests/test_gpu_stream.py -m gpu -x -q 2>&1 | tail -4) | tee gpurun_out/r2_c4_tests.log
grep -q "passed" gpurun_out/r2_c4_tests.log && ! grep -q "failed" gpurun_out/r2_c4_tests.log || { echo "tests failed: no bench"; exit 1; }
free -g | head -2 | tail -1
(timeout 1300 python bench.py --shape c4 --steps 2 --warmup 3) > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "rc=$?"
grep -E "generated|warm-up|resident steps|e2e steps|Error|error" gpurun_out/r2_bench_c4.err | tail -8
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_c4.json").read().splitlines()[-1])
    print("c4 resident %.1f ms (%.2f M rec/s, %.1f GB/s inflated)  e2e %.1f ms" % (d["ms_per_step"], d["value"] / 1e6, d["decompressed_gbs"], d["e2e"]["ms_per_step"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}); print(d["parity"]); print(d["sample_parity"])
except Exception as e:
    print("no bench line", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv,noheader
