#!/bin/bash
# GPU call: e2e timeline after the table-upload fix, streaming tests, configs[0] bench line.
set -u
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_stream.py tests/test_gpu_edge.py tests/test_gpu_parity.py tests/test_gpu_edits.py -m gpu -q --timeout 240 -p no:cacheprovider) > gpurun_out/r2d_tests.log 2>&1; tail -2 gpurun_out/r2d_tests.log
(NGSQ_TRACE=1 timeout 900 python bench.py --steps 2 --no-cpu) > gpurun_out/r2d_bench100.json 2> gpurun_out/r2d_bench100.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2d_bench100.json").read().splitlines()[-1])
    print("resident %.1f ms  e2e %.1f ms (device total %.1f, tail %.1f, h2d ceiling %.1f ms)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["device_ms_total"], d["e2e"]["ms_tail_after_last_wave_starts"], d["e2e"]["h2d_ceiling_ms"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, "|", d["parity"])
    print("e2e", {k: round(v, 1) for k, v in d["e2e"]["stage_ms"].items()})
except Exception as e:
    print("no bench line", e)
PY
grep "ngsq trace" gpurun_out/r2d_bench100.err | tail -19
(timeout 300 python bench.py --shape c1 --steps 5) > gpurun_out/r2d_bench_c1.json 2> gpurun_out/r2d_bench_c1.err; tail -1 gpurun_out/r2d_bench_c1.err; head -c 1500 gpurun_out/r2d_bench_c1.json
