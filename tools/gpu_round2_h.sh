#!/bin/bash
# GPU call: validate the current build (tests, full-size bench with golden parity), then ncu of facets.
set -u
mkdir -p gpurun_out
cp ngs_b200/libngs_cuda.so gpurun_out/r2h_libngs_cuda.so
(timeout 1500 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider -x) > gpurun_out/r2h_gpu_tests.log 2>&1; tail -3 gpurun_out/r2h_gpu_tests.log
(timeout 900 python bench.py --steps 3) > gpurun_out/r2h_bench100.json 2> gpurun_out/r2h_bench100.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2h_bench100.json").read().splitlines()[-1])
    print("resident %.1f ms  e2e %.1f ms (device total %.1f, tail %.1f, h2d ceiling %.1f ms)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["device_ms_total"], d["e2e"]["ms_tail_after_last_wave_starts"], d["e2e"]["h2d_ceiling_ms"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, "|", d["parity"], "|", d["roofline"]["frac"])
except Exception as e:
    print("no bench line", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:facets' -s 1 -c 1 -o gpurun_out/r2h_facets python tools/prof_run.py 12000000 1 2 > gpurun_out/r2h_ncu.log 2>&1; tail -1 gpurun_out/r2h_ncu.log
