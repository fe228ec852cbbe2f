#!/bin/bash
# GPU call: final validation of the shipped build — every GPU test, smoke, the default bench line, and the ncu launch list of
# the same workload (one timed step, resident only).
set -u
mkdir -p gpurun_out
cp ngs_b200/libngs_cuda.so gpurun_out/r2_final_libngs_cuda.so
(timeout 1500 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider) > gpurun_out/r2_final_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_final_gpu_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()") 2>&1 | tail -1
(timeout 900 python bench.py) > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_final_bench.json").read().splitlines()[-1])
    print("resident %.1f ms (%.1f M rec/s)  e2e %.1f ms (%.1f M rec/s)" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, "|", d["parity"], "|", d["roofline"], "|", d["cpu_baseline"], "|", d["clocks"], "| launches", d["gpu_launches"])
except Exception as e:
    print("no bench line", e)
PY
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e) > gpurun_out/r2_final_ncu.log 2>&1; tail -2 gpurun_out/r2_final_ncu.log
python tools/launch_table.py gpurun_out/r2_final_launches.csv | tail -30
