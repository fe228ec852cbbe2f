#!/bin/bash
# GPU call: configs[4] (RNA-seq, 200 M spliced 2x100 bp) and configs[3] (2 M long reads) at their named sizes, timed runs
# checked against the full-size goldens.
set -u
mkdir -p gpurun_out
for shape in ${@:-c5 c4}; do
  (timeout 1300 python bench.py --shape $shape --steps 2 --warmup 3) > gpurun_out/r2_bench_$shape.json 2> gpurun_out/r2_bench_$shape.err; echo "$shape rc=$?"
  grep -E "generated|resident steps|serial-stages|Error|error" gpurun_out/r2_bench_$shape.err | tail -5
  python - $shape <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2_bench_{sys.argv[1]}.json").read().splitlines()[-1])
    print(sys.argv[1], "resident %.1f ms (%.2f M rec/s, %.1f GB/s inflated)  e2e %.1f ms (%.2f M rec/s)" % (d["ms_per_step"], d["value"] / 1e6, d["decompressed_gbs"], d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}); print("serial", d["config"]["serial_stages"]); print(d["parity"]); print(d["sample_parity"]); print(d["cpu_baseline"])
except Exception as e:
    print("no bench line", e)
PY
done
