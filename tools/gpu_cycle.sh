#!/bin/bash
# One GPU cycle for a kernel change, meant to be run under gpurun from the repository root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_cycle.sh <tag> [tests] [bench] [ncu]'
# tests : pytest -m gpu                     -> gpurun_out/<tag>_gpu_tests.log
# bench : python bench.py (100 M records)   -> gpurun_out/<tag>_bench.json / .err, one summary line on stdout
# ncu   : launch list + --set full capture of decode / resolve / facets / crc on a 12 M-record input
#         -> gpurun_out/<tag>_launches.csv, <tag>.ncu-rep   (summarise here with tools/ncu_report.py,
#            tools/ncu_lines.py, tools/ncu_stalls.py)
# About 1.5 + 3 + 2.5 GPU-minutes.
set -u
tag=${1:?tag}
shift
what=${*:-tests bench}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests)
      (timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/${tag}_gpu_tests.log 2>&1
      tail -3 gpurun_out/${tag}_gpu_tests.log ;;
    bench)
      timeout 800 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      python - "$tag" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().splitlines()[-1])
print("resident %.1f M rec/s (%.1f ms)  e2e %.1f M rec/s (%.1f ms)" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["ms_per_step"]))
print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, d["parity"])
PY
      ;;
    ncu)
      timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
        python tools/prof_run.py 12000000 1 2 > gpurun_out/${tag}_launches.log 2>&1
      timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:inflate_|facets|crc32' -s 5 -c 4 -o gpurun_out/${tag} \
        python tools/prof_run.py 12000000 1 2 > gpurun_out/${tag}_ncu.log 2>&1
      tail -1 gpurun_out/${tag}_ncu.log ;;
  esac
done
