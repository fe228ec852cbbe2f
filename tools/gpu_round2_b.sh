#!/bin/bash
# GPU call: tests, full-size bench (timed runs checked against the golden), coverage-resolve A/B (cp.async.bulk staging vs direct
# loads), launch list + ncu --set full of every kernel of a step.
set -u
mkdir -p gpurun_out
cp ngs_b200/libngs_cuda.so gpurun_out/r2b_libngs_cuda.so
(timeout 1500 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider) > gpurun_out/r2b_gpu_tests.log 2>&1; tail -4 gpurun_out/r2b_gpu_tests.log
(timeout 900 python bench.py) > gpurun_out/r2b_bench100.json 2> gpurun_out/r2b_bench100.err; tail -2 gpurun_out/r2b_bench100.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2b_bench100.json").read().splitlines()[-1])
    print("resident %.1f ms  e2e %.1f ms (tail %.1f, h2d ceiling %.1f ms)" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["ms_tail_after_last_wave_starts"], d["e2e"]["h2d_ceiling_ms"]))
    print({k: round(v, 1) for k, v in d["config"]["stage_ms"].items()}, "|", d["parity"], "|", d["sample_parity"])
except Exception as e:
    print("no bench line", e)
PY
for v in 0 1; do
  (NGSQ_COV_BULK=$v timeout 400 python bench.py --records 30000000 --no-e2e --no-cpu --steps 5) > gpurun_out/r2b_covbulk$v.json 2> gpurun_out/r2b_covbulk$v.err
  python - $v <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2b_covbulk{sys.argv[1]}.json").read().splitlines()[-1])
    print("NGSQ_COV_BULK=%s: coverage %.2f ms, step %.1f ms" % (sys.argv[1], d["config"]["stage_ms"]["ms_coverage"], d["ms_per_step"]))
except Exception as e:
    print("no line", e)
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches.csv python tools/prof_run.py 12000000 1 2 > gpurun_out/r2b_launches.log 2>&1
(NGSQ_COV_BULK=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_resolve -s 25 -c 3 -o gpurun_out/r2b_cov_direct python tools/prof_run.py 12000000 1 2) > gpurun_out/r2b_ncu_cov0.log 2>&1
(NGSQ_COV_BULK=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cov_resolve -s 25 -c 3 -o gpurun_out/r2b_cov_bulk python tools/prof_run.py 12000000 1 2) > gpurun_out/r2b_ncu_cov1.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k 'regex:inflate_|facets|crc32|walk_kernel|find_first' -s 10 -c 9 -o gpurun_out/r2b_step python tools/prof_run.py 12000000 1 2 > gpurun_out/r2b_ncu_step.log 2>&1
tail -1 gpurun_out/r2b_ncu_step.log
