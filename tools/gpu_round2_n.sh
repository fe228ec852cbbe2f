#!/bin/bash
# GPU call on N GPUs: bench at N ranks (timed runs vs the full-size golden, one-file shard parity) + the two-device driver test.
set -u
N=${1:?n gpus}
mkdir -p gpurun_out
free -g | head -2 | tail -1
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3) > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
grep -v "^\[bench\|NCCL\|^$" gpurun_out/r2n_bench_n$N.err | tail -5
python - $N <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r2n_bench_n{sys.argv[1]}.json").read().splitlines()[-1])
    print("N=%s resident %.1f ms (%.0f M rec/s)  e2e %.1f ms (%.0f M rec/s; ceiling %.1f ms, %.1f GB/s per GPU)  reduce %.2f ms  rank ms %s" % (sys.argv[1], d["ms_per_step"], d["value"] / 1e6,
          d["e2e"]["ms_per_step"], d["e2e"]["value"] / 1e6, d["e2e"]["h2d_ceiling_ms"], d["e2e"]["h2d_ceiling_gbs_per_gpu"], d["config"]["ms_reduce"], d["config"]["rank_ms_total"]))
    print(d["parity"]); print(d["merged_parity"]); print(d["sample_parity"])
except Exception as e:
    print("no bench line", e)
PY
if [ "$N" = "2" ]; then
  (timeout 200 python -m pytest tests/test_gpu_host_driver.py -m gpu -q --timeout 180 -p no:cacheprovider -k "two_devices or devices") > gpurun_out/r2n_driver_tests.log 2>&1; tail -3 gpurun_out/r2n_driver_tests.log
fi
