#!/bin/bash
# gpu_quick.sh [pytest -k expr]: GPU parity suite + short bench lines of the four workloads (kernel only)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${1:+-k "$1"} > gpurun_out/q_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/q_gpu_tests.log
for w in bc mg sd stiff; do
  timeout 300 python bench.py --workload $w --secondary none --no-e2e --no-cpu --steps 5 --warmup 3 > gpurun_out/q_bench_$w.json 2> gpurun_out/q_bench_$w.err || tail -5 gpurun_out/q_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/q_bench_$w.json")); print("$w: %.3f ms  %.3e steps/s parity=%s" % (d["ms_per_step"], d["value"], d["shard_parity"]))
except Exception as e:
    print("$w: failed", e)
PY
done
