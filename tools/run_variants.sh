#!/bin/bash
# run_variants.sh WORKLOAD [names...]: kernel time of bench.py's device-resident leg for each library variant (GPU box)
W=${1:-bc}; shift
NAMES=${@:-$(ls delaydiffeq.jl_b200/lib/variants/*.so | xargs -n1 basename | sed 's/\.so$//')}
for n in $NAMES; do
  B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/$n.so python bench.py --workload $W --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/var_${W}_$n.json 2> gpurun_out/var_${W}_$n.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/var_${W}_$n.json")); print("$W $n: %.3f ms  %.3e steps/s" % (d["ms_per_step"], d["value"]))
except Exception as e:
    print("$W $n: failed", e)
PY
done
