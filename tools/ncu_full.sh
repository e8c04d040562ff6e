#!/bin/bash
# ncu_full.sh WORKLOAD TAG: one `--set full --import-source on` capture of the full-size kernel (after a plain run exited 0)
W=$1; T=$2
python bench.py --workload $W --secondary none --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/full_${W}_$T.plain.json 2> gpurun_out/full_${W}_$T.plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:mos_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_${W}_$T \
  python bench.py --workload $W --secondary none --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/full_${W}_$T.log 2>&1
echo "ncu full $W $T rc=$?"
