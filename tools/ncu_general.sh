#!/bin/bash
# one --set full capture of the general saveat kernel on the bc_model sweep (B2D_FORCE_GENERAL: no option in use)
export B2D_FORCE_GENERAL=1
python bench.py --workload bc --secondary none --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/gen3_bc.plain.json 2> gpurun_out/gen3_bc.plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:mos_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_r2_bc_general3 \
  python bench.py --workload bc --secondary none --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/gen3_bc.log 2>&1
echo "ncu rc=$?"
