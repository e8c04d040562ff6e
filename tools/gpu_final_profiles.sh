#!/bin/bash
# gpu_final_profiles.sh: the evidence of the committed kernels — plain default bench line, launch list of the same command,
# one `--set full --import-source on` capture per headline kernel (each only after its plain run exited 0)
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err || { tail -5 gpurun_out/r02_bench_default.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches.log 2>&1
echo "launch list rc=$?"; grep -c mos_kernel gpurun_out/r02_launches.csv
for w in bc mg sd stiff; do bash tools/ncu_full.sh $w final; done
