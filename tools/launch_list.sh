# launch_list.sh: ncu launch list of the default bench command (run only after the plain command exited 0)
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_plain.json 2> gpurun_out/launch_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launch_ncu.log 2>&1
echo "ncu rc=$?"; grep -c mos_kernel gpurun_out/launches_r1_final.csv
