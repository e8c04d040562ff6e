# mg_refill.sh: lane refill (-DB2D_REFILL_OVERRIDE=n) for Mackey-Glass: parity suite with the variant library, then kernel time
export B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/mgrefill8.so
python -m pytest tests -m gpu -q -x > gpurun_out/refill_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/refill_tests.log
unset B200DDE_LIB
run() { python bench.py --workload mg --steps 4 --warmup 3 --no-cpu > gpurun_out/mgrefill_$1.json 2> gpurun_out/mgrefill_$1.err; python -c "
import json; d=json.load(open('gpurun_out/mgrefill_$1.json')); c=d['config']; print('$1: %.2f ms  e2e %.2f ms match %s success %s accepted %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e'].get('matches_device_path'), c.get('success'), c.get('accepted_steps_per_step')))"; }
run base
for r in 4 8 16; do
  export B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/mgrefill$r.so
  run refill$r
done
