# mg_ring.sh: Mackey-Glass kernel time against an exact (non power-of-two) ring capacity (tools/build_full_variant.sh ... -DB2D_RING_CAP_FIXED=n)
run() { python bench.py --workload mg --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/mgring_$1.json 2> gpurun_out/mgring_$1.err; python -c "
import json; d=json.load(open('gpurun_out/mgring_$1.json')); c=d['config']; print('$1: %.2f ms  cap %s success %s accepted %s traffic %s' % (d['ms_per_step'], c.get('ring_capacity'), c.get('success'), c.get('accepted_steps_per_step'), d['roofline'].get('traffic')))"; }
run base
for c in 12 13 14; do
  export B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/mgring$c.so
  run cap$c
done
