run() { python bench.py --workload mg --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/mgring_$1.json 2> gpurun_out/mgring_$1.err; python -c "
import json; d=json.load(open('gpurun_out/mgring_$1.json')); c=d['config']; print('$1: %.2f ms  cap %s success %s' % (d['ms_per_step'], c.get('ring_capacity'), c.get('success')))"; }
export B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/mgr12np5.so
run r12np_capped
B2D_NO_L2_CAP=1 run r12np_5ctas
export B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/mgr12np6.so
B2D_NO_L2_CAP=1 run r12np_6ctas
