# mg_occ.sh: Mackey-Glass kernel time against CTAs per SM (shared memory is the occupancy limiter: the cp.async double
# buffer costs 9 rows), with and without the L2-residency rule
run() { B2D_DEBUG=1 python bench.py --workload mg --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/mgocc_$1.json 2> gpurun_out/mgocc_$1.err; python -c "
import json; d=json.load(open('gpurun_out/mgocc_$1.json')); print('$1: %.2f ms' % d['ms_per_step'])"; }
for v in mg5np mg6np; do
  export B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/$v.so
  run ${v}_capped
  B2D_NO_L2_CAP=1 run ${v}_nocap
  B2D_NO_L2_CAP=1 B2D_BLOCKS_PER_SM=5 run ${v}_nocap_5
done
