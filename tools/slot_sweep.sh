# slot_sweep.sh: end-to-end time of b2d_solve against the number of waves in flight (B2D_SLOTS) and the wave size
for w in mg bc; do
  for cfg in "2 0" "3 0" "4 0" "3 1" "4 1"; do
    set -- $cfg
    export B2D_SLOTS=$1
    if [ $2 = 0 ]; then unset B2D_CHUNK_GRIDS; else export B2D_CHUNK_GRIDS=$2; fi
    python bench.py --workload $w --steps 4 --warmup 3 --no-cpu > gpurun_out/slots_${w}_$1_$2.json 2> gpurun_out/slots_${w}_$1_$2.err
    python - <<PY
import json
d = json.load(open("gpurun_out/slots_${w}_$1_$2.json")); print("$w slots=$1 grids=$2 kernel %.2f ms  e2e %.2f ms  match %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("matches_device_path")))
PY
  done
done
