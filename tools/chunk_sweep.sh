# chunk_sweep.sh: end-to-end time of b2d_solve against the wave size (B2D_CHUNK_GRIDS = wave in full grids of resident trajectories)
for w in mg bc; do
  for g in 0 1 2 3 4 7; do
    if [ $g = 0 ]; then unset B2D_CHUNK_GRIDS; else export B2D_CHUNK_GRIDS=$g; fi
    python bench.py --workload $w --steps 4 --warmup 3 --no-cpu > gpurun_out/chunk_${w}_$g.json 2> gpurun_out/chunk_${w}_$g.err
    python - <<PY
import json
d = json.load(open("gpurun_out/chunk_${w}_$g.json")); print("$w grids=$g kernel %.2f ms  e2e %.2f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
  done
done
