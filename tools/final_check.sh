set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_final_bc.json 2> gpurun_out/bench_final_bc.err; echo "bench rc=$?"
for w in mg sd stiff; do python bench.py --workload $w > gpurun_out/bench_final_$w.json 2> gpurun_out/bench_final_$w.err; echo "bench $w rc=$?"; done
python - <<'PY'
import json
for w in ("ref","bc","mg","sd","stiff"):
    d=json.load(open(f"gpurun_out/bench_final_{w}.json")); print(w, "%.4g" % d.get("value"), "%.3f ms" % d.get("ms_per_step"), "e2e %.4g" % d.get("e2e",{}).get("value"), "%.2f ms" % d.get("e2e",{}).get("ms_per_step", 0), d.get("roofline",{}).get("frac"))
PY
