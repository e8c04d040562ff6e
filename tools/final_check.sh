set -x
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_smoke.py > gpurun_out/sanitizer_r1_final.log 2>&1; echo "sanitizer rc=$?"
tail -4 gpurun_out/sanitizer_r1_final.log
python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_final_bc.json 2> gpurun_out/bench_final_bc.err; echo "bench rc=$?"
python bench.py --workload mg > gpurun_out/bench_final_mg.json 2> gpurun_out/bench_final_mg.err; echo "bench mg rc=$?"
python - <<'PY'
import json
for w in ("ref","bc","mg"):
    d=json.load(open(f"gpurun_out/bench_final_{w}.json")); print(w, d.get("value"), d.get("ms_per_step"), d.get("e2e",{}).get("value"), d.get("roofline",{}).get("frac"))
PY
