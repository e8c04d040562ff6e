set -x
python -m pytest tests -m gpu -x -q > gpurun_out/final_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/final_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
