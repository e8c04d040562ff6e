#!/bin/bash
# first GPU call of round 2: GPU tests (incl. the new full-size Mackey-Glass and the many-discontinuities tests), smoke,
# the default bench line (bc_model + Mackey-Glass secondary block), the FP64 latency table
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/r2_bench_first.json 2> gpurun_out/r2_bench_first.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench_first.err
timeout 120 tools/bin/fp64_latency > gpurun_out/r2_fp64_latency.txt 2>&1; echo "lat rc=$?"
cat gpurun_out/r2_fp64_latency.txt
