#!/bin/bash
# CTA size A/B with the queue stride fixed: the whole library built with -DB2D_CTA_THREADS=64 (twice the CTAs per
# SM, same warps) against the shipped 128-thread build: GPU test suite on the variant, then the four kernels, two runs each.
mkdir -p gpurun_out
V=$PWD/delaydiffeq.jl_b200/lib/variants/all64.so
B200DDE_LIB=$V timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/all64_pytest.txt 2>&1; echo "pytest(all64) rc=$?"; tail -3 gpurun_out/all64_pytest.txt
for w in bc mg sd stiff; do
  for rep in 1 2; do
    python bench.py --workload $w --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w 128 %.3f ms' % d['ms_per_step'])"
    B200DDE_LIB=$V python bench.py --workload $w --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w 64  %.3f ms' % d['ms_per_step'])"
  done
done
