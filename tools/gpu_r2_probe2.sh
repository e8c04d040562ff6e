#!/bin/bash
# second round-2 probe batch: stiff ring capacity vs live window; CTA size re-run with the queue stride fixed
mkdir -p gpurun_out
python tools/stiff_cap_probe.py > gpurun_out/stiff_cap_probe.txt 2>&1
bash tools/run_variants.sh bc bc64x8 > gpurun_out/cta_variants.txt 2>&1
bash tools/run_variants.sh mg mg64x10 >> gpurun_out/cta_variants.txt 2>&1
python bench.py --workload bc --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('bc default %.3f ms parity %s' % (d['ms_per_step'], d.get('parity')))" >> gpurun_out/cta_variants.txt
python bench.py --workload mg --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('mg default %.3f ms' % d['ms_per_step'])" >> gpurun_out/cta_variants.txt
cat gpurun_out/stiff_cap_probe.txt gpurun_out/cta_variants.txt
