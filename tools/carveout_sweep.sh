for w in bc mg; do for c in default 100 80 65 50; do
  if [ $c = default ]; then unset B2D_SMEM_CARVEOUT; else export B2D_SMEM_CARVEOUT=$c; fi
  python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w carveout $c', round(d['ms_per_step'],3))"
done; done
