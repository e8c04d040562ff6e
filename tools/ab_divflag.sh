for w in bc mg sd stiff; do
  for rep in 1 2; do
    python bench.py --workload $w --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w flag   %.3f ms' % d['ms_per_step'])"
    B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/${w}_nodf.so python bench.py --workload $w --secondary none --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$w noflag %.3f ms' % d['ms_per_step'])"
  done
done
