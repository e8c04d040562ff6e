#!/bin/bash
# gpu_multi.sh N: the multi-GPU evidence on an N-GPU box — GPU tests (incl. the G-GPU == 1-GPU bitwise test and the library
# gather), then the bench line under torchrun with N ranks (library NCCL gather timed, shard parity on every rank)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_${N}_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q -k "multi or gather or sharding" > gpurun_out/multi_${N}_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/multi_${N}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
  > gpurun_out/multi_${N}_bench.json 2> gpurun_out/multi_${N}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/multi_${N}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/multi_${N}_bench.json"))
s = d.get("secondary") or {}
print("bc  value %.3e  e2e %.3e  gather_ms %s ok %s  parity %s" % (d["value"], d["e2e"]["value"], d.get("gather_ms"), d.get("gather_ok"), d["shard_parity"]))
print("    host_io", d["e2e"]["host_io"]["achieved_gbs"], "of", d["e2e"]["host_io"]["ceiling_gbs"])
if s: print("mg  value %.3e  e2e %.3e  gather_ms %s ok %s  parity %s" % (s["value"], s["e2e"]["value"], s.get("gather_ms"), s.get("gather_ok"), s["shard_parity"]))
PY
