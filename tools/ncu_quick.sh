#!/bin/bash
# ncu_quick.sh WORKLOAD TAG: stall / scheduler / icache counters of the full-size kernel (one launch), raw csv into gpurun_out/
W=$1; T=$2
ncu --clock-control none -k regex:mos_kernel --launch-skip 3 -c 1 \
  --section WarpStateStats --section SchedulerStats --section SpeedOfLight --section LaunchStats --section Occupancy \
  --metrics smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_drain_per_issue_active.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
  --csv --page raw --log-file gpurun_out/ncuq_${W}_$T.csv python bench.py --workload $W --secondary none --steps 1 --warmup 4 --no-e2e --no-cpu > gpurun_out/ncuq_${W}_$T.log 2>&1
echo "ncu $W $T rc=$?"
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/ncuq_${W}_$T.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h, u, v = rows[hdr], rows[hdr + 1], rows[hdr + 2]
for k, uu, x in zip(h, u, v):
    if any(t in k for t in ("stalled", "duration", "icc", "fp64", "issue_active", "occupancy_limit", "inst_executed", "warps_active", "hit_rate", "registers_per", "shared_mem_per")):
        try:
            if "stalled" in k and float(x.replace(",", "")) < 0.05: continue
        except ValueError: pass
        print(f"  {k} [{uu}] = {x}")
PY
