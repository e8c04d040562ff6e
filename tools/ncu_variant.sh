#!/bin/bash
# ncu_variant.sh WORKLOAD NAME...: warp-state / scheduler / instruction sections of the full-size kernel for library variants
W=$1; shift
for n in "$@"; do
  B200DDE_LIB=$PWD/delaydiffeq.jl_b200/lib/variants/$n.so ncu --clock-control none -k regex:mos_kernel --launch-skip 3 -c 1 \
    --section WarpStateStats --section SchedulerStats --section InstructionStats --section SpeedOfLight --section LaunchStats --section Occupancy \
    --metrics smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,sm__icc_requests.sum,sm__icc_requests_lookup_hit.sum,sm__icc_requests_lookup_miss.sum \
    --csv --page raw --log-file gpurun_out/ncu_${W}_$n.csv python bench.py --workload $W --steps 1 --warmup 4 --no-e2e --no-cpu > gpurun_out/ncu_${W}_$n.log 2>&1
  echo "$n rc=$?"
done
