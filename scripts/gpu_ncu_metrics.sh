#!/bin/bash
# Light ncu pass (a handful of metrics, no source) over the launches of one frame (frame 4) for each option set:
#   bash scripts/gpu_ncu_metrics.sh tag kernel-regex "-" "trav_kernel=1" ...
set -u
out=gpurun_out/r2
mkdir -p "$out"
tag=$1; shift
rx=$1; shift
wl=${WORKLOAD:-c2}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__inst_executed_op_shared_atom.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
i=0
for opts in "$@"; do
    [ "$opts" = "-" ] && opts=""
    CRT_CUDA_OPTIONS="$opts" timeout 600 ncu --metrics $M --clock-control none -k regex:$rx --launch-skip ${SKIP:-27} --launch-count ${COUNT:-9} --csv \
        --log-file "$out/metrics_${tag}_$i.csv" python bench.py --workload $wl --steps 1 --warmup 3 --profile-mode > "$out/metrics_${tag}_$i.log" 2>&1
    echo "== $opts"; python scripts/ncu_metrics_table.py "$out/metrics_${tag}_$i.csv"
    i=$((i+1))
done
