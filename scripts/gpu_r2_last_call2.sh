#!/bin/bash
set -u
out=gpurun_out/r2last
mkdir -p "$out"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
timeout 70 ncu --metrics $M --clock-control none -k regex:'k_lbvh_hierarchy|k_lbvh_refit|k_assemble|k_wait_words|k_set_word|k_resolve' -c 60 --csv \
    --log-file "$out/misc2.csv" python scripts/profile_misc_kernels2.py > "$out/misc2.log" 2>&1; echo "misc2 rc=$?"
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
