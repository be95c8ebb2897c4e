#!/bin/bash
# Round 2, last GPU call (the --set full captures of every kernel ran into the call's limit and were lost: 40 GPU-minutes;
# see profiles/r2_experiments.md). A LIGHT ncu pass instead — ~25 metrics per launch, a few replays each — over
#   (1) the 27 launches of one C2 frame, (2) the k_traverse launches of one C4 frame (DRAM traffic on an HBM-resident scene),
#   (3) the kernels no plain frame launches (device BVH build, shade_sort, k_assemble, flags), each step under its own timeout;
# first of all the bench line of the final build.
set -u
out=gpurun_out/r2last
mkdir -p "$out"
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
timeout 170 python bench.py --steps 20 --warmup 3 > "$out/bench_n1_final.json" 2> "$out/bench_n1_final.err"; echo "bench rc=$?"; tail -c 400 "$out/bench_n1_final.json"
timeout 100 ncu --metrics $M --clock-control none --launch-skip 81 --launch-count 27 --csv --log-file "$out/frame_c2.csv" \
    python bench.py --steps 1 --warmup 3 --profile-mode > "$out/frame_c2.log" 2>&1; echo "c2 rc=$?"
timeout 120 ncu --metrics $M --clock-control none -k regex:k_traverse --launch-skip 27 --launch-count 9 --csv --log-file "$out/trav_c4.csv" \
    python bench.py --workload c4 --steps 1 --warmup 3 --profile-mode > "$out/trav_c4.log" 2>&1; echo "c4 rc=$?"
timeout 150 ncu --metrics $M --clock-control none -k regex:'k_flatten|k_lbvh|k_radix|k_scan|k_bvh2|k_ploc|k_plan|k_emit|k_pack|k_queue|k_assemble|k_wait|k_set_word|k_resolve' \
    -c 600 --csv --log-file "$out/misc_c2.csv" python scripts/profile_misc_kernels.py c2 > "$out/misc_c2.log" 2>&1; echo "misc rc=$?"
ls -la "$out"
