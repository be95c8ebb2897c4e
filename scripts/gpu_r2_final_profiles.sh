#!/bin/bash
# Round 2, final ncu captures (1 GPU): every kernel of the backend under `ncu --set full --clock-control none`.
#   frame:  the 27 launches of one C2 frame and of one C4 frame (k_raygen, k_traverse, k_shade, k_nee_resolve, k_resolve)
#   misc:   one device set_scene (17 build kernels), a shade_sort = 2 frame (k_queue_hist / k_queue_scatter / scans),
#           k_assemble, k_wait_words, k_set_word, k_resolve writing a shared frame — on C2 and, for the build, C4
# The .ncu-rep files stay in gpurun_out/ (scratch); what is committed are the raw-page CSVs reduced to the tables under
# profiles/r2_final_* (scripts/summarize_ncu_table.py) and profiles/traffic_k_traverse.json (scripts/update_traffic.py).
set -u
out=gpurun_out/r2final
mkdir -p "$out"
for wl in c2 c4; do
    timeout 900 ncu --set full --clock-control none --launch-skip 81 --launch-count 27 -f -o "$out/frame_$wl" \
        python bench.py --workload $wl --steps 1 --warmup 3 --profile-mode > "$out/ncu_frame_$wl.log" 2>&1
    ncu -i "$out/frame_$wl.ncu-rep" --page raw --csv > "$out/frame_$wl.csv" 2>/dev/null
    tail -1 "$out/ncu_frame_$wl.log"
done
timeout 1200 ncu --set full --clock-control none -k regex:'k_flatten|k_lbvh|k_radix|k_scan|k_bvh2|k_ploc|k_plan|k_emit|k_pack|k_queue|k_assemble|k_wait|k_set_word|k_resolve' \
    -c 700 -f -o "$out/misc_c2" python scripts/profile_misc_kernels.py c2 > "$out/ncu_misc_c2.log" 2>&1
ncu -i "$out/misc_c2.ncu-rep" --page raw --csv > "$out/misc_c2.csv" 2>/dev/null
tail -1 "$out/ncu_misc_c2.log"
timeout 1200 ncu --set full --clock-control none -k regex:'k_flatten|k_lbvh|k_radix|k_scan|k_bvh2|k_ploc|k_plan|k_emit|k_pack' \
    -c 700 -f -o "$out/build_c4" python scripts/set_scene_timing.py c4 --only device --no-render > "$out/ncu_build_c4.log" 2>&1
ncu -i "$out/build_c4.ncu-rep" --page raw --csv > "$out/build_c4.csv" 2>/dev/null
tail -1 "$out/ncu_build_c4.log"
rm -f "$out"/misc_c2.ncu-rep "$out"/build_c4.ncu-rep   # (hundreds of launches: the CSVs carry what the tables need)
find "$out" -name "*.ncu-rep" -size +24M -delete   # (gpurun brings back at most 64 MiB)
ls -la "$out"
