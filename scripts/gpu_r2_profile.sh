#!/bin/bash
# Round 2, GPU call: ncu --set full captures WITH source-level stall sampling (the build has -lineinfo) of the kernels that
# decide the frame, on C2 (L2-resident scene) and C4 (HBM-resident scene). One GPU.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_r2_profile.sh'
# Outputs: gpurun_out/r2/*.ncu-rep (read here with `ncu -i ... --page raw|source --csv`, summarised into profiles/r2_*).
set -u
out=gpurun_out/r2
mkdir -p "$out"
tag=${1:-a}
# frame 4 of a profile-mode run = launches 81.. (27 per frame at depth 8): raygen, traverse(primary), shade(0),
# traverse(shadow 0 + bounce 1), nee_resolve(0), shade(1), traverse(1)
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 81 --launch-count 7 -f -o "$out/c2_frame_$tag" \
    python bench.py --steps 1 --warmup 3 --profile-mode > "$out/ncu_c2_$tag.log" 2>&1
tail -2 "$out/ncu_c2_$tag.log"
timeout 900 ncu --set full --import-source on --clock-control none --launch-skip 81 --launch-count 4 -f -o "$out/c4_frame_$tag" \
    python bench.py --workload c4 --steps 1 --warmup 3 --profile-mode > "$out/ncu_c4_$tag.log" 2>&1
tail -2 "$out/ncu_c4_$tag.log"
ls -la "$out"
