#!/bin/bash
# ncu --set full + source of the first two k_shade launches of frame 4 on C2, for each option set given as argument
# (CRT_CUDA_OPTIONS syntax; "-" = defaults).   bash scripts/gpu_r2_profile_shade.sh tag "-" "shade_unrolled=1"
set -u
out=gpurun_out/r2
mkdir -p "$out"
tag=$1; shift
i=0
for opts in "$@"; do
    [ "$opts" = "-" ] && opts=""
    CRT_CUDA_OPTIONS="$opts" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_shade --launch-skip 24 --launch-count 2 \
        -f -o "$out/shade_${tag}_$i" python bench.py --steps 1 --warmup 3 --profile-mode > "$out/ncu_shade_${tag}_$i.log" 2>&1
    tail -1 "$out/ncu_shade_${tag}_$i.log"
    i=$((i+1))
done
ls -la "$out"
