#!/bin/bash
# ncu --set full + source of ONE traversal launch (the bounce-0 merged launch of frame 4) for each option set
#   bash scripts/gpu_r2_profile_trav.sh tag "-" "trav_kernel=1"
set -u
out=gpurun_out/r2
mkdir -p "$out"
tag=$1; shift
i=0
for opts in "$@"; do
    [ "$opts" = "-" ] && opts=""
    CRT_CUDA_OPTIONS="$opts" timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_traverse --launch-skip ${SKIP:-28} --launch-count 1 \
        -f -o "$out/trav_${tag}_$i" python bench.py --workload ${WORKLOAD:-c2} --steps 1 --warmup 3 --profile-mode > "$out/ncu_trav_${tag}_$i.log" 2>&1
    tail -1 "$out/ncu_trav_${tag}_$i.log"
    i=$((i+1))
done
ls -la "$out" | tail -5
