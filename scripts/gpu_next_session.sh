#!/bin/bash
# What to run first on the next GPU box, in one gpurun call (1 GPU, ~12 minutes):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/gpu_next_session.sh'
# Everything it exercises was built after round 1's GPU budget was spent and is so far verified on the CPU only
# (profiles/README.md, "Not yet measured on a GPU"). Outputs land in gpurun_out/next/.
# A second call with --gpus 2 (or more) runs the multi-GPU part:  bash scripts/gpu_next_session.sh mgpu
set -u
out=gpurun_out/next
mkdir -p "$out"
if [ "${1:-}" = "mgpu" ]; then
    n=$(python -c 'import torch; print(torch.cuda.device_count())')
    run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 "$@"; }
    run scripts/mgpu_check.py 1280 720 > "$out/mgpu_gather.log" 2>&1
    run scripts/mgpu_check.py 1280 720 --peer > "$out/mgpu_peer.log" 2>&1
    run bench.py --gpus "$n" --steps 20 --warmup 3 > "$out/bench_n${n}_gather.json" 2> "$out/bench_n${n}_gather.err"
    CRT_BENCH_FRAME=peer run bench.py --gpus "$n" --steps 20 --warmup 3 > "$out/bench_n${n}_peer.json" 2> "$out/bench_n${n}_peer.err"
    grep -h "MGPU" "$out"/mgpu_*.log; tail -c 600 "$out"/bench_n*_*.json
    exit 0
fi
# 1. the GPU suite; the newest tests are in the last file
timeout 900 python -m pytest tests -m gpu -x -q > "$out/pytest_gpu.log" 2>&1; tail -5 "$out/pytest_gpu.log"
# 2. the bench line (its "set_scene" object compares the host and the device BVH build, in a child process)
timeout 600 python bench.py --steps 20 --warmup 3 > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -c 1500 "$out/bench_n1.json"
# 3. set_scene with every builder on C2 and C4 (phase times, tree sizes, traversal cost, frames bit-identical?)
timeout 600 python scripts/set_scene_timing.py c2 c4 > "$out/set_scene_timing.log" 2>&1; cat "$out/set_scene_timing.log"
# 4. the traversal variants (refill threshold, shadow-ray order, deferred triangle pass, device-built tree)
timeout 600 python scripts/tune_traversal.py c2 > "$out/tune_traversal_c2.log" 2>&1; cat "$out/tune_traversal_c2.log"
# 5. where a device build spends its time: launch list of one set_scene (cold, serialised: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/device_build_launches.csv" \
    python scripts/set_scene_timing.py c2 --only device --no-render > "$out/ncu_device_build.log" 2>&1
python - <<'PY'
import csv, collections, sys
rows = [r for r in csv.reader(open("gpurun_out/next/device_build_launches.csv", errors="replace")) if len(r) > 10 and r[0].isdigit()]
acc = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = r[4].split("(")[0]; acc[name] += float(r[-1].replace(",", "")); cnt[name] += 1
tot = sum(acc.values()) or 1
for k, v in acc.most_common(16):
    print(f"{k:40s} {cnt[k]:5d} launches {v / 1e3:10.1f} us {100 * v / tot:5.1f} %")
PY
