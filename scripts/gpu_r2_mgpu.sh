#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus N): the multi-GPU tests, frame assembly by gather vs by peer stores, the bench line.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 1200 -- 'bash scripts/gpu_r2_mgpu.sh'
set -u
out=gpurun_out/r2
mkdir -p "$out"
n=$(python -c 'import torch; print(torch.cuda.device_count())')
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 "$@"; }
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > "$out/pytest_mgpu_n$n.log" 2>&1; tail -3 "$out/pytest_mgpu_n$n.log"
CRT_BENCH_FRAME=gather run bench.py --gpus "$n" --steps 20 --warmup 3 > "$out/bench_n${n}_gather.json" 2> "$out/bench_n${n}_gather.err"
run bench.py --gpus "$n" --steps 20 --warmup 3 > "$out/bench_n${n}_peer.json" 2> "$out/bench_n${n}_peer.err"
for k in gather peer; do python - "$out/bench_n${n}_$k.json" $k <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
          "fbf", round(d["frame_by_frame"]["value"], 1), "fbf ms", round(d["frame_by_frame"]["ms_per_step"], 3), "bit_identical", d.get("mgpu_bit_identical"),
          "disp", {k: round(v, 3) for k, v in d["dispersion"]["ms_per_step"].items()})
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
tail -3 "$out"/bench_n${n}_*.err
