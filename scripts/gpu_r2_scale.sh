#!/bin/bash
# Round 2: the bench line at N GPUs (gpurun --gpus N), and at N = 8 BASELINE configs[4] (C5) once.
set -u
out=gpurun_out/r2
mkdir -p "$out"
n=$(python -c 'import torch; print(torch.cuda.device_count())')
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29511 "$@"; }
run bench.py --gpus "$n" --steps 20 --warmup 3 > "$out/bench_n${n}.json" 2> "$out/bench_n${n}.err"
if [ "$n" = "8" ]; then
    run bench.py --gpus "$n" --workload c5 --steps 8 --warmup 3 --repeats 2 > "$out/bench_c5_n${n}.json" 2> "$out/bench_c5_n${n}.err"
fi
for f in "$out"/bench_n${n}.json "$out"/bench_c5_n${n}.json; do [ -f "$f" ] && python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3),
          "fbf", round(d["frame_by_frame"]["value"], 1), "fbf ms", round(d["frame_by_frame"]["ms_per_step"], 3), "bit_identical", d.get("mgpu_bit_identical"),
          "disp", {k: round(v, 3) for k, v in d["dispersion"]["ms_per_step"].items()}, "stages", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -n 3 "$out"/bench_*n${n}.err
