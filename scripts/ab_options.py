"""A/B timing of renderer options on a bench workload (run on the GPU box). Every option set renders the same frames;
the float framebuffer of each set is compared bit for bit with the first one's.
    python scripts/ab_options.py c2 "" "shade_unrolled=1" "tri_pass_defer=16,refill_idle=2"
An empty string is the default configuration. Prints the mean stage times of frames 3..10 (ms)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from chameleonrt_b200 import RenderCUDA  # noqa: E402

bench.select_workload(sys.argv[1])
scene, view = bench.make_workload()
sets = sys.argv[2:] or [""]
print(f"# {bench.WORKLOAD}; mean of frames 3..10, ms", flush=True)
ref = None
base = None
for spec in sets:
    opts = dict(kv.split("=") for kv in spec.split(",") if kv)
    ctor = {}
    if "bvh_builder" in opts:
        ctor["bvh_builder"] = opts.pop("bvh_builder")
    gpu = RenderCUDA(0, max_depth=bench.MAX_DEPTH, **ctor)
    for k, v in opts.items():
        gpu.set_option(k, int(v))
    gpu.initialize(bench.WIDTH, bench.HEIGHT)
    gpu.set_scene(scene)
    acc = {}
    n = 0
    for f in range(11):
        gpu.render(*view, f == 0, False)
        if f >= 3:
            n += 1
            for k, v in gpu.stage_times().items():
                acc[k] = acc.get(k, 0.0) + v
    acc = {k: v / n for k, v in acc.items()}
    a = gpu.read_accum().view(np.uint32)
    same = "ref" if ref is None else ("bit-identical" if np.array_equal(a, ref) else "DIFFERENT")
    if ref is None:
        ref = a
    base = base or acc["frame"]
    print(f"{spec or 'default':48s} frame {acc['frame']:7.3f} ({acc['frame'] / base - 1:+.1%}) " +
          " ".join(f"{k}={v:.3f}" for k, v in acc.items() if k != "frame") + f"  [{same}]", flush=True)
    del gpu
