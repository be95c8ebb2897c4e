"""Sweep the traversal scheduling knobs on the C2 workload (run on the GPU box)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_workload, WIDTH, HEIGHT, MAX_DEPTH
from chameleonrt_b200 import RenderCUDA
scene, view = make_workload()
for refill in (1, 2, 4, 8, 12):
    gpu = RenderCUDA(0, max_depth=MAX_DEPTH)
    gpu._check(gpu.lib.crtc_set_option(gpu.h, b"refill_idle", refill))
    gpu.initialize(WIDTH, HEIGHT); gpu.set_scene(scene)
    acc = {}
    for f in range(8):
        st = gpu.render(*view, f == 0, False)
        if f >= 3:
            for k, v in gpu.stage_times().items(): acc[k] = acc.get(k, 0) + v / 5
    print(f"refill_idle={refill:2d} " + ' '.join(f"{k}={v:.3f}" for k, v in acc.items()), flush=True)
