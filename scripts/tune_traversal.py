"""Sweep the traversal scheduling knobs on the C2 workload (run on the GPU box)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import make_workload, WIDTH, HEIGHT, MAX_DEPTH
from chameleonrt_b200 import RenderCUDA
scene, view = make_workload()
for variant, tri_lanes, refill in [(0, 8, 4), (1, 1, 4), (2, 1, 4), (2, 1, 2), (2, 1, 8), (2, 1, 1), (2, 1, 12)]:
    gpu = RenderCUDA(0, max_depth=MAX_DEPTH)
    gpu._check(gpu.lib.crtc_set_option(gpu.h, b"trav_variant", variant))
    gpu._check(gpu.lib.crtc_set_option(gpu.h, b"tri_lanes", tri_lanes))
    gpu._check(gpu.lib.crtc_set_option(gpu.h, b"refill_idle", refill))
    gpu.initialize(WIDTH, HEIGHT); gpu.set_scene(scene)
    acc = {}
    for f in range(8):
        st = gpu.render(*view, f == 0, False)
        if f >= 3:
            for k, v in gpu.stage_times().items(): acc[k] = acc.get(k, 0) + v / 5
    print(f"variant={variant} tri_lanes={tri_lanes:2d} refill_idle={refill:2d} frame={acc['frame']:.3f} closest={acc['traverse_closest']:.3f} any={acc['traverse_any']:.3f} shade={acc['shade']:.3f}", flush=True)
