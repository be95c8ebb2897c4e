"""How well does one GPU do on 1/N of the frame (what each rank sees at N GPUs)?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload, WIDTH, HEIGHT, MAX_DEPTH
from chameleonrt_b200 import RenderCUDA
scene, view = make_workload()
base = None
for world, lanes in [(1, 1), (1, 2), (2, 1), (2, 2), (4, 1), (4, 2), (4, 3), (8, 1), (8, 2), (8, 3), (8, 4)]:
    gpu = RenderCUDA(0, max_depth=MAX_DEPTH, rank=0, world_size=world)
    gpu._check(gpu.lib.crtc_set_option(gpu.h, b"lanes", lanes))
    gpu.initialize(WIDTH, HEIGHT); gpu.set_scene(scene)
    acc = {}; wall = 0.0
    for f in range(13):
        t0 = time.perf_counter()
        st = gpu.render(*view, f == 0, False)
        dt = time.perf_counter() - t0
        if f >= 3:
            wall += dt / 10
            for k, v in gpu.stage_times().items(): acc[k] = acc.get(k, 0) + v / 10
    base = base or acc['frame']
    print(f"world={world} lanes={lanes} frame_ms={acc['frame']:.3f} wall_ms={wall*1e3:.3f} ideal={base/world:.3f} eff={base/world/ (wall*1e3):.2f}", flush=True)
