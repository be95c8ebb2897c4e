"""How well does one GPU do on 1/N of the frame (what each rank sees at N GPUs), with F frames per wavefront?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import make_workload, WIDTH, HEIGHT, MAX_DEPTH
from chameleonrt_b200 import RenderCUDA
scene, view = make_workload()
base = None
for world, F in [(1, 1), (1, 2), (1, 4), (2, 1), (2, 2), (4, 1), (4, 4), (8, 1), (8, 2), (8, 4), (8, 8)]:
    gpu = RenderCUDA(0, max_depth=MAX_DEPTH, rank=0, world_size=world)
    gpu.initialize(WIDTH, HEIGHT); gpu.set_scene(scene)
    gpu.render_async(*view, True, F); gpu.render_async(*view, False, F); gpu.sync()
    t0 = time.perf_counter()
    for _ in range(24 // F):
        gpu.render_async(*view, False, F)
    tot, stages, cnt, n = gpu.sync()
    wall = (time.perf_counter() - t0) / n * 1e3
    base = base or wall
    print(f"world={world} frames_in_flight={F} ms_per_frame={wall:.3f} ideal={base/world:.3f} eff={base/world/wall:.2f} frames={n}", flush=True)
