"""First-light check on a GPU box: kernel-level parity, frame parity, stage times."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chameleonrt_b200 import ArcballCamera, RenderCUDA
from chameleonrt_b200.scenes import cornell_box, sponza_like
from oracle import OracleBackend
from oracle.oracle import primary_rays

def run(name, scene, cam, w, h, frames, depth):
    camera = ArcballCamera(cam["eye"], cam["center"], cam["up"])
    gpu = RenderCUDA(0, max_depth=depth); cpu = OracleBackend(max_depth=depth)
    for r in (gpu, cpu):
        r.initialize(w, h); t = time.time(); r.set_scene(scene); print(name, r.name(), 'set_scene s', round(time.time()-t, 3))
    print(gpu.scene_info())
    rays = primary_rays(w, h, camera.eye(), camera.dir(), camera.up(), cam["fov_y"])
    hg = gpu.trace_closest(rays); hc = cpu.trace_closest(rays)
    same = (hg.view(np.uint32) == hc.view(np.uint32)).all(axis=1)
    print(name, 'primary hits bit-identical:', same.mean(), 'hit frac', (hc[:,3].view(np.uint32) != 0xffffffff).mean())
    for f in range(frames):
        sg = gpu.render(camera.eye(), camera.dir(), camera.up(), cam["fov_y"], f == 0, True)
        sc = cpu.render(camera.eye(), camera.dir(), camera.up(), cam["fov_y"], f == 0, True)
        a, b = gpu.read_accum(), cpu.read_accum()
        d = np.abs(a - b); tol = 1e-4 + 1e-3*np.abs(b)
        print(name, f, 'gpu rays', sg.num_rays, 'cpu rays', sc.num_rays, 'gpu ms', round(sg.render_time, 3), 'cpu ms', round(sc.render_time, 1),
              'within tol', (d <= tol).all(axis=2).mean(), 'bit-equal px', (a.view(np.uint32) == b.view(np.uint32)).all(axis=2).mean(),
              'max abs', d.max(), 'rel_l1', d.sum()/np.abs(b).sum(), 'MRays/s gpu', round(sg.rays_per_second/1e6, 1))
        print('  stages', {k: round(v, 3) for k, v in gpu.stage_times().items()}, gpu.counters())
    return gpu

scene, cam = cornell_box(spp=1)
run('cornell', scene, cam, 512, 512, 3, 5)
scene, cam = sponza_like(spp=1, tex_size=256)
run('sponza_small', scene, cam, 320, 180, 2, 5)
scene, cam = sponza_like(spp=4)
camera = ArcballCamera(cam["eye"], cam["center"], cam["up"])
for depth in (5, 8):
    gpu = RenderCUDA(0, max_depth=depth)
    gpu.initialize(1280, 720); gpu.set_scene(scene)
    for f in range(6):
        sg = gpu.render(camera.eye(), camera.dir(), camera.up(), cam["fov_y"], f == 0, True)
    print('C2 depth', depth, 'ms', sg.render_time, 'rays', sg.num_rays, 'MRays/s', sg.rays_per_second/1e6)
    print('  stages', {k: round(v, 3) for k, v in gpu.stage_times().items()}, gpu.counters())
from PIL import Image as PI
os.makedirs('gpurun_out', exist_ok=True)
PI.fromarray(gpu.img.view(np.uint8).reshape(720, 1280, 4)[..., :3]).save('gpurun_out/c2_gpu.png')
