"""Tiny render for compute-sanitizer (memcheck / racecheck on the compaction + traversal kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chameleonrt_b200 import ArcballCamera, RenderCUDA
from chameleonrt_b200.scenes import cornell_box, sponza_like
for make, (w, h) in ((lambda: cornell_box(spp=2), (96, 64)), (lambda: sponza_like(spp=1, detail=0.15, tex_size=32), (64, 48))):
    scene, cam = make()
    c = ArcballCamera(cam["eye"], cam["center"], cam["up"])
    r = RenderCUDA(0, max_depth=5)
    r.initialize(w, h); r.set_scene(scene)
    for f in range(2):
        st = r.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0, True)
    a = r.read_accum()
    print("SANITIZE_RUN_OK", st.num_rays, float(a.mean()))
