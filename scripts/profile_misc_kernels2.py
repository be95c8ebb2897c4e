"""The kernels scripts/profile_misc_kernels.py did not reach within its time limit: the LBVH builder's two kernels and the
multi-GPU assembly / completion-flag kernels, on a small scene (Cornell box) so that an ncu pass takes seconds."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chameleonrt_b200 import ArcballCamera, RenderCUDA  # noqa: E402
from chameleonrt_b200.scenes import sponza_like  # noqa: E402

scene, cam = sponza_like(spp=2, detail=0.5, tex_size=64)
c = ArcballCamera(cam["eye"], cam["center"], cam["up"])
view = (c.eye(), c.dir(), c.up(), cam["fov_y"])
r = RenderCUDA(0, bvh_builder="device_lbvh")
r.initialize(640, 360)
r.set_scene(scene)
r.render(*view, True, False)
del r
ranks = []
for rank in range(2):
    q = RenderCUDA(0, rank=rank, world_size=2)
    q.initialize(640, 360)
    q.set_scene(scene)
    ranks.append(q)
for q in ranks:
    q.render(*view, True, False)
for src, q in enumerate(ranks):
    acc, img, _ = q.local_buffers()
    ranks[0].assemble_rank(src, 2, acc, img)
ranks[0].read_img()
ranks[0].share_frame_with(ranks[1])
for f in range(3):
    for q in ranks:
        q.render_async(*view, f == 0, 1)
    for q in ranks:
        q.sync()
    ranks[0].read_img()
print("ok")
