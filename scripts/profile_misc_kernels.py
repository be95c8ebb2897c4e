"""Exercises the kernels no single-GPU frame launches, so that one ncu run can capture them (profiles/r2_final_*):
k_assemble (the gather path's scatter of another rank's tiles), k_wait_words / k_set_word (completion flags of a shared
frame) and k_resolve's peer-store form — two renderers on ONE GPU sharing a frame (crtc_share_frame), as
tests/test_z_new_gpu_paths.py does — and the device BVH build + shade_sort = 2 on the bench scene given as argument."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from chameleonrt_b200 import RenderCUDA  # noqa: E402

bench.select_workload(sys.argv[1] if len(sys.argv) > 1 else "c2")
scene, view = bench.make_workload()
# device BVH build + material-sorted shade queue
r = RenderCUDA(0, max_depth=bench.MAX_DEPTH, bvh_builder="device", shade_sort=2)
r.initialize(bench.WIDTH, bench.HEIGHT)
r.set_scene(scene)
for f in range(2):
    r.render(*view, f == 0, False)
del r
# two ranks on one GPU: gather-style assembly, then a shared (peer-written) frame with its flags
ranks = []
for rank in range(2):
    q = RenderCUDA(0, max_depth=bench.MAX_DEPTH, rank=rank, world_size=2)
    q.initialize(bench.WIDTH, bench.HEIGHT)
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
