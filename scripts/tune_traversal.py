"""Sweep the scheduling variants of k_traverse on a bench workload (run on the GPU box): the refill threshold, the
shadow-ray order, the deferred triangle pass, and — through the tree they traverse — the BVH builder. Every variant
renders the same image (tested); only the stage times differ.
    python scripts/tune_traversal.py [c2|c3|c4] [quick]"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from chameleonrt_b200 import RenderCUDA  # noqa: E402

args = [a for a in sys.argv[1:] if a != "quick"]
quick = "quick" in sys.argv
bench.select_workload(args[0] if args else "c2")
scene, view = bench.make_workload()
refills = (4,) if quick else (1, 2, 4, 8)
grid = itertools.product(("host", "device"), refills, (0, 1), (0, 16, 24))
print(f"# {bench.WORKLOAD}; mean of frames 3..7, ms", flush=True)
base = None
for builder, refill, far, defer in grid:
    if builder == "device" and (refill != 4 or quick and (far or defer)):
        continue  # the tree matters through its quality; one line per ordering is enough
    gpu = RenderCUDA(0, max_depth=bench.MAX_DEPTH, bvh_builder=builder, any_far_first=far, tri_pass_defer=defer)
    gpu._check(gpu.lib.crtc_set_option(gpu.h, b"refill_idle", refill))
    gpu.initialize(bench.WIDTH, bench.HEIGHT)
    gpu.set_scene(scene)
    acc = {}
    for f in range(8):
        gpu.render(*view, f == 0, False)
        if f >= 3:
            for k, v in gpu.stage_times().items():
                acc[k] = acc.get(k, 0) + v / 5
    trav = acc["traverse"] + acc["traverse_primary"]
    base = base or trav
    print(f"builder={builder:6s} refill_idle={refill:2d} far_first={far} tri_pass_defer={defer:2d}  traverse {trav:7.3f} "
          f"({trav / base - 1:+.1%})  frame {acc['frame']:7.3f}  " + " ".join(f"{k}={v:.3f}" for k, v in acc.items()
                                                                                if k not in ("frame",)), flush=True)

# The shade queue bucketed by material id (option shade_sort): the sort's launches are inside the "shade" stage, so the
# shade time says whether the more coherent k_shade pays for them; the traversal that follows sees a differently
# ordered queue of continuation rays, so its time is printed as well.
print("# shade_sort (host tree, default traversal)", flush=True)
base_shade = None
for mode in (0, 1, 2):
    gpu = RenderCUDA(0, max_depth=bench.MAX_DEPTH, shade_sort=mode)
    gpu.initialize(bench.WIDTH, bench.HEIGHT)
    gpu.set_scene(scene)
    acc = {}
    for f in range(8):
        gpu.render(*view, f == 0, False)
        if f >= 3:
            for k, v in gpu.stage_times().items():
                acc[k] = acc.get(k, 0) + v / 5
    base_shade = base_shade or (acc["shade"], acc["frame"])
    print(f"shade_sort={mode}  shade {acc['shade']:7.3f} ({acc['shade'] / base_shade[0] - 1:+.1%})  traverse {acc['traverse']:7.3f}  "
          f"frame {acc['frame']:7.3f} ({acc['frame'] / base_shade[1] - 1:+.1%})", flush=True)

