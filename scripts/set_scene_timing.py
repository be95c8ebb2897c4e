"""set_scene of the bench scenes with each BVH builder (host binned SAH / device PLOC / device LBVH): wall-clock around the
call, the builder's own time, tree size, and what each tree costs to traverse (instrumented node visits and the
traverse stage time of two frames at a reduced size), frames compared bit for bit. GPU only.
    python scripts/set_scene_timing.py [c2 c3 c4 ...]        (default: c2 c4)
For an ncu launch list of one device build:  ncu --metrics gpu__time_duration.sum --clock-control none --csv
    --log-file gpurun_out/device_build_launches.csv python scripts/set_scene_timing.py c2 --only device --no-render"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import bench  # noqa: E402
from chameleonrt_b200 import RenderCUDA  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else None
    if only in args:
        args.remove(only)
    render = "--no-render" not in sys.argv
    for key in args or ["c2", "c4"]:
        bench.select_workload(key)
        t0 = time.perf_counter()
        scene, view = bench.make_workload()
        print(f"# {bench.WORKLOAD}: generated in {time.perf_counter() - t0:.1f} s, {scene.total_tris()} triangles", flush=True)
        rows, frames = {}, {}
        for builder in ("host", "device", "device_lbvh"):
            if only and builder != only:
                continue
            r = RenderCUDA(0, max_depth=bench.MAX_DEPTH, bvh_builder=builder, count_traversal=render, any_far_first=0)
            r.initialize(640, 360)
            walls = []
            for _ in range(3):  # the first call pays for allocations and, on the host path, cold pages
                t0 = time.perf_counter()
                r.set_scene(scene)
                walls.append((time.perf_counter() - t0) * 1e3)
            info = r.scene_info()
            row = {"set_scene_ms": [round(w, 2) for w in walls], "bvh_build_ms": round(info["bvh_build_ms"], 3),
                   "phases_ms": {k: round(info[k + "_ms"], 3) for k in ("flatten", "sort", "tree", "emit", "pack")},
                   "bvh8_nodes": int(info["bvh8_nodes"]), "bvh8_depth": int(info["bvh8_depth"]),
                   "ploc_rounds": int(info["ploc_rounds"]) if builder == "device" else None}
            if render:
                for f in range(2):
                    r.render(*view, f == 0, False)
                c, st = r.counters(), r.stage_times()
                row.update(closest_nodes_per_ray=round(c["closest_nodes_visited"] / max(1, c["closest_rays"]), 3),
                           any_hit_nodes_per_ray=round(c["any_nodes_visited"] / max(1, c["occlusion_rays"]), 3),
                           traverse_ms_instrumented=round(st["traverse"] + st["traverse_primary"], 3))
                frames[builder] = r.read_accum()
            rows[builder] = row
            print(json.dumps({builder: row}), flush=True)
            del r
        if render and "host" in frames:
            for b in frames:
                if b != "host":
                    print(f"# {b}: frames bit-identical to the host-built tree's: "
                          f"{np.array_equal(frames['host'].view(np.uint32), frames[b].view(np.uint32))}", flush=True)


if __name__ == "__main__":
    main()
