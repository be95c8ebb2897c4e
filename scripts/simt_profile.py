"""Warp-level work of k_traverse, measured WITHOUT a GPU: runs the production kernel under the SIMT emulation of
chameleonrt_b200/csrc/simt_hostcheck.cpp on rays of the bench scene and counts, per warp, how often the node phase
and the triangle pass execute and how many lanes do useful work in them. ncu shows the kernel bound by instruction
issue (issue-active ~78 %, 21-24 of 32 lanes), so these counts are a proxy for its run time; scheduling and ordering
choices can be compared on them (refill threshold, far-first shadow rays, sorting the rays of a launch).
    python scripts/simt_profile.py [--width 96 --height 54] [--scene sponza_like]"""
import argparse
import ctypes as C
import os
import sys
import time
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)

from bvh_quality import cosine_bounce, shadow_rays  # noqa: E402
from chameleonrt_b200 import scenes  # noqa: E402
from chameleonrt_b200.scene import CScene  # noqa: E402
import helpers  # noqa: E402
from oracle import OracleBackend  # noqa: E402
from oracle.oracle import primary_rays  # noqa: E402

C_NODE, C_TRI = 1.0, 0.45  # relative cost of one node-phase execution and one triangle pass (SASS instruction counts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="sponza_like")
    ap.add_argument("--width", type=int, default=96)
    ap.add_argument("--height", type=int, default=54)
    args = ap.parse_args()
    lib = C.CDLL(os.path.join(ROOT, "chameleonrt_b200", "csrc", "libcrt_simt_hostcheck.so"))
    lib.crt_simt_create.restype = C.c_void_p
    lib.crt_simt_create.argtypes = [C.POINTER(CScene)]
    lib.crt_simt_traverse.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.crt_simt_profile.argtypes = [C.c_void_p, C.c_int]
    scene, cam = (scenes.san_miguel_like(spp=1, scale=0.1, tex_size=64) if args.scene == "san_miguel_like" else getattr(scenes, args.scene)(spp=1))
    c = helpers.camera_for(cam)
    ms = scene.to_c()
    h = lib.crt_simt_create(C.byref(ms.c))
    o = OracleBackend(fast=True)
    o.initialize(8, 8)
    o.set_scene(scene)
    rng = np.random.default_rng(7)
    prim = primary_rays(args.width, args.height, c.eye(), c.dir(), c.up(), cam["fov_y"])
    hits, normals = o.trace_closest(prim, True)
    b1, p0 = cosine_bounce(prim, hits, normals, rng)
    sh0 = shadow_rays(p0, scene.lights[0], rng)
    sh0[:, 3] = 1e-4

    def run(closest, shadow, sched, label):
        prof = (C.c_ulonglong * 4)()
        lib.crt_simt_profile(prof, 1)
        out, vis = np.zeros((max(1, len(closest)), 4), np.float32), np.zeros(max(1, len(shadow)), np.uint8)
        t = time.time()
        lib.crt_simt_traverse(h, closest.ctypes.data, len(closest), None, shadow.ctypes.data, len(shadow), 1, sched, out.ctypes.data, vis.ctypes.data)
        lib.crt_simt_profile(prof, 0)
        pn, ln, pt, lt = (int(x) for x in prof)
        n = len(closest) + len(shadow)
        cost = (pn * C_NODE + pt * C_TRI) / max(1, n)
        print(f"  {label:54s} node phases/ray {pn / n:6.3f} ({ln / max(1, pn):4.1f} lanes)  tri passes/ray {pt / n:6.3f} ({lt / max(1, pt):4.1f} lanes)  "
              f"cost/ray {cost:6.3f}   [{time.time() - t:.0f}s]")
        return cost

    def sort_key(r):
        octant = (r[:, 4] < 0) * 4 + (r[:, 5] < 0) * 2 + (r[:, 6] < 0) * 1
        return np.lexsort((r[:, 2], r[:, 0], octant))

    empty = np.zeros((0, 8), np.float32)
    print(f"scene {args.scene}, {len(prim)} primary rays, {len(b1)} bounce rays, {len(sh0)} shadow rays (one block of 4 warps drains each launch)")
    print("primary rays (coherent):")
    run(prim, empty, 4, "refill at 4 idle lanes (default)")
    run(prim, empty, 4 | (16 << 16), "triangle pass deferred until 16 pairs are pooled")
    print("the merged launch of bounce 0: its shadow rays + the continuation rays of bounce 1:")
    base = run(b1, sh0, 4, "default: refill at 4 idle lanes, near-first")
    for ri in (1, 8, 16, 32):
        run(b1, sh0, ri, f"refill at {ri} idle lanes")
    run(b1, sh0, 4 | 0x100, "shadow rays far-first")
    # bits 16-23: the instantiation (test-only selector of simt_hostcheck.cpp) = option tri_pass_defer
    run(b1, sh0, 4 | (16 << 16), "triangle pass deferred until 16 pairs are pooled")
    run(b1, sh0, 4 | (24 << 16), "triangle pass deferred until 24 pairs are pooled")
    run(b1, sh0, 4 | 0x100 | (16 << 16), "deferred (16) + shadow rays far-first")
    run(b1, sh0, 1 | 0x100 | (16 << 16), "deferred (16) + far-first + refill at 1 idle lane")
    run(np.ascontiguousarray(b1[sort_key(b1)]), np.ascontiguousarray(sh0[sort_key(sh0)]), 4, "rays sorted by octant, then position")
    run(np.ascontiguousarray(b1[sort_key(b1)]), np.ascontiguousarray(sh0[sort_key(sh0)]), 4 | 0x100, "sorted + shadow rays far-first")
    print("the same rays as two launches (shadow rays alone, continuation rays alone):")
    run(empty, sh0, 4, "shadow rays only")
    run(b1, empty, 4, "continuation rays only")


if __name__ == "__main__":
    main()
