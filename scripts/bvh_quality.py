"""CPU-side evaluation of BVH8 quality on the C2 ray population (no GPU needed).

Runs the host instantiation of the product's traversal (libcrt_bvh8_hostcheck.so, the same bvh8_traverse.h the
kernels compile) over primary rays, two generations of cosine-distributed bounce rays and light-sampling
shadow rays of the bench scene, and reports nodes visited / triangles tested per ray: the two quantities the
traversal kernel's time is proportional to. Use it to compare builder variants:
    python scripts/bvh_quality.py [--lib /path/to/other/libcrt_bvh8_hostcheck.so] [--scene sponza_like]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cosine_bounce(rays, hits, normals, rng):
    hit = hits[:, 3].view(np.uint32) != 0xFFFFFFFF
    r, h, n = rays[hit], hits[hit], normals[hit]
    p = r[:, :3] + h[:, :1] * r[:, 4:7]
    n = np.where((np.sum(n * r[:, 4:7], axis=1) > 0)[:, None], -n, n)  # face the incoming ray
    u1, u2 = rng.random(len(p)), rng.random(len(p))
    rad, phi = np.sqrt(u1), 2 * np.pi * u2
    a = np.where(np.abs(n[:, :1]) < 0.6, [[1.0, 0, 0]], [[0, 1.0, 0]])
    t = np.cross(a, n)
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    b = np.cross(n, t)
    d = (rad * np.cos(phi))[:, None] * t + (rad * np.sin(phi))[:, None] * b + np.sqrt(1 - u1)[:, None] * n
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = np.concatenate([p, np.full((len(p), 1), 1e-4), d, np.full((len(p), 1), 1e20)], 1)
    return np.ascontiguousarray(out, np.float32), p.astype(np.float32)


def shadow_rays(points, light, rng):
    pos, vx, vy = (np.array(v, np.float64)[:3] for v in (light.position, light.v_x, light.v_y))
    s = rng.random((len(points), 2))
    lp = pos + (s[:, :1] * light.width) * vx + (s[:, 1:] * light.height) * vy
    d = lp - points
    dist = np.linalg.norm(d, axis=1, keepdims=True)
    d /= dist
    return np.ascontiguousarray(np.concatenate([points, np.full((len(points), 1), 1e-4), d, dist], 1), np.float32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=None)
    ap.add_argument("--scene", default="sponza_like")
    ap.add_argument("--width", type=int, default=480)
    ap.add_argument("--height", type=int, default=270)
    args = ap.parse_args()
    import helpers
    from chameleonrt_b200 import scenes
    from oracle import OracleBackend
    from oracle.oracle import primary_rays

    if args.lib:
        real_cdll = C.CDLL
        helpers.C.CDLL = lambda path, *a, **k: real_cdll(args.lib if "hostcheck" in path else path, *a, **k)
    if args.scene == "materials":
        scene, cam = helpers.synthetic_material_scene()
    elif args.scene == "san_miguel_like":
        scene, cam = scenes.san_miguel_like(spp=1, scale=0.1, tex_size=64)
    elif args.scene == "cornell":
        scene, cam = scenes.cornell_box(spp=1)
    else:
        scene, cam = getattr(scenes, args.scene)(spp=1)
    c = helpers.camera_for(cam)
    hc = helpers.HostCheck(scene)
    st = hc.stats()
    o = OracleBackend(fast=True)
    o.initialize(8, 8)
    o.set_scene(scene)
    rng = np.random.default_rng(7)
    gens = []
    rays = primary_rays(args.width, args.height, c.eye(), c.dir(), c.up(), cam["fov_y"])
    points = []
    for g in range(3):
        hits, normals = o.trace_closest(rays, True)
        gens.append(rays)
        rays, p = cosine_bounce(rays, hits, normals, rng)
        points.append(p)
    sh = shadow_rays(np.concatenate(points), scene.lights[0], rng)
    print(f"scene {args.scene}: {scene.total_tris()} tris, {st['nodes']} nodes, depth {st['depth']}, sah {st['sah']:.3f}, "
          f"build {st['build_ms']:.0f} ms")
    tot_n = tot_t = tot_r = 0
    for name, r in zip(["primary", "bounce1", "bounce2"], gens):
        _, _, cnt = hc.trace(r, counters=True)
        print(f"  closest {name:8s} {len(r):7d} rays: {cnt[:, 0].mean():6.2f} nodes  {cnt[:, 1].mean():6.2f} tris")
        tot_n += cnt[:, 0].sum(); tot_t += cnt[:, 1].sum(); tot_r += len(r)
    print(f"  closest all      {tot_r:7d} rays: {tot_n / tot_r:6.2f} nodes  {tot_t / tot_r:6.2f} tris")
    h_near, _, cnt = hc.trace(sh, any_hit=True, counters=True)
    print(f"  any-hit shadow   {len(sh):7d} rays: {cnt[:, 0].mean():6.2f} nodes  {cnt[:, 1].mean():6.2f} tris  (near-first)")
    h_far, _, cnt = hc.trace(sh, any_hit=True, counters=True, far_first=True)
    occ_near, occ_far = (h[:, 3].view(np.uint32) != 0xFFFFFFFF for h in (h_near, h_far))
    print(f"  any-hit shadow   {len(sh):7d} rays: {cnt[:, 0].mean():6.2f} nodes  {cnt[:, 1].mean():6.2f} tris  (far-first; "
          f"{100 * occ_far.mean():.1f} % occluded, same answers: {np.array_equal(occ_near, occ_far)})")


if __name__ == "__main__":
    main()
