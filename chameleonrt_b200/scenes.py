"""Seeded procedural stand-ins for the benchmark scenes of BASELINE.json.

The real Sponza / San Miguel / Rungholt assets are not on the box and there is no network
(SURVEY.md §8d), so every config has a deterministic generator that produces a ``Scene`` with the
same *structure* the reference's loaders would produce for the real asset:

* OBJ-class scenes (Cornell, Sponza-like, Rungholt-like) follow ``Scene::load_obj``
  (util/scene.cpp:94-228): ONE mesh whose geometries are the OBJ shapes (one per material
  group), ONE parameterized mesh, ONE identity instance, materials mapped from MTL
  (``base_color=Kd``, ``specular=clamp(Ns/500)``, ``roughness=1-specular``, ``map_Kd`` -> sRGB
  RGBA8 texture), and the single generated quad light (scene.cpp:218-227).
* The glTF-class scene (San-Miguel-like) follows ``Scene::load_gltf`` (scene.cpp:230-415):
  many meshes, instances with non-identity transforms, baseColor (sRGB) + metallicRoughness
  (linear; metallic = B channel, roughness = G channel) textures.

Each generator returns ``(scene, camera_args)`` where camera_args = dict(eye, center, up, fov_y).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .scene import (LINEAR, SRGB, DisneyMaterial, Geometry, Image, Instance, Mesh, ParameterizedMesh,
                    QuadLight, Scene, default_obj_light, textured_param)


# ---------------------------------------------------------------------------------------
# mesh building blocks (all return (vertices (n,3) f32, uvs (n,2) f32, indices (m,3) u32))
# ---------------------------------------------------------------------------------------
class MeshBuilder:
    def __init__(self):
        self.v: List[np.ndarray] = []
        self.uv: List[np.ndarray] = []
        self.idx: List[np.ndarray] = []
        self.nv = 0

    def add(self, v, uv, idx):
        v = np.asarray(v, dtype=np.float32).reshape(-1, 3)
        uv = np.asarray(uv, dtype=np.float32).reshape(-1, 2)
        idx = np.asarray(idx, dtype=np.uint32).reshape(-1, 3)
        self.v.append(v)
        self.uv.append(uv)
        self.idx.append(idx + np.uint32(self.nv))
        self.nv += v.shape[0]

    def num_tris(self):
        return sum(i.shape[0] for i in self.idx)

    def geometry(self) -> Geometry:
        return Geometry(
            vertices=np.concatenate(self.v).astype(np.float32),
            indices=np.concatenate(self.idx).astype(np.uint32),
            uvs=np.concatenate(self.uv).astype(np.float32),
        )


def grid_indices(nu: int, nv: int, flip: bool = False) -> np.ndarray:
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * (nv + 1) + j).reshape(-1)
    b = a + (nv + 1)
    c = b + 1
    d = a + 1
    if flip:
        t = np.stack([np.stack([a, c, b], 1), np.stack([a, d, c], 1)], 1)
    else:
        t = np.stack([np.stack([a, b, c], 1), np.stack([a, c, d], 1)], 1)
    return t.reshape(-1, 3).astype(np.uint32)


def grid(p0, du, dv, nu, nv, uv_scale=(1.0, 1.0), flip=False):
    p0, du, dv = (np.asarray(x, dtype=np.float64) for x in (p0, du, dv))
    s = np.linspace(0, 1, nu + 1)
    t = np.linspace(0, 1, nv + 1)
    S, T = np.meshgrid(s, t, indexing="ij")
    v = p0[None, None, :] + S[..., None] * du[None, None, :] + T[..., None] * dv[None, None, :]
    uv = np.stack([S * uv_scale[0], T * uv_scale[1]], -1)
    return v.reshape(-1, 3), uv.reshape(-1, 2), grid_indices(nu, nv, flip)


def box(lo, hi, sub=(1, 1, 1), uv_scale=1.0):
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    d = hi - lo
    ex, ey, ez = np.array([d[0], 0, 0]), np.array([0, d[1], 0]), np.array([0, 0, d[2]])
    sx, sy, sz = sub
    parts = [
        grid(lo, ey, ex, sy, sx, (uv_scale, uv_scale)),  # z = lo (faces -z)
        grid(lo + ez, ex, ey, sx, sy, (uv_scale, uv_scale)),  # z = hi
        grid(lo, ez, ey, sz, sy, (uv_scale, uv_scale)),  # x = lo
        grid(lo + ex, ey, ez, sy, sz, (uv_scale, uv_scale)),  # x = hi
        grid(lo, ex, ez, sx, sz, (uv_scale, uv_scale)),  # y = lo
        grid(lo + ey, ez, ex, sz, sx, (uv_scale, uv_scale)),  # y = hi
    ]
    return parts


def cylinder(base, radius, height, nseg, nring, uv_scale=(1.0, 1.0), radius_fn=None):
    base = np.asarray(base, dtype=np.float64)
    nseg = max(3, nseg)  # two segments would be two coincident sheets
    a = np.linspace(0, 2 * np.pi, nseg + 1)
    h = np.linspace(0, 1, nring + 1)
    A, H = np.meshgrid(a, h, indexing="ij")
    r = radius if radius_fn is None else radius * radius_fn(H, A)
    v = np.stack([base[0] + r * np.cos(A), base[1] + H * height, base[2] + r * np.sin(A)], -1)
    uv = np.stack([A / (2 * np.pi) * uv_scale[0], H * uv_scale[1]], -1)
    return v.reshape(-1, 3), uv.reshape(-1, 2), grid_indices(nseg, nring, flip=True)


def sphere(center, radius, nu, nv, squash=(1.0, 1.0, 1.0)):
    center = np.asarray(center, dtype=np.float64)
    a = np.linspace(0, 2 * np.pi, nu + 1)
    b = np.linspace(1e-3, np.pi - 1e-3, nv + 1)
    A, B = np.meshgrid(a, b, indexing="ij")
    v = np.stack([
        center[0] + radius * squash[0] * np.sin(B) * np.cos(A),
        center[1] + radius * squash[1] * np.cos(B),
        center[2] + radius * squash[2] * np.sin(B) * np.sin(A),
    ], -1)
    uv = np.stack([A / (2 * np.pi), B / np.pi], -1)
    return v.reshape(-1, 3), uv.reshape(-1, 2), grid_indices(nu, nv)


def arch(c0, span_axis, span, rise, thickness, depth_axis, depth, nseg, nprof=2):
    """A half-ring (extruded) spanning ``span`` along span_axis, rising ``rise``."""
    c0 = np.asarray(c0, dtype=np.float64)
    sa = np.asarray(span_axis, dtype=np.float64)
    da = np.asarray(depth_axis, dtype=np.float64)
    upv = np.array([0.0, 1.0, 0.0])
    th = np.linspace(0, np.pi, nseg + 1)
    parts = []
    for (r0, r1, d0, d1, flip) in (
        (1.0, 1.0, 0.0, 1.0, False),  # outer surface
        (1.0 - thickness / (span * 0.5), 1.0 - thickness / (span * 0.5), 0.0, 1.0, True),  # inner
        (1.0 - thickness / (span * 0.5), 1.0, 0.0, 0.0, True),  # front face
        (1.0 - thickness / (span * 0.5), 1.0, 1.0, 1.0, False),  # back face
    ):
        q = np.linspace(0, 1, nprof + 1)
        TH, Q = np.meshgrid(th, q, indexing="ij")
        rr = r0 + (r1 - r0) * Q
        dd = d0 + (d1 - d0) * Q
        if r0 == r1:
            dd = Q
        x = -np.cos(TH) * rr * span * 0.5
        y = np.sin(TH) * rr * rise
        v = (c0[None, None, :] + x[..., None] * sa[None, None, :] + y[..., None] * upv[None, None, :]
             + (dd * depth)[..., None] * da[None, None, :])
        uv = np.stack([TH / np.pi * 4.0, Q], -1)
        parts.append((v.reshape(-1, 3), uv.reshape(-1, 2), grid_indices(nseg, nprof, flip)))
    return parts


# ---------------------------------------------------------------------------------------
# procedural textures (RGBA8, sRGB like map_Kd)
# ---------------------------------------------------------------------------------------
def _value_noise(rng, size, cells):
    g = rng.random((cells + 1, cells + 1))
    x = np.linspace(0, cells, size, endpoint=False)
    xi = x.astype(int)
    xf = x - xi
    xf = xf * xf * (3 - 2 * xf)
    a = g[xi][:, xi]
    b = g[xi + 1][:, xi]
    c = g[xi][:, xi + 1]
    d = g[xi + 1][:, xi + 1]
    fx = xf[:, None]
    fy = xf[None, :]
    return a * (1 - fx) * (1 - fy) + b * fx * (1 - fy) + c * (1 - fx) * fy + d * fx * fy


def _fbm(rng, size, octaves=4, base=4):
    out = np.zeros((size, size))
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        out += amp * _value_noise(rng, size, base * (2 ** o))
        tot += amp
        amp *= 0.5
    return out / tot


def make_texture(kind: str, seed: int, size: int = 1024) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    n = _fbm(rng, size)
    if kind == "bricks":
        bw, bh = size // 8, size // 16
        row = y // bh
        xo = (x + (row % 2) * (bw // 2)) % bw
        mortar = (xo < 4) | (y % bh < 4)
        base = np.array([150, 70, 50])[None, None, :] * (0.7 + 0.5 * n[..., None])
        rgb = np.where(mortar[..., None], np.array([190, 185, 170])[None, None, :], base)
    elif kind == "tiles":
        tw = size // 8
        edge = (x % tw < 3) | (y % tw < 3)
        chk = ((x // tw) + (y // tw)) % 2
        base = np.where(chk[..., None] == 0, np.array([200, 190, 170]), np.array([120, 110, 100]))
        rgb = np.where(edge[..., None], np.array([60, 60, 60])[None, None, :], base * (0.8 + 0.3 * n[..., None]))
    elif kind == "marble":
        v = np.sin((x / size * 6 + n * 5) * np.pi) * 0.5 + 0.5
        rgb = (np.array([230, 225, 215])[None, None, :] * (0.55 + 0.45 * v[..., None]))
    elif kind == "fabric":
        w = (np.sin(x * 0.6) * 0.5 + 0.5) * (np.sin(y * 0.6) * 0.5 + 0.5)
        col = np.array([[170, 30, 30], [30, 120, 50], [40, 60, 160]][seed % 3])
        rgb = col[None, None, :] * (0.6 + 0.4 * w[..., None]) * (0.8 + 0.3 * n[..., None])
    elif kind == "plaster":
        rgb = np.array([205, 195, 170])[None, None, :] * (0.75 + 0.35 * n[..., None])
    elif kind == "wood":
        v = np.sin((y / size * 24 + n * 3) * np.pi) * 0.5 + 0.5
        rgb = np.array([140, 95, 55])[None, None, :] * (0.6 + 0.4 * v[..., None])
    else:  # "stone"
        rgb = np.array([150, 150, 145])[None, None, :] * (0.6 + 0.6 * n[..., None])
    rgba = np.empty((size, size, 4), dtype=np.uint8)
    rgba[..., :3] = np.clip(rgb, 0, 255).astype(np.uint8)
    rgba[..., 3] = 255
    return rgba


def _obj_material(kd, ns, tex_id=None) -> DisneyMaterial:
    """MTL -> DisneyMaterial exactly as load_obj does (util/scene.cpp:190-215)."""
    spec = float(np.clip(np.float32(ns) / np.float32(500.0), 0.0, 1.0))
    rough = float(np.clip(np.float32(1.0) - np.float32(spec), 0.0, 1.0))
    base = tuple(float(x) for x in kd)
    if tex_id is not None:
        base = (textured_param(tex_id), base[1], base[2])
    return DisneyMaterial(base_color=base, specular=spec, roughness=rough, specular_transmission=0.0)


def _obj_scene(builders: List[Tuple[MeshBuilder, int]], materials, textures, spp) -> Scene:
    mesh = Mesh([b.geometry() for b, _ in builders])
    scene = Scene(
        meshes=[mesh],
        parameterized_meshes=[ParameterizedMesh(0, [m for _, m in builders])],
        instances=[Instance(np.eye(4, dtype=np.float32), 0)],
        materials=materials,
        textures=textures,
        lights=[default_obj_light()],
        samples_per_pixel=spp,
    )
    scene.validate_materials()
    return scene


# ---------------------------------------------------------------------------------------
# C1: Cornell box
# ---------------------------------------------------------------------------------------
def cornell_box(spp: int = 1):
    """5 walls + 2 boxes, 3 diffuse Kd materials (white/red/green). 34 triangles."""
    mats = [
        _obj_material((0.73, 0.73, 0.73), 1.0),  # white (tinyobj default Ns=1)
        _obj_material((0.65, 0.05, 0.05), 1.0),  # red
        _obj_material((0.12, 0.45, 0.15), 1.0),  # green
    ]
    white, red, green = MeshBuilder(), MeshBuilder(), MeshBuilder()
    # room x in [-1,1], y in [0,2], z in [-1,1], open toward +z
    white.add(*grid((-1, 0, -1), (0, 0, 2), (2, 0, 0), 1, 1))  # floor (normal +y)
    white.add(*grid((-1, 2, -1), (2, 0, 0), (0, 0, 2), 1, 1))  # ceiling
    white.add(*grid((-1, 0, -1), (2, 0, 0), (0, 2, 0), 1, 1))  # back wall
    red.add(*grid((-1, 0, -1), (0, 2, 0), (0, 0, 2), 1, 1))  # left wall
    green.add(*grid((1, 0, -1), (0, 0, 2), (0, 2, 0), 1, 1))  # right wall

    def rot_box(lo, hi, angle, builder):
        c, s = np.cos(angle), np.sin(angle)
        ctr = (np.asarray(lo) + np.asarray(hi)) * 0.5
        for v, uv, idx in box(lo, hi):
            p = v - ctr
            q = np.stack([c * p[:, 0] + s * p[:, 2], p[:, 1], -s * p[:, 0] + c * p[:, 2]], 1) + ctr
            builder.add(q, uv, idx)

    rot_box((-0.7, 0.0, -0.6), (-0.1, 1.2, 0.0), 0.3, white)  # tall box
    rot_box((0.1, 0.0, 0.0), (0.7, 0.6, 0.6), -0.3, white)  # short box
    scene = _obj_scene([(white, 0), (red, 1), (green, 2)], mats, [], spp)
    cam = dict(eye=(0.0, 1.0, 3.4), center=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), fov_y=40.0)
    return scene, cam


# ---------------------------------------------------------------------------------------
# C2: Sponza-like atrium (~262 K triangles, ~25 materials, 8 map_Kd textures)
# ---------------------------------------------------------------------------------------
def sponza_like(spp: int = 4, seed: int = 0x5002A, detail: float = 1.0, tex_size: int = 1024):
    """Two-storey colonnaded atrium, x in [-15,15], z in [-7,7], y in [0,12].

    ``detail`` scales tessellation (1.0 -> ~262 K triangles like Crytek Sponza's 262,267).
    """
    rng = np.random.default_rng(seed)
    kinds = ["bricks", "tiles", "marble", "fabric", "fabric", "fabric", "plaster", "stone"]
    textures = [Image(f"tex_{k}_{i}.png", make_texture(k, seed + i, tex_size), SRGB) for i, k in enumerate(kinds)]
    T_BRICK, T_TILE, T_MARBLE, T_FAB_R, T_FAB_G, T_FAB_B, T_PLASTER, T_STONE = range(8)

    materials: List[DisneyMaterial] = []
    builders: List[Tuple[MeshBuilder, int]] = []

    def new_group(kd, ns, tex=None) -> MeshBuilder:
        materials.append(_obj_material(kd, ns, tex))
        b = MeshBuilder()
        builders.append((b, len(materials) - 1))
        return b

    def n_(x):
        return max(1, int(round(x * detail)))

    floor = new_group((0.6, 0.6, 0.6), 40.0, T_TILE)
    walls = new_group((0.6, 0.5, 0.4), 5.0, T_BRICK)
    plaster = new_group((0.7, 0.7, 0.6), 2.0, T_PLASTER)
    gallery = new_group((0.5, 0.5, 0.5), 10.0, T_STONE)
    col_a = new_group((0.8, 0.8, 0.75), 120.0, T_MARBLE)
    col_b = new_group((0.75, 0.75, 0.8), 200.0, T_MARBLE)
    col_base = new_group((0.5, 0.5, 0.5), 20.0, T_STONE)
    arches = new_group((0.7, 0.65, 0.55), 8.0, T_PLASTER)
    arches2 = new_group((0.65, 0.6, 0.55), 8.0, T_BRICK)
    drape_r = new_group((0.6, 0.1, 0.1), 3.0, T_FAB_R)
    drape_g = new_group((0.1, 0.5, 0.2), 3.0, T_FAB_G)
    drape_b = new_group((0.1, 0.2, 0.6), 3.0, T_FAB_B)
    vase_gold = new_group((0.9, 0.7, 0.3), 420.0)
    vase_cu = new_group((0.8, 0.45, 0.3), 350.0)
    vase_cer = new_group((0.85, 0.85, 0.9), 480.0)
    plain = [new_group(tuple(0.25 + 0.6 * rng.random(3)), float(rng.choice([1.0, 15.0, 60.0, 250.0])))
             for _ in range(10)]

    # floor, outer walls, ceiling ring
    floor.add(*grid((-15, 0, -7), (0, 0, 14), (30, 0, 0), n_(28), n_(60), (7, 15)))
    walls.add(*grid((-15, 0, -7), (30, 0, 0), (0, 12, 0), n_(60), n_(24), (10, 4)))
    walls.add(*grid((15, 0, 7), (-30, 0, 0), (0, 12, 0), n_(60), n_(24), (10, 4)))
    walls.add(*grid((-15, 0, 7), (0, 0, -14), (0, 12, 0), n_(28), n_(24), (5, 4)))
    walls.add(*grid((15, 0, -7), (0, 0, 14), (0, 12, 0), n_(28), n_(24), (5, 4)))
    # roof over the side aisles (the nave stays open to the sky)
    plaster.add(*grid((-15, 12, -7), (30, 0, 0), (0, 0, 3.5), n_(60), n_(7), (10, 1)))
    plaster.add(*grid((-15, 12, 3.5), (30, 0, 0), (0, 0, 3.5), n_(60), n_(7), (10, 1)))
    # gallery slabs (second floor) along both aisles
    for z0, z1 in ((-7.0, -3.2), (3.2, 7.0)):
        for part in box((-15, 5.6, z0), (15, 6.0, z1), (n_(40), 1, n_(6)), 4.0):
            gallery.add(*part)
    # columns: 2 rows x 12 x 2 storeys
    xs = np.linspace(-13.2, 13.2, 12)
    for storey, (y0, hgt) in enumerate(((0.0, 5.6), (6.0, 5.4))):
        for zi, z in enumerate((-3.5, 3.5)):
            for ci, x in enumerate(xs):
                col = col_a if (ci + zi + storey) % 2 == 0 else col_b
                flute = lambda H, A: 1.0 + 0.04 * np.cos(A * 12) - 0.12 * H  # noqa: E731
                col.add(*cylinder((x, y0 + 0.4, z), 0.32, hgt - 0.8, n_(36), n_(18), (2, 4), flute))
                for part in box((x - 0.45, y0, z - 0.45), (x + 0.45, y0 + 0.4, z + 0.45), (2, 1, 2)):
                    col_base.add(*part)
                for part in box((x - 0.42, y0 + hgt - 0.4, z - 0.42), (x + 0.42, y0 + hgt, z + 0.42), (2, 1, 2)):
                    col_base.add(*part)
    # arches between neighbouring columns
    dx = xs[1] - xs[0]
    for storey, ytop in enumerate((4.2, 10.2)):
        for zi, z in enumerate((-3.5, 3.5)):
            for ci in range(len(xs) - 1):
                tgt = arches if (ci + storey) % 2 == 0 else arches2
                for part in arch((xs[ci] + dx / 2, ytop, z - 0.3), (1, 0, 0), dx - 0.7, 1.2, 0.35, (0, 0, 1), 0.6,
                                 n_(40), 3):
                    tgt.add(*part)
    # drapes hanging across the nave
    drapes = (drape_r, drape_g, drape_b)
    for di in range(12):
        x = -12.5 + di * 2.27
        nu, nv = n_(40), n_(40)
        v, uv, idx = grid((x, 11.0, -3.0), (0, 0, 6.0), (0, -4.5, 0), nu, nv, (2, 2))
        s = (v[:, 2] + 3.0) / 6.0
        sag = 4.0 * s * (1 - s)
        v[:, 1] -= 1.2 * sag
        v[:, 0] += 0.18 * np.sin(v[:, 2] * 4.0 + di) * (11.0 - v[:, 1]) / 4.5
        drapes[di % 3].add(v, uv, idx)
        drapes[di % 3].add(v + np.array([0.02, 0, 0]), uv, idx[:, ::-1])  # back side
    # vases / ornaments along the nave
    vases = (vase_gold, vase_cu, vase_cer)
    for vi in range(8):
        x = -12.0 + vi * 3.4
        z = 1.6 if vi % 2 else -1.6
        tgt = vases[vi % 3]
        prof = lambda H, A: 0.55 + 0.45 * np.sin(H * np.pi * 0.9 + 0.3)  # noqa: E731
        tgt.add(*cylinder((x, 0.0, z), 0.5, 1.3, n_(48), n_(32), (1, 1), prof))
        tgt.add(*sphere((x, 1.75, z), 0.38, n_(40), n_(24)))
    # loose blocks / crates / debris with the plain materials (incoherent small geometry)
    for bi in range(int(220 * detail)):
        tgt = plain[bi % len(plain)]
        c = np.array([rng.uniform(-14, 14), 0.0, rng.uniform(-6.5, 6.5)])
        if abs(abs(c[2]) - 3.5) < 0.8:
            c[2] += 1.2 * np.sign(c[2])
        sz = rng.uniform(0.1, 0.45, 3)
        yb = 6.0 if (bi % 5 == 0 and abs(c[2]) > 3.4) else 0.0
        for part in box(c - sz * [1, 0, 1] + [0, yb, 0], c + sz * [1, 2, 1] + [0, yb, 0], (n_(3), n_(3), n_(3))):
            tgt.add(*part)

    builders = [(b, m) for b, m in builders if b.num_tris() > 0]
    scene = _obj_scene(builders, materials, textures, spp)
    cam = dict(eye=(-12.5, 2.2, 0.3), center=(6.0, 4.2, -0.4), up=(0.0, 1.0, 0.0), fov_y=65.0)
    return scene, cam


# ---------------------------------------------------------------------------------------
# C4: Rungholt-like voxel city (~6.7 M triangles at scale=1, untextured, ~80 materials)
# ---------------------------------------------------------------------------------------
def rungholt_like(spp: int = 4, seed: int = 0x2C401, scale: float = 1.0):
    """Voxel city from a value-noise height map; only exposed block faces are emitted."""
    rng = np.random.default_rng(seed)
    n = max(16, int(round(1100 * np.sqrt(scale))))
    size = 1
    while size < n:
        size *= 2
    h = _fbm(rng, size, octaves=5, base=4)[:n, :n]
    h = (h - h.min()) / (h.max() - h.min())
    height = np.floor(2 + 14 * h ** 2).astype(np.int32)
    # buildings: random boxes raised above the terrain
    nb = int(2600 * scale)
    bx = rng.integers(0, n - 12, nb)
    bz = rng.integers(0, n - 12, nb)
    bw = rng.integers(3, 12, nb)
    bd = rng.integers(3, 12, nb)
    bh = rng.integers(4, 40, nb)
    matmap = (rng.integers(0, 20, (n, n))).astype(np.int32)
    for i in range(nb):
        sl = (slice(bx[i], bx[i] + bw[i]), slice(bz[i], bz[i] + bd[i]))
        base = height[sl].min()
        height[sl] = base + bh[i]
        matmap[sl] = 20 + (i % 60)
    nmat = 80
    materials = [_obj_material(tuple(0.15 + 0.75 * rng.random(3)), float(rng.choice([1.0, 10.0, 80.0, 300.0])))
                 for _ in range(nmat)]
    vox = 0.25
    quads_by_mat: Dict[int, List[np.ndarray]] = {}

    def emit(mat_ids, p0, du, dv):
        # p0,du,dv: (k,3) arrays
        for m in np.unique(mat_ids):
            sel = mat_ids == m
            q = np.stack([p0[sel], p0[sel] + du[sel], p0[sel] + du[sel] + dv[sel], p0[sel] + dv[sel]], 1)
            quads_by_mat.setdefault(int(m), []).append(q)

    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    ii = ii.reshape(-1)
    jj = jj.reshape(-1)
    hh = height.reshape(-1)
    mm = matmap.reshape(-1)
    x0 = (ii - n / 2) * vox
    z0 = (jj - n / 2) * vox
    k = len(ii)
    zeros = np.zeros(k)
    ones = np.ones(k)
    # top faces
    emit(mm, np.stack([x0, hh * vox, z0], 1), np.stack([zeros, zeros, ones * vox], 1),
         np.stack([ones * vox, zeros, zeros], 1))
    # side faces: for each of 4 neighbours, one quad per exposed voxel layer
    hp = np.pad(height, 1, constant_values=0)
    for (di, dj, p_off, du_, dv_) in (
        (1, 0, (vox, 0, 0), (0, 0, vox), (0, vox, 0)),
        (-1, 0, (0, 0, vox), (0, 0, -vox), (0, vox, 0)),
        (0, 1, (vox, 0, vox), (-vox, 0, 0), (0, vox, 0)),
        (0, -1, (0, 0, 0), (vox, 0, 0), (0, vox, 0)),
    ):
        nh = hp[1 + di:n + 1 + di, 1 + dj:n + 1 + dj].reshape(-1)
        expo = np.maximum(hh - nh, 0)
        maxe = int(expo.max())
        for layer in range(maxe):
            sel = expo > layer
            if not sel.any():
                break
            yb = (hh[sel] - 1 - layer) * vox
            p0 = np.stack([x0[sel] + p_off[0], yb + p_off[1], z0[sel] + p_off[2]], 1)
            ks = int(sel.sum())
            emit(mm[sel], p0, np.tile(np.array(du_), (ks, 1)), np.tile(np.array(dv_), (ks, 1)))
    builders = []
    for m, lst in sorted(quads_by_mat.items()):
        q = np.concatenate(lst).astype(np.float32)  # (nq,4,3)
        nq = q.shape[0]
        v = q.reshape(-1, 3)
        base = (np.arange(nq, dtype=np.uint32) * 4)[:, None]
        idx = np.concatenate([base + np.array([0, 1, 2], dtype=np.uint32), base + np.array([0, 2, 3], dtype=np.uint32)], 0)
        b = MeshBuilder()
        b.add(v, np.zeros((v.shape[0], 2), dtype=np.float32), idx)
        builders.append((b, m % nmat))
    mesh = Mesh([Geometry(b.v[0], b.idx[0], None) for b, _ in builders])  # untextured: no uvs
    scene = Scene(meshes=[mesh], parameterized_meshes=[ParameterizedMesh(0, [m for _, m in builders])],
                  instances=[Instance(np.eye(4, dtype=np.float32), 0)], materials=materials, textures=[],
                  lights=[default_obj_light()], samples_per_pixel=spp)
    scene.validate_materials()
    ext = n * vox / 2
    cam = dict(eye=(-0.62 * ext, 0.42 * ext, 0.75 * ext), center=(0.0, 1.0, 0.0), up=(0.0, 1.0, 0.0), fov_y=55.0)
    return scene, cam


# ---------------------------------------------------------------------------------------
# C3/C5: San-Miguel-like instanced, textured glTF-class scene
# ---------------------------------------------------------------------------------------
def _trs(t, angle_y, s):
    c, si = np.cos(angle_y), np.sin(angle_y)
    m = np.array([[c * s, 0, si * s, t[0]], [0, s, 0, t[1]], [-si * s, 0, c * s, t[2]], [0, 0, 0, 1]], dtype=np.float32)
    return m


def san_miguel_like(spp: int = 8, seed: int = 0x5A11, scale: float = 1.0, tex_size: int = 1024):
    """Courtyard with instanced foliage. ~10 M instanced triangles at scale=1.

    glTF-class structure (util/scene.cpp:230-415): several meshes, one parameterized mesh per
    (mesh, material set), many instances with non-identity TRS, baseColor (sRGB) and
    metallicRoughness (linear: G=roughness, B=metallic) textures.
    """
    rng = np.random.default_rng(seed)
    textures: List[Image] = []
    materials: List[DisneyMaterial] = []

    def add_tex(kind, s, cs):
        textures.append(Image(f"{kind}_{len(textures)}", make_texture(kind, s, tex_size), cs))
        return len(textures) - 1

    def gltf_material(base_tex=None, mr_tex=None, base=(0.8, 0.8, 0.8), metallic=0.0, roughness=0.8):
        bc = (textured_param(base_tex), base[1], base[2]) if base_tex is not None else base
        m = DisneyMaterial(base_color=bc, metallic=metallic, roughness=roughness, specular=0.0)
        if mr_tex is not None:
            m.metallic = textured_param(mr_tex, 2)
            m.roughness = textured_param(mr_tex, 1)
        materials.append(m)
        return len(materials) - 1

    kinds = ["tiles", "bricks", "plaster", "wood", "stone", "marble", "fabric"]
    base_tex = [add_tex(k, seed + i, SRGB) for i, k in enumerate(kinds)]
    mr_tex = [add_tex("stone", seed + 100 + i, LINEAR) for i in range(4)]
    mats_struct = [gltf_material(base_tex[i % len(base_tex)], mr_tex[i % len(mr_tex)]) for i in range(60)]
    leaf_mats = [gltf_material(None, None, (0.1 + 0.2 * rng.random(), 0.35 + 0.4 * rng.random(), 0.08), 0.0, 0.6)
                 for _ in range(30)]
    trunk_mats = [gltf_material(base_tex[3], mr_tex[1]) for _ in range(10)]

    meshes: List[Mesh] = []
    pms: List[ParameterizedMesh] = []
    instances: List[Instance] = []

    def d_(x):
        return max(1, int(round(x * np.sqrt(scale))))

    # mesh 0: courtyard architecture (single instance, identity-ish transform)
    arch_geoms = []
    arch_mats = []

    def add_arch_geom(parts, mat):
        b = MeshBuilder()
        for p in parts:
            b.add(*p)
        arch_geoms.append(b.geometry())
        arch_mats.append(mat)

    add_arch_geom([grid((-20, 0, -20), (0, 0, 40), (40, 0, 0), d_(160), d_(160), (20, 20))], mats_struct[0])
    for wi, (p0, du) in enumerate((((-20, 0, -20), (40, 0, 0)), ((20, 0, 20), (-40, 0, 0)),
                                   ((-20, 0, 20), (0, 0, -40)), ((20, 0, -20), (0, 0, 40)))):
        add_arch_geom([grid(p0, du, (0, 14, 0), d_(160), d_(56), (12, 4))], mats_struct[1 + wi])
    for ci in range(28):
        ang = ci / 28 * 2 * np.pi
        x, z = 13.5 * np.cos(ang), 13.5 * np.sin(ang)
        add_arch_geom([cylinder((x, 0, z), 0.4, 6.0, d_(32), d_(24), (2, 4))], mats_struct[5 + ci % 20])
        add_arch_geom(box((x - 0.6, 6.0, z - 0.6), (x + 0.6, 6.5, z + 0.6), (2, 1, 2)), mats_struct[25 + ci % 10])
    for ti in range(16):
        x, z = rng.uniform(-10, 10), rng.uniform(-10, 10)
        add_arch_geom(box((x - 0.8, 0.0, z - 0.5), (x + 0.8, 0.8, z + 0.5), (d_(6), d_(4), d_(4))),
                      mats_struct[35 + ti % 20])
    meshes.append(Mesh(arch_geoms))
    pms.append(ParameterizedMesh(0, arch_mats))
    instances.append(Instance(_trs((0, 0, 0), 0.0, 1.0), 0))

    # foliage meshes: a few tree / bush prototypes, instanced many times
    n_proto = 6
    for pi in range(n_proto):
        trunk = MeshBuilder()
        trunk.add(*cylinder((0, 0, 0), 0.18, 3.0, d_(16), d_(12), (1, 3), lambda H, A: 1.0 - 0.5 * H))
        leaves = MeshBuilder()
        nleaf = int(5200 * scale)
        prng = np.random.default_rng(seed + 1000 + pi)
        ctr = prng.normal(0, 1, (nleaf, 3)) * np.array([1.3, 0.9, 1.3]) + np.array([0, 3.6, 0])
        a = prng.normal(0, 1, (nleaf, 3))
        a /= np.linalg.norm(a, axis=1, keepdims=True)
        b = np.cross(a, prng.normal(0, 1, (nleaf, 3)))
        b /= np.linalg.norm(b, axis=1, keepdims=True)
        sz = prng.uniform(0.08, 0.2, (nleaf, 1))
        q = np.stack([ctr - a * sz, ctr + b * sz * 0.5, ctr + a * sz, ctr - b * sz * 0.5], 1)
        v = q.reshape(-1, 3)
        base = (np.arange(nleaf, dtype=np.uint32) * 4)[:, None]
        idx = np.concatenate([base + np.array([0, 1, 2], dtype=np.uint32), base + np.array([0, 2, 3], dtype=np.uint32)], 0)
        uv = np.tile(np.array([[0, 0], [1, 0], [1, 1], [0, 1]], dtype=np.float32), (nleaf, 1))
        leaves.add(v, uv, idx)
        meshes.append(Mesh([trunk.geometry(), leaves.geometry()]))
    n_inst = int(900 * min(1.0, scale) + 60)
    for ii in range(n_inst):
        proto = ii % n_proto
        pms.append(ParameterizedMesh(1 + proto, [trunk_mats[ii % len(trunk_mats)], leaf_mats[ii % len(leaf_mats)]]))
        r = rng.uniform(2.0, 19.0)
        ang = rng.uniform(0, 2 * np.pi)
        instances.append(Instance(_trs((r * np.cos(ang), 0.0, r * np.sin(ang)), rng.uniform(0, 2 * np.pi),
                                       rng.uniform(0.6, 1.5)), len(pms) - 1))
    # glTF scenes carry no light; load_gltf adds the same default quad light when none exists
    # (scene.cpp:402-414)
    scene = Scene(meshes=meshes, parameterized_meshes=pms, instances=instances, materials=materials,
                  textures=textures, lights=[default_obj_light()], samples_per_pixel=spp)
    scene.validate_materials()
    cam = dict(eye=(16.0, 3.0, 17.0), center=(0.0, 2.5, 0.0), up=(0.0, 1.0, 0.0), fov_y=60.0)
    return scene, cam


SCENES = {
    "cornell": cornell_box,
    "sponza_like": sponza_like,
    "rungholt_like": rungholt_like,
    "san_miguel_like": san_miguel_like,
}
