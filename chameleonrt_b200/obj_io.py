"""Write a procedural OBJ-class ``Scene`` as .obj + .mtl + .png files that ChameleonRT's own
loader (``Scene::load_obj``, util/scene.cpp:94-228, tinyobjloader) reads back into the same
``Scene``: one ``g`` group + one ``usemtl`` per geometry (the loader takes a shape's material
from its first face, scene.cpp:125-139), ``Kd``/``Ns``/``map_Kd`` materials (scene.cpp:190-215),
textures flipped vertically because ``Image(file)`` loads with
``stbi_set_flip_vertically_on_load(1)`` (util/material.cpp:8-10).

This is the "data format on the caller's side" of the hot path: it lets
``./chameleonrt cuda scene.obj`` (or the headless twin) render exactly the scenes the Python
tests build in memory.
"""
from __future__ import annotations

import os
import struct

import numpy as np

from .scene import Scene


def _tex_id(x: float):
    bits = struct.unpack("<I", struct.pack("<f", x))[0]
    return (bits & 0x1FFFFFFF) if bits & 0x80000000 else None


def _ns_for_specular(specular: float) -> np.float32:
    spec = np.float32(specular)
    ns = np.float32(spec * np.float32(500.0))
    cands = [ns]
    up = down = ns
    for _ in range(4):
        up = np.nextafter(up, np.float32(np.inf), dtype=np.float32)
        down = np.nextafter(down, np.float32(-np.inf), dtype=np.float32)
        cands += [up, down]
    for c in cands:
        if np.float32(c / np.float32(500.0)) == spec:
            return c
    return ns


def write_obj(scene: Scene, path: str) -> str:
    """Writes ``path`` (.obj), a sibling .mtl and the textures; returns ``path``."""
    from PIL import Image as PILImage

    assert len(scene.meshes) == 1 and len(scene.instances) == 1, "OBJ-class scenes have one mesh, one instance"
    base = os.path.splitext(os.path.basename(path))[0]
    out_dir = os.path.dirname(os.path.abspath(path))
    os.makedirs(out_dir, exist_ok=True)
    for i, t in enumerate(scene.textures):
        PILImage.fromarray(np.ascontiguousarray(t.img[::-1])).save(os.path.join(out_dir, f"{base}_tex{i}.png"))
    with open(os.path.join(out_dir, base + ".mtl"), "w") as f:
        for i, m in enumerate(scene.materials):
            f.write(f"newmtl mat{i}\n")
            tid = _tex_id(m.base_color[0])
            kd = (0.0, m.base_color[1], m.base_color[2]) if tid is not None else m.base_color
            f.write("Kd %.9g %.9g %.9g\n" % tuple(kd))
            # specular = clamp(Ns / 500, 0, 1) (scene.cpp:193): pick the float Ns that maps back to
            # exactly this specular
            f.write("Ns %.9g\n" % _ns_for_specular(m.specular))
            if tid is not None:
                f.write(f"map_Kd {base}_tex{tid}.png\n")
            f.write("\n")
    mat_ids = scene.parameterized_meshes[0].material_ids
    with open(path, "w") as f:
        f.write(f"mtllib {base}.mtl\n")
        voff = 1
        for gi, g in enumerate(scene.meshes[0].geometries):
            v = np.asarray(g.vertices, np.float32)
            has_uv = g.uvs is not None and len(g.uvs)
            f.write(f"g geom{gi}\nusemtl mat{mat_ids[gi]}\n")
            f.write("".join("v %.9g %.9g %.9g\n" % (p[0], p[1], p[2]) for p in v))
            if has_uv:
                f.write("".join("vt %.9g %.9g\n" % (p[0], p[1]) for p in np.asarray(g.uvs, np.float32)))
            idx = np.asarray(g.indices, np.int64) + voff
            if has_uv:
                f.write("".join("f %d/%d %d/%d %d/%d\n" % (a, a, b, b, c, c) for a, b, c in idx))
            else:
                f.write("".join("f %d %d %d\n" % (a, b, c) for a, b, c in idx))
            voff += len(v)
    return path
