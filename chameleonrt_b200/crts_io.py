"""Writes a ``Scene`` in the reference's own binary .crts format (util/scene.cpp:417-620: uint64 header
size, JSON header, data block), the only format whose importer reads EVERY Disney parameter, textured scalar
parameters (texture id + channel per parameter) and explicit quad lights. Used to push the all-lobes test scene
through the reference's loader (oracle/_ref/crt_headless <backend> scene.crts).

Constraints of the format, mirrored here: one geometry per mesh; objects carry (mesh, material, matrix), so a
parameterized mesh is a (mesh, material) pair; images are embedded encoded files which the loader decodes with a
vertical flip and expands to RGBA; a LIGHT object's frame is its matrix (x axis, y axis, -normal, position), which
the loader normalises — pass axis-aligned lights if the loaded scene has to be bit-identical.
"""
from __future__ import annotations

import io
import json
import struct

import numpy as np

from .scene import LINEAR, Scene, f32_bits


def _handle(v: float):
    bits = f32_bits(v)
    if bits & 0x80000000:
        return bits & 0x1FFFFFFF, (bits >> 29) & 0x3
    return None


def _num(v: float) -> float:
    return float(np.float32(v))


def write_crts(scene: Scene, path: str, extra_objects=(), align: bool = True) -> str:
    """``extra_objects``: header objects appended as they are (e.g. CAMERA objects, lights with an arbitrary frame);
    ``align=False`` leaves the data block wherever the header ends (the loader must not rely on alignment)."""
    from PIL import Image as PILImage

    data = bytearray()
    views = []

    def add_view(raw: bytes, dtype: str) -> int:
        while len(data) % 8:
            data.append(0)
        views.append({"byte_offset": len(data), "byte_length": len(raw), "type": dtype})
        data.extend(raw)
        return len(views) - 1

    # CRTS meshes hold one geometry each: split every Mesh into its geometries
    geom_mesh_id = {}
    meshes = []
    for mi, mesh in enumerate(scene.meshes):
        for gi, g in enumerate(mesh.geometries):
            m = {"positions": add_view(np.ascontiguousarray(g.vertices, np.float32).tobytes(), "VEC3_F32"),
                 "indices": add_view(np.ascontiguousarray(g.indices, np.uint32).tobytes(), "VEC3_U32")}
            if g.uvs is not None:
                m["texcoords"] = add_view(np.ascontiguousarray(g.uvs, np.float32).tobytes(), "VEC2_F32")
            geom_mesh_id[(mi, gi)] = len(meshes)
            meshes.append(m)

    images = []
    for t in scene.textures:
        img = np.ascontiguousarray(np.asarray(t.img, np.uint8)[::-1])  # the loader flips vertically
        buf = io.BytesIO()
        mode = {1: "L", 3: "RGB", 4: "RGBA"}[img.shape[2]]
        PILImage.fromarray(img[:, :, 0] if mode == "L" else img, mode).save(buf, format="PNG")
        images.append({"name": t.name, "view": add_view(buf.getvalue(), "UINT_8"),
                       "color_space": "LINEAR" if t.color_space == LINEAR else "SRGB"})

    names = [("metallic", "metallic"), ("specular", "specular"), ("roughness", "roughness"),
             ("specular_tint", "specular_tint"), ("anisotropic", "anisotropy"), ("sheen", "sheen"),
             ("sheen_tint", "sheen_tint"), ("clearcoat", "clearcoat"), ("clearcoat_roughness", "clearcoat_gloss"),
             ("ior", "ior"), ("transmission", "specular_transmission")]
    materials = []
    for m in scene.materials:
        jm = {}
        h = _handle(m.base_color[0])
        jm["base_color"] = [0.0 if h is not None else _num(m.base_color[0]), _num(m.base_color[1]), _num(m.base_color[2])]
        if h is not None:
            jm["base_color_texture"] = h[0]
        for key, attr in names:
            v = getattr(m, attr)
            h = _handle(v)
            jm[key] = 0.0 if h is not None else _num(v)
            if h is not None:
                jm[key + "_texture"] = {"texture": h[0], "channel": h[1]}
        materials.append(jm)

    objects = []
    for inst in scene.instances:
        pm = scene.parameterized_meshes[inst.parameterized_mesh_id]
        mat = np.asarray(inst.transform, np.float32).reshape(4, 4)
        for gi in range(len(scene.meshes[pm.mesh_id].geometries)):
            objects.append({"type": "MESH", "mesh": geom_mesh_id[(pm.mesh_id, gi)], "material": int(pm.material_ids[gi]),
                            "matrix": [_num(x) for x in mat.T.reshape(-1)]})
    for l in scene.lights:
        e = np.array(l.emission[:3], np.float32)
        mat = np.zeros((4, 4), np.float32)
        mat[:3, 0] = l.v_x[:3]
        mat[:3, 1] = l.v_y[:3]
        mat[:3, 2] = -np.array(l.normal[:3], np.float32)
        mat[:3, 3] = l.position[:3]
        mat[3, 3] = 1.0
        objects.append({"type": "LIGHT", "color": [_num(x) for x in e], "energy": 1.0,
                        "size": [_num(l.width), _num(l.height)], "matrix": [_num(x) for x in mat.T.reshape(-1)]})

    objects.extend(extra_objects)
    header = {"meshes": meshes, "images": images, "materials": materials, "objects": objects, "buffer_views": views}
    js = json.dumps(header).encode()
    js += b" " * ((-(len(js) + 8)) % 8)  # keep the data block 8-byte aligned
    if not align:
        js += b" "
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(js)))
        f.write(js)
        f.write(bytes(data))
    return path


def crts_scene_view(scene: Scene) -> Scene:
    """The Scene the reference's load_crts builds from ``write_crts(scene)``: one mesh per geometry, one
    parameterized mesh per distinct (mesh, material) pair in order of first use, one instance per (instance,
    geometry), RGBA textures, light emission.w = 1 and position.w = 1. Rendering it is bit-identical to rendering
    the file through the loader (and, geometry order being preserved, to rendering ``scene`` itself whenever
    every instance's geometries keep their order — flattened primitive ids follow instance order)."""
    from .scene import Image, Instance, Mesh, ParameterizedMesh, QuadLight

    meshes, gid = [], {}
    for mi, mesh in enumerate(scene.meshes):
        for gi, g in enumerate(mesh.geometries):
            gid[(mi, gi)] = len(meshes)
            meshes.append(Mesh([g]))
    pms, pm_ids, instances = [], {}, []
    for inst in scene.instances:
        pm = scene.parameterized_meshes[inst.parameterized_mesh_id]
        for gi in range(len(scene.meshes[pm.mesh_id].geometries)):
            key = (gid[(pm.mesh_id, gi)], int(pm.material_ids[gi]))
            if key not in pm_ids:
                pm_ids[key] = len(pms)
                pms.append(ParameterizedMesh(key[0], [key[1]]))
            instances.append(Instance(np.asarray(inst.transform, np.float32), pm_ids[key]))
    textures = []
    for t in scene.textures:
        img = np.asarray(t.img, np.uint8)
        if img.shape[2] != 4:
            rgb = img if img.shape[2] == 3 else np.repeat(img[:, :, :1], 3, axis=2)
            img = np.concatenate([rgb, np.full(img.shape[:2] + (1,), 255, np.uint8)], axis=2)
        textures.append(Image(t.name, np.ascontiguousarray(img), t.color_space))
    lights = [QuadLight(emission=(l.emission[0], l.emission[1], l.emission[2], 1.0),
                        position=(l.position[0], l.position[1], l.position[2], 1.0), normal=l.normal, v_x=l.v_x,
                        width=l.width, v_y=l.v_y, height=l.height) for l in scene.lights]
    return Scene(meshes=meshes, parameterized_meshes=pms, instances=instances, materials=list(scene.materials),
                 textures=textures, lights=lights, samples_per_pixel=scene.samples_per_pixel)
