"""Writes a glTF-class ``Scene`` as .gltf + .bin + PNG images, in exactly the subset the reference's
``Scene::load_gltf`` reads (util/scene.cpp:230-415): one glTF mesh per ChameleonRT *parameterized mesh*
(primitives = geometries, POSITION / TEXCOORD_0 / uint32 indices, one material each), pbrMetallicRoughness
materials (baseColorFactor, metallicFactor, roughnessFactor, baseColorTexture -> sRGB,
metallicRoughnessTexture -> linear, B = metallic, G = roughness), a flat list of nodes with a ``matrix`` each
(= instances), no lights (the loader generates its default quad light, scene.cpp:404-414).

Used to push the synthetic San-Miguel-like scene through the reference's own loader
(oracle/_ref/crt_headless <backend> scene.gltf) — tests/test_reference_embree.py, tests/test_reference_plugin.py.
All numbers are float32 values written with enough digits to round-trip, geometry goes through the binary
buffer, so the loaded scene is bit-identical to the in-memory one (textures come back as RGBA: tinygltf asks
stb_image for 4 components).
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np

from .scene import SRGB, Scene, f32_bits

_TEXTURED = 0x80000000


def _handle(v: float):
    """(texture id, channel) if the float is a texture handle (util/texture_channel_mask.h), else None."""
    bits = f32_bits(v)
    if bits & _TEXTURED:
        return bits & 0x1FFFFFFF, (bits >> 29) & 0x3
    return None


def _num(v: float) -> float:
    """A Python float that is exactly the float32 value (json writes the shortest round-trip repr)."""
    return float(np.float32(v))


def write_gltf(scene: Scene, path: str) -> str:
    from PIL import Image as PILImage

    out_dir = os.path.dirname(os.path.abspath(path))
    os.makedirs(out_dir, exist_ok=True)
    base = os.path.splitext(os.path.basename(path))[0]
    blob = bytearray()
    views, accessors = [], []

    def add_accessor(arr: np.ndarray, comp: int, typ: str, target: int, with_bounds: bool = False) -> int:
        while len(blob) % 4:
            blob.append(0)
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": arr.nbytes, "target": target})
        blob.extend(arr.tobytes())
        acc = {"bufferView": len(views) - 1, "componentType": comp, "count": int(arr.shape[0]), "type": typ}
        if with_bounds:
            acc["min"] = [_num(x) for x in arr.min(axis=0)]
            acc["max"] = [_num(x) for x in arr.max(axis=0)]
        accessors.append(acc)
        return len(accessors) - 1

    # geometry arrays are shared by every parameterized mesh that uses the same Mesh
    geom_acc = {}
    for mi, mesh in enumerate(scene.meshes):
        for gi, g in enumerate(mesh.geometries):
            v = np.ascontiguousarray(g.vertices, np.float32).reshape(-1, 3)
            idx = np.ascontiguousarray(g.indices, np.uint32).reshape(-1)
            entry = {"POSITION": add_accessor(v, 5126, "VEC3", 34962, True), "indices": add_accessor(idx, 5125, "SCALAR", 34963)}
            if g.uvs is not None:
                entry["TEXCOORD_0"] = add_accessor(np.ascontiguousarray(g.uvs, np.float32).reshape(-1, 2), 5126, "VEC2", 34962)
            geom_acc[(mi, gi)] = entry

    images, textures = [], []
    for ti, t in enumerate(scene.textures):
        fname = f"{base}_tex{ti}.png"
        img = np.ascontiguousarray(t.img, np.uint8)
        mode = {1: "L", 3: "RGB", 4: "RGBA"}[img.shape[2]]
        PILImage.fromarray(img[:, :, 0] if mode == "L" else img, mode).save(os.path.join(out_dir, fname))
        images.append({"uri": fname, "name": t.name})
        textures.append({"source": ti})

    materials = []
    for m in scene.materials:
        pbr = {}
        h = _handle(m.base_color[0])
        if h is not None:
            pbr["baseColorTexture"] = {"index": h[0]}
            assert scene.textures[h[0]].color_space == SRGB, "load_gltf marks base colour textures sRGB"
            pbr["baseColorFactor"] = [1.0, _num(m.base_color[1]), _num(m.base_color[2]), 1.0]
        else:
            pbr["baseColorFactor"] = [_num(m.base_color[0]), _num(m.base_color[1]), _num(m.base_color[2]), 1.0]
        hm, hr = _handle(m.metallic), _handle(m.roughness)
        if hm is not None or hr is not None:
            assert hm is not None and hr is not None and hm[0] == hr[0] and hm[1] == 2 and hr[1] == 1, \
                "glTF packs metallic (B) and roughness (G) into one texture"
            pbr["metallicRoughnessTexture"] = {"index": hm[0]}
        else:
            pbr["metallicFactor"] = _num(m.metallic)
            pbr["roughnessFactor"] = _num(m.roughness)
        materials.append({"pbrMetallicRoughness": pbr})

    meshes = []
    for pm in scene.parameterized_meshes:
        prims = []
        for gi in range(len(scene.meshes[pm.mesh_id].geometries)):
            e = geom_acc[(pm.mesh_id, gi)]
            attrs = {"POSITION": e["POSITION"]}
            if "TEXCOORD_0" in e:
                attrs["TEXCOORD_0"] = e["TEXCOORD_0"]
            prims.append({"attributes": attrs, "indices": e["indices"], "material": int(pm.material_ids[gi]), "mode": 4})
        meshes.append({"primitives": prims})

    nodes = []
    for inst in scene.instances:
        m = np.asarray(inst.transform, np.float32).reshape(4, 4)
        nodes.append({"mesh": int(inst.parameterized_mesh_id), "matrix": [_num(x) for x in m.T.reshape(-1)]})  # column-major

    doc = {
        "asset": {"version": "2.0", "generator": "chameleonrt_b200.gltf_io"},
        "scene": 0,
        "scenes": [{"nodes": list(range(len(nodes)))}],
        "nodes": nodes,
        "meshes": meshes,
        "materials": materials,
        "textures": textures,
        "images": images,
        "accessors": accessors,
        "bufferViews": views,
        "buffers": [{"uri": base + ".bin", "byteLength": len(blob)}],
    }
    if not textures:
        del doc["textures"], doc["images"]
    with open(os.path.join(out_dir, base + ".bin"), "wb") as f:
        f.write(bytes(blob))
    with open(path, "w") as f:
        json.dump(doc, f)
    return path
