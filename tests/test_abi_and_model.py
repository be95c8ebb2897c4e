"""The C-ABI library loads and exports every symbol include/crt_cuda.h declares (no compute
calls without a GPU), and the Python mirror of the reference's scene model marshals correctly."""
import ctypes as C
import os
import re
import struct

import numpy as np
import pytest

from chameleonrt_b200 import backend, tiles
from chameleonrt_b200.scene import (CGeometry, CImage, CInstance, CMaterial, CQuadLight, CRenderStats, CScene,
                                    DisneyMaterial, default_obj_light, textured_param)
from chameleonrt_b200.scenes import cornell_box, sponza_like

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(built):
    header = open(os.path.join(ROOT, "include", "crt_cuda.h")).read()
    declared = set(re.findall(r"\b(crtc_[a-z_]+)\s*\(", header))
    assert declared == set(backend.C_ABI_SYMBOLS), declared ^ set(backend.C_ABI_SYMBOLS)
    lib = C.CDLL(backend.lib_path())
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} not exported by libcrt_cuda_core.so"


def test_scene_io_header_symbols_exported(built):
    """include/crt_scene_io.h (the native scene loader): every declared entry point is exported by libcrt_scene_io.so."""
    from chameleonrt_b200 import scene_io

    header = open(os.path.join(ROOT, "include", "crt_scene_io.h")).read()
    declared = set(re.findall(r"\b(crtio_[a-z_]+)\s*\(", header))
    assert declared == {"crtio_load_obj", "crtio_load_crts", "crtio_load_gltf", "crtio_load", "crtio_load_mode", "crtio_cameras", "crtio_texture_name", "crtio_scene_view", "crtio_timings",
                        "crtio_warnings", "crtio_free", "crtio_last_error"}
    lib = C.CDLL(scene_io.lib_path())
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} not exported by libcrt_scene_io.so"


def test_no_gpu_fails_loudly(built):
    """On a machine without a CUDA device creation must fail with a message, never fall back."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="CUDA"):
        backend.RenderCUDA(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chameleonrt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) and f != "hostcheck.cpp":
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_struct_layouts_match_header():
    assert C.sizeof(CMaterial) == 64 and C.sizeof(CQuadLight) == 80  # util/material.h:29-46, util/lights.h:6-18
    assert C.sizeof(CInstance) == 68 and C.sizeof(CRenderStats) == 16
    assert C.sizeof(CGeometry) == 32 and C.sizeof(CImage) == 24 and C.sizeof(CScene) == 80


def test_textured_param_encoding():
    x = textured_param(5, 2)
    bits = struct.unpack("<I", struct.pack("<f", x))[0]
    assert bits & 0x80000000 and (bits >> 29) & 3 == 2 and bits & 0x1FFFFFFF == 5  # texture_channel_mask.h:16-23


def test_scene_marshalling_and_obj_conventions():
    scene, _ = cornell_box(spp=3)
    assert len(scene.meshes) == 1 and len(scene.instances) == 1 and len(scene.parameterized_meshes) == 1
    assert scene.total_tris() == scene.unique_tris() == 34
    ms = scene.to_c()
    c = ms.c
    assert c.num_meshes == 1 and c.samples_per_pixel == 3 and c.num_lights == 1
    g0 = c.meshes[0].geometries[0]
    assert g0.num_tris == scene.meshes[0].geometries[0].num_tris()
    assert [c.instances[0].transform[i] for i in range(16)] == [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    # load_obj material mapping (scene.cpp:191-195) with tinyobj's default Ns = 1
    m = scene.materials[0]
    assert m.specular == pytest.approx(0.002) and m.roughness == pytest.approx(0.998) and m.specular_transmission == 0
    # the generated light (scene.cpp:218-227)
    l = default_obj_light()
    n = np.array(l.normal)
    assert np.allclose(n, np.array([0.5, -0.8, -0.5]) / np.linalg.norm([0.5, -0.8, -0.5]), atol=1e-6)
    assert np.allclose(np.array(l.position[:3]), -10 * n, atol=1e-5) and l.width == 5 and l.height == 5
    assert abs(np.dot(l.v_x, l.v_y)) < 1e-6 and abs(np.dot(l.v_x, n)) < 1e-6


def test_validate_materials_default():
    scene, _ = cornell_box()
    scene.parameterized_meshes[0].material_ids[1] = -1
    n = len(scene.materials)
    scene.validate_materials()
    assert len(scene.materials) == n + 1 and scene.parameterized_meshes[0].material_ids[1] == n
    d = scene.materials[-1]
    assert d.base_color == (0.9, 0.9, 0.9) and d.roughness == 1.0  # util/material.h:29-46


def test_sponza_like_is_deterministic_and_sized():
    a, _ = sponza_like(tex_size=16)
    b, _ = sponza_like(tex_size=16)
    assert a.total_tris() == b.total_tris() and 255_000 < a.total_tris() < 270_000  # Crytek Sponza: 262,267
    assert 20 <= len(a.materials) <= 30 and len(a.textures) == 8
    for ga, gb in zip(a.meshes[0].geometries, b.meshes[0].geometries):
        assert (ga.vertices == gb.vertices).all() and (ga.indices == gb.indices).all()


@pytest.mark.parametrize("w,h,world", [(1280, 720, 1), (1280, 720, 8), (100, 70, 3), (64, 64, 2), (129, 65, 4)])
def test_tile_layout_roundtrip(w, h, world):
    """tiles.py must be a bijection between the frame and the union of the ranks' local buffers."""
    rng = np.random.default_rng(0)
    full = rng.random((h, w, 3)).astype(np.float32)
    chunks = [tiles.to_local(full, r, world) for r in range(world)]
    ntx, nty = tiles.num_tiles(w, h)
    assert sum(len(c) for c in chunks) == ntx * nty * 4096
    assert (tiles.assemble(chunks, w, h, world) == full).all()
    seen = np.zeros((h, w), np.int32)
    for r in range(world):
        x, y, valid = tiles.local_pixel_coords(w, h, r, world)
        np.add.at(seen, (y[valid], x[valid]), 1)
        # a warp (32 consecutive local pixels) covers one 8x4 rectangle
        if len(x):
            assert x[:32].max() - x[:32].min() == 7 and y[:32].max() - y[:32].min() == 3
    assert (seen == 1).all()
