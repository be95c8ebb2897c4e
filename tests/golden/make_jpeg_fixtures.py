"""Writes tests/golden/jpeg_*.jpg and jpeg_expected.npz: small JPEG files (PIL's encoder) and the RGBA pixels the REFERENCE'S
stb_image decodes from them (through oracle/_ref/libcrt_refscene.so, i.e. Scene::load_obj -> Image -> stbi_load with the
vertical flip). Run in the build container (needs /root/reference built by oracle/ref_build/Makefile); the GPU box only reads
the committed files (tests/test_scene_io.py::test_native_jpeg_decoder_against_committed_stb_output)."""
import os
import sys

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import test_scene_io as t  # noqa: E402

y, x = np.mgrid[0:37, 0:61]
ramp = np.stack([(x * 3) % 256, (y * 5) % 256, (x ^ y) * 4 % 256], 2).astype(np.uint8)
smooth = np.stack([128 + 100 * np.sin(x / 9.0), 128 + 90 * np.cos(y / 7.0), 128 + 80 * np.sin((x + y) / 11.0)], 2).astype(np.uint8)
files = {"jpeg_baseline_420.jpg": (ramp, "RGB", dict(quality=80, subsampling=2)),
         "jpeg_progressive_422.jpg": (smooth, "RGB", dict(quality=88, subsampling=1, progressive=True, optimize=True)),
         "jpeg_restart_444.jpg": (smooth, "RGB", dict(quality=70, subsampling=0, restart_marker_blocks=5)),
         "jpeg_grey.jpg": (smooth[:, :, 1], "L", dict(quality=75))}
for name, (img, mode, opts) in files.items():
    Image.fromarray(img, mode).save(os.path.join(HERE, name), format="JPEG", **opts)
names = sorted(files)
with open(os.path.join(HERE, "jpeg_fixture.mtl"), "w") as f:
    f.write("".join(f"newmtl m{i}\nKd 1 1 1\nmap_Kd {n}\n" for i, n in enumerate(names)))
with open(os.path.join(HERE, "jpeg_fixture.obj"), "w") as f:
    f.write("mtllib jpeg_fixture.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\n" +
            "".join(f"g g{i}\nusemtl m{i}\nf 1/1 2/2 3/3\n" for i in range(len(names))))
ref = t._reference_arrays(os.path.join(HERE, "jpeg_fixture.obj"))
np.savez_compressed(os.path.join(HERE, "jpeg_expected.npz"), **{n: px for n, (px, _) in zip(names, ref["textures"])})
print({n: px.shape for n, (px, _) in zip(names, ref["textures"])})
