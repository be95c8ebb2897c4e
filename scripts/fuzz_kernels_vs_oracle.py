"""Randomised comparison of the product's KERNEL SOURCE run on the host (libcrt_wavefront_hostcheck*.so,
tests/test_wavefront_host.py) with the CPU oracle (which the other fuzzer keeps bit-identical to the reference's
own backend): random scenes as in fuzz_oracle_vs_reference.py — every Disney parameter, textured scalars,
transmission, several lights, geometry with and without texture coordinates — with identity instance transforms
(the product flattens instances to world space, so sheared instances only agree statistically). With the
reference's Schlick formula and one sample per pixel the frames must be bit-identical, ray counts included; with
the product's arithmetic all pixels must be within the parity tolerance and the ray counts equal up to NaN paths.
    python scripts/fuzz_kernels_vs_oracle.py [n] [first_seed]"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "scripts")):
    sys.path.insert(0, p)

import fuzz_oracle_vs_reference as base  # noqa: E402
from chameleonrt_b200 import ArcballCamera  # noqa: E402
from helpers import parity  # noqa: E402
from oracle import OracleBackend  # noqa: E402
from test_wavefront_host import HostWavefront, _load  # noqa: E402

LIBS = None


def one(seed):
    global LIBS
    if LIBS is None:
        LIBS = (_load("libcrt_wavefront_hostcheck.so"), _load("libcrt_wavefront_hostcheck_powf.so"))
    rng = np.random.default_rng(seed)
    scene = base.random_scene(rng)
    for inst in scene.instances:
        inst.transform = np.eye(4, dtype=np.float32)
    w, h, depth, frames = int(rng.integers(8, 90)), int(rng.integers(8, 80)), int(rng.integers(1, 9)), int(rng.integers(1, 4))
    spp1 = rng.random() < 0.6
    if spp1:
        scene.samples_per_pixel = 1
    cam = ArcballCamera(tuple(rng.uniform(-9, 9, 3)), tuple(rng.uniform(-1, 1, 3)), (0.0, 1.0, 0.0))
    view = (cam.eye(), cam.dir(), cam.up(), float(rng.uniform(20, 90)))
    cpu = OracleBackend(max_depth=depth)
    cpu.initialize(w, h)
    cpu.set_scene(scene)
    prod, powf = HostWavefront(LIBS[0], scene, w, h, depth), HostWavefront(LIBS[1], scene, w, h, depth)
    for f in range(frames):
        so = cpu.render(*view, f == 0, True)
        rp, rf = prod.render(view, f == 0), powf.render(view, f == 0)
    want = cpu.read_accum()
    a_prod, _ = prod.read()
    a_powf, i_powf = powf.read()
    has_nan = bool(np.isnan(want).any())
    frac, rel_l1 = parity(a_prod, want)
    ok = frac >= 0.999 and rel_l1 <= 1e-4 and (has_nan or rp == so.num_rays)
    exact = None
    if spp1:
        same = (a_powf.view(np.uint32) == want.view(np.uint32)) | (np.isnan(a_powf) & np.isnan(want))
        exact = bool(same.all()) and (has_nan or rf == so.num_rays) and (has_nan or np.array_equal(i_powf, cpu.img))
        ok = ok and exact
    return ok, (frac, rel_l1, rp, so.num_rays, exact, has_nan, scene.total_tris(), w, h, depth, frames, scene.samples_per_pixel)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = nan = exact = 0
    for seed in range(first, first + n):
        ok, info = one(seed)
        nan += info[5]
        exact += 1 if info[4] else 0
        if not ok:
            bad += 1
            print("MISMATCH seed", seed, info)
    print(f"{n} random scenes, {bad} mismatches; {exact} one-sample scenes bit-identical with the reference's Schlick formula; "
          f"{nan} scenes with NaN pixels")
