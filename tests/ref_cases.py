"""The frame cases and known-answer inputs shared by tests/golden/make_ref_embree_golden.py (which
runs them through the reference's own Embree backend) and tests/test_reference_embree.py (which runs
them through the oracle, and on the GPU box through the CUDA backend)."""
import numpy as np

from helpers import synthetic_material_scene

# name -> (generator, kwargs, width, height, frames, max_depth, spp)
FRAME_CASES = {
    "cornell": ("cornell_box", {}, 96, 64, 2, 5, 2),
    "cornell_d8": ("cornell_box", {}, 64, 64, 1, 8, 1),
    "materials": ("materials", {}, 96, 72, 2, 5, 2),
    "sponza_like": ("sponza_like", dict(detail=0.25, tex_size=64), 128, 72, 1, 5, 1),
    "sponza_like_d8": ("sponza_like", dict(detail=0.25, tex_size=64), 96, 54, 2, 8, 2),
    "san_miguel_like_instances": ("san_miguel_like", dict(scale=0.02, tex_size=64), 96, 54, 1, 5, 2),
    "rungholt_like": ("rungholt_like", dict(scale=0.001), 96, 54, 1, 5, 1),
    "ragged_70x50": ("cornell_box", {}, 70, 50, 3, 5, 1),
}


def make_case(name):
    from chameleonrt_b200 import ArcballCamera, scenes

    gen, kwargs, w, h, frames, depth, spp = FRAME_CASES[name]
    if gen == "materials":
        scene, cam = synthetic_material_scene(spp=spp)
    else:
        scene, cam = getattr(scenes, gen)(spp=spp, **kwargs)
    c = ArcballCamera(cam["eye"], cam["center"], cam["up"])
    return scene, (c.eye(), c.dir(), c.up(), cam["fov_y"]), w, h, frames, depth


def kat_inputs():
    rng = np.random.default_rng(20260923)
    mats = []
    for metallic in (0.0, 0.6):
        for rough in (0.02, 0.4, 1.0):
            for aniso in (0.0, 0.7):
                for trans in (0.0, 0.85):
                    mats.append([0.75, 0.4, 0.25, metallic, 0.55, rough, 0.35, aniso, 0.45, 0.5, 0.65, 0.7, 1.5, trans, 0, 0])
    n = np.array([0.25, 0.85, -0.35], np.float32)
    n /= np.linalg.norm(n)
    dirs = rng.normal(size=(20, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    light = np.array([5, 5, 5, 0, 0.0, 1.98, 0.0, 0, 0, -1, 0, 0, 1, 0, 0, 0.5, 0, 0, 1, 0.4], np.float32)
    out = dict(
        mats=np.array(mats, np.float32), n=n, dirs=dirs.astype(np.float32),
        seeds=rng.integers(1, 2**32 - 1, size=6, dtype=np.uint64).astype(np.uint32),
        light=light, light_s=rng.random((5, 2)).astype(np.float32), light_orig=np.array([0.1, 0.3, -0.2], np.float32),
        uv=(rng.random((64, 2)) * 3.0 - 1.0).astype(np.float32),
        miss_dirs=np.concatenate([dirs, np.array([[0, 1, 0], [0, -1, 0], [1, 0, 0], [0, 0, -1], [0, 0, 1]], np.float32)]),
        rng_keys=np.array([(0, 1), (1, 1), (921599, 7), (12345, 4 * 19 + 3)], np.uint32),
    )
    for ch in (1, 3, 4):
        out[f"tex{ch}"] = rng.integers(0, 256, size=(7, 5, ch), dtype=np.uint8)
    return out
