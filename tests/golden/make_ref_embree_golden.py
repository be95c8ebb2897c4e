"""Generates tests/golden/ref_embree_frames.npz from the REFERENCE'S OWN Embree/ISPC backend.

oracle/_ref/libcrt_embree.so is /root/reference/backends/embree/{render_embree.cpp,embree_utils.cpp,
render_embree.ispc,*.ih} compiled where they lie (oracle/ref_build/Makefile: ISPC kernels as scalar
C++, Embree/TBB/GLM replaced by third_party/ stand-ins). It exists only where /root/reference does, so
its outputs are frozen here: float framebuffers, per-pixel ray counts and sRGB8 images of small
frames over every scene class, plus known-answer tables of the reference's pure functions. The
oracle must reproduce all of them bit for bit (tests/test_reference_embree.py), which pins it.
Run from the repo root:  python tests/golden/make_ref_embree_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ref_cases import FRAME_CASES, kat_inputs, make_case  # noqa: E402
from oracle.ref_embree import RefEmbreeBackend, load_ref_embree_lib  # noqa: E402


def main():
    out = {}
    for name in FRAME_CASES:
        scene, view, w, h, frames, depth = make_case(name)
        ref = RefEmbreeBackend(max_depth=depth)
        ref.initialize(w, h)
        ref.set_scene(scene)
        rays = []
        for f in range(frames):
            st = ref.render(*view, f == 0, True)
            rays.append(st.num_rays)
        out[f"{name}.accum"] = ref.read_accum()
        out[f"{name}.ray_stats"] = ref.read_ray_stats()
        out[f"{name}.img"] = ref.img.copy()
        out[f"{name}.rays"] = np.array(rays, np.uint64)
        print(name, w, h, "frames", frames, "depth", depth, "rays", rays)

    lib = load_ref_embree_lib()
    k = kat_inputs()
    ev = np.zeros((len(k["mats"]), len(k["dirs"]), len(k["dirs"]), 4), np.float32)
    for mi, m in enumerate(k["mats"]):
        for oi, wo in enumerate(k["dirs"]):
            for ii, wi in enumerate(k["dirs"]):
                lib.refispc_kat_disney_eval(m.ctypes.data, k["n"].ctypes.data, wo.ctypes.data, wi.ctypes.data,
                                            ev[mi, oi, ii].ctypes.data)
    out["kat.disney_eval"] = ev
    sm = np.zeros((len(k["mats"]), len(k["dirs"]), len(k["seeds"]), 8), np.float32)
    for mi, m in enumerate(k["mats"]):
        for oi, wo in enumerate(k["dirs"]):
            for si, seed in enumerate(k["seeds"]):
                st = C.c_uint32(int(seed))
                lib.refispc_kat_disney_sample(m.ctypes.data, k["n"].ctypes.data, wo.ctypes.data, C.addressof(st),
                                              sm[mi, oi, si].ctypes.data)
                sm[mi, oi, si, 7] = np.array([st.value], np.uint32).view(np.float32)[0]
    out["kat.disney_sample"] = sm
    lt = np.zeros((len(k["light_s"]), len(k["dirs"]), 9), np.float32)
    for si, s2 in enumerate(k["light_s"]):
        for di, d in enumerate(k["dirs"]):
            lib.refispc_kat_light(k["light"].ctypes.data, s2.ctypes.data, k["light_orig"].ctypes.data, d.ctypes.data,
                                  lt[si, di].ctypes.data)
    out["kat.light"] = lt
    for ch in (1, 3, 4):
        tex = k[f"tex{ch}"]
        tx = np.zeros((len(k["uv"]), 4), np.float32)
        lib.refispc_kat_texture(tex.ctypes.data, tex.shape[1], tex.shape[0], ch, k["uv"].ctypes.data, len(k["uv"]),
                                tx.ctypes.data)
        out[f"kat.texture{ch}"] = tx
    ms = np.zeros((len(k["miss_dirs"]), 3), np.float32)
    lib.refispc_kat_miss(k["miss_dirs"].ctypes.data, len(k["miss_dirs"]), ms.ctypes.data)
    out["kat.miss"] = ms
    ob = np.zeros((len(k["dirs"]), 6), np.float32)
    for di, d in enumerate(k["dirs"]):
        lib.refispc_kat_ortho_basis(d.ctypes.data, ob[di].ctypes.data)
    out["kat.ortho_basis"] = ob
    rs = np.zeros((len(k["rng_keys"]), 16), np.uint32)
    rf = np.zeros((len(k["rng_keys"]), 16), np.float32)
    for i, (pix, frame) in enumerate(k["rng_keys"]):
        lib.refispc_kat_rng(int(pix), int(frame), 16, rs[i].ctypes.data, rf[i].ctypes.data)
    out["kat.rng_states"], out["kat.rng_floats"] = rs, rf
    path = os.path.join(ROOT, "tests", "golden", "ref_embree_frames.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
