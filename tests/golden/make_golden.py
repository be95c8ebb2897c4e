"""Generates tests/golden/oracle_kat.npz from the CPU oracle (strict build).

The reference ships no golden vectors and cannot be built here, so these
known-answer vectors are produced by the oracle itself and frozen: they pin the oracle against
regressions and give the GPU tests fixed targets that do not need the oracle at run time.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from chameleonrt_b200 import ArcballCamera  # noqa: E402
from chameleonrt_b200.scene import default_obj_light, textured_param  # noqa: E402
from chameleonrt_b200.scenes import cornell_box, make_texture  # noqa: E402
from oracle import OracleBackend, load_oracle_lib  # noqa: E402
from oracle.oracle import primary_rays  # noqa: E402


def main():
    lib = load_oracle_lib()
    out = {}
    rng = np.random.default_rng(20260922)
    # RNG streams (SURVEY §8c list)
    for k, (pix, frame) in enumerate([(0, 1), (1, 1), (921599, 7)]):
        states = np.zeros(16, np.uint32)
        floats = np.zeros(16, np.float32)
        lib.oracle_kat_rng(pix, frame, 16, states.ctypes.data, floats.ctypes.data)
        out[f"rng_states_{k}"] = states
        out[f"rng_floats_{k}"] = floats
    out["rng_keys"] = np.array([(0, 1), (1, 1), (921599, 7)], np.uint32)
    # camera basis
    cam = ArcballCamera((0.0, 1.0, 3.4), (0.0, 1.0, 0.0), (0.0, 1.0, 0.0))
    basis = np.zeros(12, np.float32)
    import ctypes as C
    fp = C.POINTER(C.c_float)
    e, d, u = (np.ascontiguousarray(x, np.float32) for x in (cam.eye(), cam.dir(), cam.up()))
    lib.oracle_kat_camera(e.ctypes.data_as(fp), d.ctypes.data_as(fp), u.ctypes.data_as(fp), C.c_float(40.0), 512, 512,
                          basis.ctypes.data)
    out["camera_in"] = np.concatenate([e, d, u, [40.0, 512, 512]]).astype(np.float32)
    out["camera_basis"] = basis
    # Disney BSDF tables: materials x directions
    mats = []
    for metallic in (0.0, 0.7):
        for rough in (0.05, 0.5, 1.0):
            for aniso in (0.0, 0.6):
                for trans in (0.0, 0.8):
                    mats.append([0.8, 0.45, 0.2, metallic, 0.5, rough, 0.3, aniso, 0.4, 0.5, 0.6, 0.7, 1.45, trans, 0, 0])
    mats = np.array(mats, np.float32)
    n = np.array([0.2, 0.9, 0.3], np.float32)
    n /= np.linalg.norm(n)
    dirs = rng.normal(size=(24, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    ev = np.zeros((len(mats), len(dirs), len(dirs), 4), np.float32)
    for mi, m in enumerate(mats):
        for oi, wo in enumerate(dirs):
            for ii, wi in enumerate(dirs):
                o4 = np.zeros(4, np.float32)
                lib.oracle_kat_disney_eval(m.ctypes.data, n.ctypes.data, np.ascontiguousarray(wo).ctypes.data,
                                           np.ascontiguousarray(wi).ctypes.data, o4.ctypes.data)
                ev[mi, oi, ii] = o4
    out["bsdf_mats"] = mats
    out["bsdf_n"] = n
    out["bsdf_dirs"] = dirs
    out["bsdf_eval"] = ev
    smp = np.zeros((len(mats), 8, 16, 7), np.float32)
    smp_state = np.zeros((len(mats), 8, 16), np.uint32)
    for mi, m in enumerate(mats):
        for oi in range(8):
            wo = np.ascontiguousarray(dirs[oi])
            st = np.array([12345 + 977 * mi + oi], np.uint32)
            for k in range(16):
                o7 = np.zeros(7, np.float32)
                lib.oracle_kat_disney_sample(m.ctypes.data, n.ctypes.data, wo.ctypes.data, st.ctypes.data, o7.ctypes.data)
                smp[mi, oi, k] = o7
                smp_state[mi, oi, k] = st[0]
    out["bsdf_sample"] = smp
    out["bsdf_sample_state"] = smp_state
    # light
    light = np.array(default_obj_light().as_floats(), np.float32)
    ls = rng.random((16, 2)).astype(np.float32)
    lo = (rng.normal(size=(16, 3)) * 3).astype(np.float32)
    ld = rng.normal(size=(16, 3)).astype(np.float32)
    ld /= np.linalg.norm(ld, axis=1, keepdims=True)
    # aim half of the rays at the light so quad_intersect hits
    for i in range(0, 16, 2):
        tgt = light[4:7] + 0.5 * light[12:15] + 0.3 * light[16:19]
        v = tgt - lo[i]
        ld[i] = (v / np.linalg.norm(v)).astype(np.float32)
    lres = np.zeros((16, 9), np.float32)
    for i in range(16):
        lib.oracle_kat_light(light.ctypes.data, np.ascontiguousarray(ls[i]).ctypes.data, np.ascontiguousarray(lo[i]).ctypes.data,
                             np.ascontiguousarray(ld[i]).ctypes.data, lres[i].ctypes.data)
    out.update(light=light, light_s=ls, light_o=lo, light_d=ld, light_res=lres)
    # texture (incl. negative uv and the truncation quirk region)
    tex = make_texture("bricks", 7, 16)
    uv = np.concatenate([rng.uniform(-2, 2, (24, 2)), [[0.0, 0.0], [0.01, 0.01], [-0.01, -0.01], [1.0, 1.0], [0.999, 0.5]]]).astype(np.float32)
    tres = np.zeros((len(uv), 4), np.float32)
    lib.oracle_kat_texture(tex.ctypes.data, 16, 16, 4, uv.ctypes.data, len(uv), tres.ctypes.data)
    out.update(tex_data=tex, tex_uv=uv, tex_res=tres)
    # miss shader, ortho basis, srgb8, triangle test
    md = rng.normal(size=(32, 3)).astype(np.float32)
    md /= np.linalg.norm(md, axis=1, keepdims=True)
    mres = np.zeros((32, 3), np.float32)
    lib.oracle_kat_miss(md.ctypes.data, 32, mres.ctypes.data)
    out.update(miss_dirs=md, miss_res=mres)
    ob = np.zeros((32, 6), np.float32)
    for i in range(32):
        lib.oracle_kat_ortho_basis(np.ascontiguousarray(md[i]).ctypes.data, ob[i].ctypes.data)
    out["ortho_res"] = ob
    sx = np.concatenate([np.linspace(-0.1, 1.2, 64), [0.0031308, 0.5, 1.0, np.nan]]).astype(np.float32)
    sres = np.zeros(len(sx), np.uint8)
    lib.oracle_kat_srgb8(sx.ctypes.data, len(sx), sres.ctypes.data)
    out.update(srgb_in=sx, srgb_res=sres)
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32)
    trays = np.zeros((12, 8), np.float32)
    trays[:, :3] = rng.uniform(-0.2, 1.2, (12, 3))
    trays[:, 2] = 1.0
    trays[:, 4:7] = [0, 0, -1]
    trays[:, 7] = 1e20
    trays[10, 4:7] = [1, 0, 0]  # parallel to the triangle: det == 0
    trays[11, :3] = [0.25, 0.25, 1.0]
    trays[11, 7] = 0.5  # tfar in front of the triangle
    tout = np.zeros((12, 4), np.float32)
    for i in range(12):
        lib.oracle_kat_tri(tri.ctypes.data, trays[i].ctypes.data, tout[i].ctypes.data)
    out.update(tri=tri, tri_rays=trays, tri_res=tout)

    # Cornell primary-visibility AOVs + a small accumulated float framebuffer
    scene, c = cornell_box(spp=2)
    camera = ArcballCamera(c["eye"], c["center"], c["up"])
    o = OracleBackend(max_depth=5)
    o.initialize(48, 48)
    o.set_scene(scene)
    rays = primary_rays(48, 48, camera.eye(), camera.dir(), camera.up(), c["fov_y"])
    hits, normals = o.trace_closest(rays, True)
    out.update(cornell_rays=rays, cornell_hits=hits, cornell_normals=normals)
    for f in range(2):
        st = o.render(camera.eye(), camera.dir(), camera.up(), c["fov_y"], f == 0)
    out["cornell_accum_48_spp2_f2"] = o.read_accum()
    out["cornell_img_48_spp2_f2"] = o.img.copy()
    out["cornell_rays_last_frame"] = np.array([st.num_rays], np.uint64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_kat.npz"), **out)
    print("wrote oracle_kat.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
