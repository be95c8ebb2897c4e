"""The CPU oracle against the frozen golden vectors and against independent numpy/Python
restatements of the reference's pure functions (the reference ships no vectors of its own, SURVEY.md
§8c; the pin against the reference's own code is tests/test_reference_embree.py, these are the older,
independent ones)."""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import OracleBackend, load_oracle_lib
from oracle.oracle import primary_rays

fp = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def lib(built):
    return load_oracle_lib()


# ---- lcg_rng.ih:8-59, independent integer restatement ----
def _mix(h, k):
    k = (k * 0xCC9E2D51) & 0xFFFFFFFF
    k = ((k << 15) | (k >> 17)) & 0xFFFFFFFF
    k = (k * 0x1B873593) & 0xFFFFFFFF
    h ^= k
    h = ((h << 13) | (h >> 19)) & 0xFFFFFFFF
    return (h * 5 + 0xE6546B64) & 0xFFFFFFFF


def _fin(h):
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    return h ^ (h >> 16)


def test_rng_stream(lib, golden):
    for k, (pix, frame) in enumerate(golden["rng_keys"]):
        states = np.zeros(16, np.uint32)
        floats = np.zeros(16, np.float32)
        lib.oracle_kat_rng(int(pix), int(frame), 16, states.ctypes.data, floats.ctypes.data)
        assert (states == golden[f"rng_states_{k}"]).all()
        assert (floats.view(np.uint32) == golden[f"rng_floats_{k}"].view(np.uint32)).all()
        s = _fin(_mix(_mix(0, int(pix)), int(frame)))
        for i in range(16):
            s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
            assert s == int(states[i])
            assert np.float32(np.float32(s) * np.float32(2.0 ** -32)) == floats[i]


def test_rng_can_return_one(lib):
    # App. A #2: state >= 2^32 - 128 rounds to 2^32 -> exactly 1.0f
    assert np.float32(np.float32(0xFFFFFFFF) * np.float32(2.0 ** -32)) == np.float32(1.0)


def test_camera_basis(lib, golden):
    cin = golden["camera_in"]
    e, d, u = (np.ascontiguousarray(cin[i:i + 3]) for i in (0, 3, 6))
    out = np.zeros(12, np.float32)
    lib.oracle_kat_camera(e.ctypes.data_as(fp), d.ctypes.data_as(fp), u.ctypes.data_as(fp), C.c_float(float(cin[9])),
                          int(cin[10]), int(cin[11]), out.ctypes.data)
    assert (out.view(np.uint32) == golden["camera_basis"].view(np.uint32)).all()
    # independent float64 restatement of render_embree.cpp:149-159
    py = 2.0 * math.tan(math.radians(0.5 * float(cin[9])))
    px = py * cin[10] / cin[11]
    du = np.cross(d.astype(np.float64), u.astype(np.float64))
    du = du / np.linalg.norm(du) * px
    dv = -np.cross(du, d.astype(np.float64))
    dv = dv / np.linalg.norm(dv) * py
    tl = d - 0.5 * du - 0.5 * dv
    np.testing.assert_allclose(out[3:6], du, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[6:9], dv, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out[9:12], tl, rtol=1e-5, atol=1e-6)


def test_bsdf_eval_golden_and_properties(lib, golden):
    mats, n, dirs, ev = golden["bsdf_mats"], golden["bsdf_n"], golden["bsdf_dirs"], golden["bsdf_eval"]
    got = np.zeros_like(ev)
    for mi, m in enumerate(mats):
        m = np.ascontiguousarray(m)
        for oi in range(len(dirs)):
            for ii in range(len(dirs)):
                lib.oracle_kat_disney_eval(m.ctypes.data, n.ctypes.data, np.ascontiguousarray(dirs[oi]).ctypes.data,
                                           np.ascontiguousarray(dirs[ii]).ctypes.data, got[mi, oi, ii].ctypes.data)
    assert (got.view(np.uint32) == ev.view(np.uint32)).all()
    # properties of the model (disney_bsdf.ih:311-359)
    cos_o = dirs @ n
    cos_i = dirs @ n
    same = (cos_o[:, None] * cos_i[None, :]) > 0
    for mi, m in enumerate(mats):
        f, pdf = got[mi, ..., :3], got[mi, ..., 3]
        assert np.isfinite(f).all() and np.isfinite(pdf).all()
        up = cos_o > 0  # the renderer flips n toward w_o for opaque materials (ispc:297-299)
        assert (pdf[up] >= 0).all()
        if m[13] == 0.0:  # no transmission: opposite hemispheres carry nothing
            assert (f[~same] == 0).all()
            # reflection lobes only exist for w_i in n's hemisphere when w_o is too
        assert (f[same] >= 0).all()


def test_bsdf_sample_golden_and_consistency(lib, golden):
    mats, n, dirs = golden["bsdf_mats"], golden["bsdf_n"], golden["bsdf_dirs"]
    smp, smp_state = golden["bsdf_sample"], golden["bsdf_sample_state"]
    for mi, m in enumerate(mats):
        m = np.ascontiguousarray(m)
        for oi in range(8):
            wo = np.ascontiguousarray(dirs[oi])
            st = np.array([12345 + 977 * mi + oi], np.uint32)
            for k in range(16):
                o7 = np.zeros(7, np.float32)
                lib.oracle_kat_disney_sample(m.ctypes.data, n.ctypes.data, wo.ctypes.data, st.ctypes.data, o7.ctypes.data)
                assert (o7.view(np.uint32) == smp[mi, oi, k].view(np.uint32)).all()
                assert st[0] == smp_state[mi, oi, k]  # exactly three draws per sample (App. A #3)
                if o7[3] > 0:
                    # sample's (f, pdf) equals eval at the sampled direction
                    o4 = np.zeros(4, np.float32)
                    lib.oracle_kat_disney_eval(m.ctypes.data, n.ctypes.data, wo.ctypes.data,
                                               np.ascontiguousarray(o7[4:7]).ctypes.data, o4.ctypes.data)
                    assert (o4.view(np.uint32) == o7[:4].view(np.uint32)).all()
                    assert abs(float(np.linalg.norm(o7[4:7])) - 1.0) < 1e-4


def test_light_functions(lib, golden):
    light, ls, lo, ld = golden["light"], golden["light_s"], golden["light_o"], golden["light_d"]
    res = np.zeros((16, 9), np.float32)
    for i in range(16):
        lib.oracle_kat_light(light.ctypes.data, np.ascontiguousarray(ls[i]).ctypes.data,
                             np.ascontiguousarray(lo[i]).ctypes.data, np.ascontiguousarray(ld[i]).ctypes.data, res[i].ctypes.data)
    assert (res.view(np.uint32) == golden["light_res"].view(np.uint32)).all()
    pos, nrm, vx, w, vy, h = light[4:7], light[8:11], light[12:15], light[15], light[16:19], light[19]
    assert res[::2, 4].sum() >= 6  # the rays aimed at the quad hit it
    for i in range(16):
        p = ls[i, 0] * vx * w + ls[i, 1] * vy * h + pos  # lights.ih:26-30
        np.testing.assert_allclose(res[i, :3], p, rtol=1e-5, atol=1e-5)
        # quad_light_pdf quirk: to_pt = p - dir (lights.ih:41)
        ndw = float(np.dot(nrm, -ld[i]))
        expect = 0.0 if ndw < 1e-4 else float(np.dot(p - ld[i], p - ld[i])) / (ndw * w * h)
        assert res[i, 3] == pytest.approx(expect, rel=1e-4, abs=1e-6)
        denom = float(np.dot(ld[i], nrm))
        t = float(np.dot(pos - lo[i], nrm)) / denom
        hv = lo[i] + ld[i] * t - pos
        hit = t >= 0 and abs(np.dot(hv, vx)) < w and abs(np.dot(hv, vy)) < h  # lights.ih:62-65
        assert bool(res[i, 4]) == bool(hit)


def test_texture_filter(lib, golden):
    tex, uv = golden["tex_data"], golden["tex_uv"]
    res = np.zeros((len(uv), 4), np.float32)
    lib.oracle_kat_texture(tex.ctypes.data, 16, 16, 4, uv.ctypes.data, len(uv), res.ctypes.data)
    assert (res.view(np.uint32) == golden["tex_res"].view(np.uint32)).all()
    # independent restatement incl. the float->int truncation quirk (App. A #12)
    for k, (u, v) in enumerate(uv):
        ux, uy = np.float32(u * 16 - 0.5), np.float32(v * 16 - 0.5)
        tx, ty = ux - np.floor(ux), uy - np.floor(uy)
        x0, y0 = int(ux) % 16, int(uy) % 16  # int() truncates toward zero; % is non-negative in Python
        x1, y1 = int(np.float32(ux + 1)) % 16, int(np.float32(uy + 1)) % 16
        t = tex.astype(np.float64) / 255.0
        e = (t[y0, x0] * (1 - tx) * (1 - ty) + t[y0, x1] * tx * (1 - ty) + t[y1, x0] * (1 - tx) * ty + t[y1, x1] * tx * ty)
        np.testing.assert_allclose(res[k], e, rtol=1e-5, atol=1e-6)


def test_miss_ortho_srgb_tri(lib, golden):
    md = golden["miss_dirs"]
    mres = np.zeros((32, 3), np.float32)
    lib.oracle_kat_miss(md.ctypes.data, 32, mres.ctypes.data)
    assert (mres == golden["miss_res"]).all()
    for i, d in enumerate(md):  # render_embree.ispc:183-196
        u = (1.0 + math.atan2(d[0], -d[2]) / math.pi) * 0.5
        v = math.acos(d[1]) / math.pi
        e = 0.5 if (d[1] > -0.1 and (int(u * 10) + int(v * 10)) % 2 == 0) else 0.1
        assert mres[i, 0] == pytest.approx(e)
    ob = np.zeros((32, 6), np.float32)
    for i in range(32):
        lib.oracle_kat_ortho_basis(np.ascontiguousarray(md[i]).ctypes.data, ob[i].ctypes.data)
    assert (ob.view(np.uint32) == golden["ortho_res"].view(np.uint32)).all()
    for i in range(32):  # orthonormal frame around n
        vx, vy = ob[i, :3], ob[i, 3:]
        assert abs(np.dot(vx, vy)) < 1e-5 and abs(np.dot(vx, md[i])) < 1e-5 and abs(np.dot(vy, md[i])) < 1e-5
        assert np.linalg.norm(vx) == pytest.approx(1, abs=1e-5) and np.linalg.norm(vy) == pytest.approx(1, abs=1e-5)
    sx = golden["srgb_in"]
    sres = np.zeros(len(sx), np.uint8)
    lib.oracle_kat_srgb8(sx.ctypes.data, len(sx), sres.ctypes.data)
    assert (sres == golden["srgb_res"]).all()
    for x, r in zip(sx, sres):
        if not (x > 0):
            e = 0
        elif x >= 1:
            e = 255
        else:
            s = 12.92 * x if x <= 0.0031308 else 1.055 * x ** (1 / 2.4) - 0.055
            e = int(s * 255 + 0.5)
        assert abs(int(r) - e) <= 1
    tri, trays = golden["tri"], golden["tri_rays"]
    tout = np.zeros((12, 4), np.float32)
    for i in range(12):
        lib.oracle_kat_tri(tri.ctypes.data, np.ascontiguousarray(trays[i]).ctypes.data, tout[i].ctypes.data)
    g = golden["tri_res"]
    assert ((tout.view(np.uint32) == g.view(np.uint32)) | (np.isnan(tout) & np.isnan(g))).all()
    for i in range(10):  # rays along -z onto the unit right triangle in z=0
        x, y = trays[i, 0], trays[i, 1]
        inside = x >= 0 and y >= 0 and x + y <= 1
        assert bool(tout[i, 3]) == bool(inside)
        if inside:
            assert tout[i, 0] == pytest.approx(1.0) and tout[i, 1] == pytest.approx(x, abs=1e-6) and tout[i, 2] == pytest.approx(y, abs=1e-6)
    assert tout[10, 3] == 0  # parallel ray (det == 0) must miss, NaN-safe
    assert tout[11, 3] == 0  # hit beyond tfar is rejected


def test_cornell_golden_frame(built, golden):
    from chameleonrt_b200.scenes import cornell_box
    from helpers import camera_for

    scene, cam = cornell_box(spp=2)
    c = camera_for(cam)
    o = OracleBackend(max_depth=5)
    o.initialize(48, 48)
    o.set_scene(scene)
    rays = primary_rays(48, 48, c.eye(), c.dir(), c.up(), cam["fov_y"])
    assert (rays.view(np.uint32) == golden["cornell_rays"].view(np.uint32)).all()
    hits, normals = o.trace_closest(rays, True)
    assert (hits.view(np.uint32) == golden["cornell_hits"].view(np.uint32)).all()
    assert (normals.view(np.uint32) == golden["cornell_normals"].view(np.uint32)).all()
    for f in range(2):
        st = o.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0)
    assert (o.read_accum().view(np.uint32) == golden["cornell_accum_48_spp2_f2"].view(np.uint32)).all()
    assert (o.img == golden["cornell_img_48_spp2_f2"]).all()
    assert st.num_rays == int(golden["cornell_rays_last_frame"][0])
    # thread-count independence (tiles are independent, render_embree.cpp:178)
    o1 = OracleBackend(max_depth=5, num_threads=1)
    o1.initialize(48, 48)
    o1.set_scene(scene)
    for f in range(2):
        o1.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0)
    assert (o1.read_accum().view(np.uint32) == golden["cornell_accum_48_spp2_f2"].view(np.uint32)).all()


def test_oracle_bvh_matches_brute_force(built):
    """The oracle's own BVH2 (which stands in for Embree) against exhaustive intersection."""
    from chameleonrt_b200.scenes import cornell_box, sponza_like
    from helpers import bounce_rays, camera_for

    for make in (lambda: cornell_box(), lambda: sponza_like(detail=0.12, tex_size=16)):
        scene, cam = make()
        c = camera_for(cam)
        fast = OracleBackend()
        slow = OracleBackend(brute_force=True)
        for o in (fast, slow):
            o.initialize(8, 8)
            o.set_scene(scene)
        rays = primary_rays(40, 24, c.eye(), c.dir(), c.up(), cam["fov_y"])
        h = fast.trace_closest(rays)
        rays = np.concatenate([rays, bounce_rays(rays, h)])
        hf, hs = fast.trace_closest(rays), slow.trace_closest(rays)
        assert (hf.view(np.uint32) == hs.view(np.uint32)).all()
        assert (fast.trace_any(rays) == slow.trace_any(rays)).all()
