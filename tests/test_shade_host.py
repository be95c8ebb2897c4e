"""The product's DEVICE shading source (chameleonrt_b200/csrc/shade_math.cuh), compiled for the host
(TEST-ONLY libcrt_shade_hostcheck.so), against tables computed by the reference's own functions
(tests/golden/ref_embree_frames.npz: disney_bsdf.ih, lights.ih, texture2d.ih, lcg_rng.ih, util.ih, miss_shader).
Everything is bit-exact except where DESIGN.md §4 says otherwise: the Schlick weight (1-cos)^5 is evaluated by
multiplication on the device, by pow() in the reference, which moves BSDF values by a few ulp."""
import ctypes as C
import os

import numpy as np
import pytest

from ref_cases import kat_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(built):
    path = os.path.join(ROOT, "chameleonrt_b200", "csrc", "libcrt_shade_hostcheck.so")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.shadekat_rng.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    lib.shadekat_disney_eval.argtypes = [vp] * 5
    lib.shadekat_disney_sample.argtypes = [vp] * 5
    lib.shadekat_light.argtypes = [vp] * 5
    lib.shadekat_texture.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    lib.shadekat_miss.argtypes = [vp, C.c_int, vp]
    lib.shadekat_ortho_basis.argtypes = [vp, vp]
    lib.shadekat_srgb8.argtypes = [vp, C.c_int, vp]
    return lib


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_embree_frames.npz"))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _close(got, want, rel=2e-6, abs_=1e-12):
    """Few-ulp agreement, NaN / inf in the same places."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    d = np.abs(got[fin] - want[fin])
    return bool((d <= abs_ + rel * np.abs(want[fin])).all()), float((d / np.maximum(np.abs(want[fin]), 1e-30)).max())


def test_exact_functions(lib, ref):
    k = kat_inputs()
    for i, (pix, frame) in enumerate(k["rng_keys"]):
        states, floats = np.zeros(16, np.uint32), np.zeros(16, np.float32)
        lib.shadekat_rng(int(pix), int(frame), 16, states.ctypes.data, floats.ctypes.data)
        assert np.array_equal(states, ref["kat.rng_states"][i]) and np.array_equal(_bits(floats), _bits(ref["kat.rng_floats"][i]))
    lt = np.zeros_like(ref["kat.light"])
    for si, s2 in enumerate(k["light_s"]):
        for di, d in enumerate(k["dirs"]):
            lib.shadekat_light(k["light"].ctypes.data, s2.ctypes.data, k["light_orig"].ctypes.data, d.ctypes.data,
                               lt[si, di].ctypes.data)
    assert np.array_equal(_bits(lt), _bits(ref["kat.light"]))
    for ch in (1, 3, 4):
        tex = k[f"tex{ch}"]
        tx = np.zeros((len(k["uv"]), 4), np.float32)
        lib.shadekat_texture(tex.ctypes.data, tex.shape[1], tex.shape[0], ch, k["uv"].ctypes.data, len(k["uv"]), tx.ctypes.data)
        assert np.array_equal(_bits(tx), _bits(ref[f"kat.texture{ch}"]))
    ms = np.zeros((len(k["miss_dirs"]), 3), np.float32)
    lib.shadekat_miss(k["miss_dirs"].ctypes.data, len(k["miss_dirs"]), ms.ctypes.data)
    assert np.array_equal(_bits(ms), _bits(ref["kat.miss"]))
    ob = np.zeros((len(k["dirs"]), 6), np.float32)
    for di, d in enumerate(k["dirs"]):
        lib.shadekat_ortho_basis(d.ctypes.data, ob[di].ctypes.data)
    assert np.array_equal(_bits(ob), _bits(ref["kat.ortho_basis"]))


def test_disney_bsdf_eval_and_sample(lib, ref):
    k = kat_inputs()
    ev = np.zeros_like(ref["kat.disney_eval"])
    for mi, m in enumerate(k["mats"]):
        for oi, wo in enumerate(k["dirs"]):
            for ii, wi in enumerate(k["dirs"]):
                lib.shadekat_disney_eval(m.ctypes.data, k["n"].ctypes.data, wo.ctypes.data, wi.ctypes.data,
                                         ev[mi, oi, ii].ctypes.data)
    ok, worst = _close(ev, ref["kat.disney_eval"])
    assert ok, f"BSDF eval/pdf: worst relative difference {worst:.3e}"
    # the pdf involves no Schlick weight: bit-exact
    assert np.array_equal(_bits(ev[..., 3]), _bits(ref["kat.disney_eval"][..., 3]))
    sm = np.zeros_like(ref["kat.disney_sample"])
    for mi, m in enumerate(k["mats"]):
        for oi, wo in enumerate(k["dirs"]):
            for si, seed in enumerate(k["seeds"]):
                st = C.c_uint32(int(seed))
                lib.shadekat_disney_sample(m.ctypes.data, k["n"].ctypes.data, wo.ctypes.data, C.addressof(st),
                                           sm[mi, oi, si].ctypes.data)
                sm[mi, oi, si, 7] = np.array([st.value], np.uint32).view(np.float32)[0]
    want = ref["kat.disney_sample"]
    # sampled direction, pdf and the rng state after the call are bit-exact; f differs only by the Schlick weight
    assert np.array_equal(_bits(sm[..., 3:]), _bits(want[..., 3:]))
    ok, worst = _close(sm[..., :3], want[..., :3])
    assert ok, f"sampled BSDF value: worst relative difference {worst:.3e}"
