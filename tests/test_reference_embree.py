"""The oracle against the REFERENCE'S OWN Embree/ISPC backend.

tests/golden/ref_embree_frames.npz was produced by oracle/_ref/libcrt_embree.so, i.e. by
/root/reference/backends/embree/{render_embree.cpp,embree_utils.cpp,render_embree.ispc,*.ih} compiled where
they lie (ISPC kernels as scalar C++; Embree, TBB, GLM replaced by third_party/ stand-ins; see
oracle/ref_build/Makefile and tests/golden/make_ref_embree_golden.py). The oracle has to reproduce those
float framebuffers, per-pixel ray counts, sRGB8 images and pure-function tables BIT FOR BIT: that is what pins
it. Where the library itself is present (this container, and the GPU box as a prebuilt file) further cases
are compared live, including the plugin loaded by the reference's own RenderPlugin in the headless twin.
"""
import ctypes as C
import os

import numpy as np
import pytest

from ref_cases import FRAME_CASES, kat_inputs, make_case

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_embree_frames.npz"))


def _same(a, b):
    """Bitwise equality of float arrays, NaNs in the same places."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", list(FRAME_CASES))
def test_oracle_reproduces_reference_frames_bit_for_bit(built, ref_golden, name):
    from oracle import OracleBackend

    scene, view, w, h, frames, depth = make_case(name)
    cpu = OracleBackend(max_depth=depth)
    cpu.initialize(w, h)
    cpu.set_scene(scene)
    rays = [cpu.render(*view, f == 0, True).num_rays for f in range(frames)]
    assert rays == [int(r) for r in ref_golden[f"{name}.rays"]]
    assert np.array_equal(cpu.read_ray_stats(), ref_golden[f"{name}.ray_stats"])
    assert _same(cpu.read_accum(), ref_golden[f"{name}.accum"])
    assert np.array_equal(cpu.img, ref_golden[f"{name}.img"])


def test_oracle_pure_functions_equal_the_references(built, ref_golden):
    """disney_bsdf.ih, lights.ih, texture2d.ih, lcg_rng.ih, util.ih, miss_shader: the oracle's restatements
    against tables computed by the reference's functions themselves."""
    from oracle import load_oracle_lib

    lib = load_oracle_lib()
    k = kat_inputs()
    ev = np.zeros_like(ref_golden["kat.disney_eval"])
    for mi, m in enumerate(k["mats"]):
        for oi, wo in enumerate(k["dirs"]):
            for ii, wi in enumerate(k["dirs"]):
                lib.oracle_kat_disney_eval(m.ctypes.data, k["n"].ctypes.data, wo.ctypes.data, wi.ctypes.data,
                                           ev[mi, oi, ii].ctypes.data)
    assert _same(ev, ref_golden["kat.disney_eval"])
    sm = np.zeros_like(ref_golden["kat.disney_sample"])
    for mi, m in enumerate(k["mats"]):
        for oi, wo in enumerate(k["dirs"]):
            for si, seed in enumerate(k["seeds"]):
                st = C.c_uint32(int(seed))
                lib.oracle_kat_disney_sample(m.ctypes.data, k["n"].ctypes.data, wo.ctypes.data, C.addressof(st),
                                             sm[mi, oi, si].ctypes.data)
                sm[mi, oi, si, 7] = np.array([st.value], np.uint32).view(np.float32)[0]
    assert _same(sm, ref_golden["kat.disney_sample"])
    lt = np.zeros_like(ref_golden["kat.light"])
    for si, s2 in enumerate(k["light_s"]):
        for di, d in enumerate(k["dirs"]):
            lib.oracle_kat_light(k["light"].ctypes.data, s2.ctypes.data, k["light_orig"].ctypes.data, d.ctypes.data,
                                 lt[si, di].ctypes.data)
    assert _same(lt, ref_golden["kat.light"])
    for ch in (1, 3, 4):
        tex = k[f"tex{ch}"]
        tx = np.zeros((len(k["uv"]), 4), np.float32)
        lib.oracle_kat_texture(tex.ctypes.data, tex.shape[1], tex.shape[0], ch, k["uv"].ctypes.data, len(k["uv"]),
                               tx.ctypes.data)
        assert _same(tx, ref_golden[f"kat.texture{ch}"])
    ms = np.zeros((len(k["miss_dirs"]), 3), np.float32)
    lib.oracle_kat_miss(k["miss_dirs"].ctypes.data, len(k["miss_dirs"]), ms.ctypes.data)
    assert _same(ms, ref_golden["kat.miss"])
    ob = np.zeros((len(k["dirs"]), 6), np.float32)
    for di, d in enumerate(k["dirs"]):
        lib.oracle_kat_ortho_basis(d.ctypes.data, ob[di].ctypes.data)
    assert _same(ob, ref_golden["kat.ortho_basis"])
    for i, (pix, frame) in enumerate(k["rng_keys"]):
        states, floats = np.zeros(16, np.uint32), np.zeros(16, np.float32)
        lib.oracle_kat_rng(int(pix), int(frame), 16, states.ctypes.data, floats.ctypes.data)
        assert np.array_equal(states, ref_golden["kat.rng_states"][i]) and _same(floats, ref_golden["kat.rng_floats"][i])


# ------------------------------------------------------------------ live, where the library exists
def _ref_mod():
    from oracle import ref_embree

    if not ref_embree.available():
        pytest.skip("oracle/_ref/libcrt_embree.so not built (needs /root/reference at build time)")
    return ref_embree


def test_golden_file_is_what_the_reference_build_produces(built, ref_golden):
    """Guards the fixture itself: regenerating one case from the library gives the committed arrays."""
    ref_embree = _ref_mod()
    scene, view, w, h, frames, depth = make_case("materials")
    ref = ref_embree.RefEmbreeBackend(max_depth=depth)
    assert ref.name() == "Embree (w/ TBB & ISPC)"  # render_embree.cpp:33-36
    ref.initialize(w, h)
    ref.set_scene(scene)
    for f in range(frames):
        ref.render(*view, f == 0, True)
    assert _same(ref.read_accum(), ref_golden["materials.accum"])
    assert np.array_equal(ref.read_ray_stats(), ref_golden["materials.ray_stats"])


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_equals_reference_build_on_random_views(built, seed):
    """Other cameras, sizes, spp and frame counts than the frozen ones, camera reset included."""
    ref_embree = _ref_mod()
    from chameleonrt_b200 import ArcballCamera
    from chameleonrt_b200.scenes import cornell_box, sponza_like
    from helpers import synthetic_material_scene
    from oracle import OracleBackend

    rng = np.random.default_rng(seed)
    spp = int(rng.integers(1, 4))
    scene, cam = [cornell_box, synthetic_material_scene, lambda spp: sponza_like(spp=spp, detail=0.2, tex_size=32)][seed % 3](spp=spp)
    w, h = int(rng.integers(40, 150)), int(rng.integers(40, 110))
    depth = int(rng.integers(1, 9))
    ref, cpu = ref_embree.RefEmbreeBackend(max_depth=depth), OracleBackend(max_depth=depth)
    for r in (ref, cpu):
        r.initialize(w, h)
        r.set_scene(scene)
    for k in range(2):  # second pass: camera moved, accumulation restarts
        eye = np.array(cam["eye"], np.float32) + rng.normal(scale=0.15, size=3).astype(np.float32)
        c = ArcballCamera(eye, cam["center"], cam["up"])
        for f in range(int(rng.integers(1, 4))):
            sr = ref.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0, True)
            so = cpu.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0, True)
            assert sr.num_rays == so.num_rays
        assert _same(ref.read_accum(), cpu.read_accum())
        assert np.array_equal(ref.read_ray_stats(), cpu.read_ray_stats())
        assert np.array_equal(ref.img, cpu.img)


@pytest.mark.parametrize("name", ["cornell", "sponza"])
def test_headless_embree_plugin_equals_oracle_plugin(built, tmp_path, name):
    """`crt_headless embree scene.obj`: the reference's RenderPlugin loads libcrt_embree.so, the reference's
    Scene::load_obj loads the file, the reference's RenderEmbree renders it (calling thread in FTZ/DAZ mode
    like the real application); the oracle plugin on the same file gives the same float framebuffer bit for
    bit."""
    _ref_mod()
    from test_reference_plugin import HEADLESS, run_headless

    if not os.path.exists(HEADLESS):
        pytest.skip("oracle/_ref/crt_headless not built")
    from chameleonrt_b200.obj_io import write_obj
    from chameleonrt_b200.scenes import cornell_box, sponza_like

    scene, cam = cornell_box(spp=2) if name == "cornell" else sponza_like(spp=2, detail=0.2, tex_size=64)
    obj = write_obj(scene, str(tmp_path / "scene.obj"))
    a_ref, v1, out = run_headless("embree", obj, cam, 100, 76, 2, 2, tmp_path)
    a_cpu, v2, _ = run_headless("oracle", obj, cam, 100, 76, 2, 2, tmp_path)
    assert v1 == v2 and "Embree (w/ TBB & ISPC)" in out
    assert _same(a_ref, a_cpu)


def test_headless_embree_plugin_on_gltf_instances(built, tmp_path):
    """The reference's glTF loader + the reference's Embree backend (instanced scene: rtcSetGeometryTransform,
    world_to_object normals) against the oracle plugin on the same .gltf: bit for bit."""
    _ref_mod()
    from test_reference_plugin import HEADLESS, run_headless

    if not os.path.exists(HEADLESS):
        pytest.skip("oracle/_ref/crt_headless not built")
    from chameleonrt_b200.gltf_io import write_gltf
    from chameleonrt_b200.scenes import san_miguel_like

    scene, cam = san_miguel_like(spp=2, scale=0.02, tex_size=64)
    gltf = write_gltf(scene, str(tmp_path / "scene.gltf"))
    a_ref, v1, out = run_headless("embree", gltf, cam, 96, 54, 2, 2, tmp_path)
    a_cpu, v2, _ = run_headless("oracle", gltf, cam, 96, 54, 2, 2, tmp_path)
    assert v1 == v2 and "Embree (w/ TBB & ISPC)" in out
    assert _same(a_ref, a_cpu)


def test_headless_embree_plugin_on_crts_all_lobes(built, tmp_path):
    """Every BSDF lobe, textured scalar parameters and an explicit light, loaded by the reference's .crts
    importer and rendered by the reference's Embree backend, against the oracle plugin: bit for bit."""
    _ref_mod()
    from test_reference_plugin import HEADLESS, _crts_case, run_headless

    if not os.path.exists(HEADLESS):
        pytest.skip("oracle/_ref/crt_headless not built")
    scene, cam, crts = _crts_case(tmp_path)
    a_ref, v1, out = run_headless("embree", crts, cam, 96, 72, 2, 2, tmp_path)
    a_cpu, v2, _ = run_headless("oracle", crts, cam, 96, 72, 2, 2, tmp_path)
    assert v1 == v2 and "Embree (w/ TBB & ISPC)" in out
    assert _same(a_ref, a_cpu)


def test_fuzz_random_scenes_against_reference_build(built):
    """scripts/fuzz_oracle_vs_reference.py on a fixed set of seeds: random triangle soups and primitives, every
    Disney parameter random (textured scalars, transmission -> NaN paths), sheared / mirrored instances, several
    lights, random spp / depth / frames / ragged sizes. 2,100 seeds were clean when this was written; seed 1583 is
    the one that showed the oracle needs ISPC's min / max NaN semantics (minps / maxps return the second operand)."""
    _ref_mod()
    import sys

    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import fuzz_oracle_vs_reference as fuzz

    for seed in (0, 2, 5, 20, 31, 1583, 5007, 5333):
        ok, info = fuzz.one(seed)
        assert ok, (seed, info)
