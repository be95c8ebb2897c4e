"""The PRODUCT's renderer object and C ABI (chameleonrt_b200/csrc/crt_cuda_core.cu + kernels.cuh) running end to end
on a machine without a GPU: tests/simt_emu builds the same sources for the host, executes every kernel launch under
the SIMT environment of chameleonrt_b200/csrc/simt_env.h (an OS thread per CUDA thread, barrier-backed warp
collectives, real shared memory and atomics) and replaces the CUDA runtime by two dozen host stubs. A RenderCUDA is
pointed at that library (here only — nothing in chameleonrt_b200/ can load it) and put through the SAME test functions
the B200 runs (tests/test_gpu_parity.py), at the sizes the emulation can afford, plus
small versions of the frames-in-flight, tile-sharding and shadow-ray-order tests (under a second per 48x48x2spp frame). This checks the launch sequence,
queue hand-offs, counters, frame records, options and error paths of the real orchestration code, and dry-runs the
GPU tests' own code; what stays GPU-only is timing and CUDA's floating-point library.
Set CRT_SIMT_FULL=1 for the longer cases."""
import os
import sys

import numpy as np
import pytest

from helpers import assert_parity, camera_for, parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.environ.get("CRT_SIMT_FULL") == "1"


@pytest.fixture(scope="module")
def mods(built):
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    import build as simt_build

    lib_path = simt_build.build()
    import chameleonrt_b200.backend as backend
    from oracle import OracleBackend
    from oracle.oracle import primary_rays

    saved = (backend._LIB_PATH, backend._lib)
    backend._LIB_PATH, backend._lib = lib_path, None
    yield backend.RenderCUDA, OracleBackend, primary_rays
    backend._LIB_PATH, backend._lib = saved


@pytest.fixture(scope="module")
def gpu_tests():
    import test_gpu_parity

    return test_gpu_parity


def test_the_gpu_suites_own_tests_on_the_emulated_renderer(mods, gpu_tests, golden):
    gpu_tests.test_error_behaviour(mods)
    gpu_tests.test_traversal_empty_batch_and_misses(mods)
    gpu_tests.test_traversal_golden_cornell(mods, golden)
    gpu_tests.test_cornell_frame_matches_golden_and_oracle(mods, golden)
    gpu_tests.test_camera_changed_resets_accumulation(mods)
    gpu_tests.test_ragged_framebuffer_and_resize(mods)
    gpu_tests.test_async_frames_match_blocking_frames(mods)


def test_reference_frames_on_the_emulated_renderer(mods, gpu_tests):
    ref_golden = np.load(os.path.join(ROOT, "tests", "golden", "ref_embree_frames.npz"))
    for name in ("cornell_d8", "ragged_70x50", "cornell", "materials", "rungholt_like") + (("sponza_like", "sponza_like_d8", "san_miguel_like_instances") if FULL else ()):
        # (the device-built tree on two of them: the emulated device build of a larger scene takes minutes)
        for builder in ("host", "device") if name in ("cornell_d8", "ragged_70x50") else ("host",):
            gpu_tests.test_cuda_matches_reference_embree_frames(mods, ref_golden, name, builder)


def test_frames_in_flight_sharding_and_shadow_order_on_the_emulated_renderer(mods):
    from chameleonrt_b200.scenes import cornell_box

    RenderCUDA = mods[0]
    scene, cam = cornell_box(spp=1)
    c = camera_for(cam)
    args = (c.eye(), c.dir(), c.up(), cam["fov_y"])
    w, h = 70, 40  # two tiles, ragged
    single = RenderCUDA(0)
    single.initialize(w, h)
    single.set_scene(scene)
    rays = sum(single.render(*args, f == 0, True).num_rays for f in range(3))
    want, want_img = single.read_accum(), single.read_img()
    # render_async + sync, one frame at a time and as a 1 + 2 batch
    for batches in ((1, 1, 1), (1, 2)):
        r = RenderCUDA(0)
        r.initialize(w, h)
        r.set_scene(scene)
        for i, nb in enumerate(batches):
            r.render_async(*args, i == 0, nb)
        totals, stages, counters, n = r.sync()
        assert n == 3 and totals.num_rays == rays and counters["kernel_launches"] == len(batches) * (2 + 3 * 5 + 1)
        assert (r.read_accum().view(np.uint32) == want.view(np.uint32)).all() and (r.read_img() == want_img).all()
    # two ranks, each rendering its tile of the three frames as wavefronts of 2 + 1, assembled on rank 0
    ranks, got = [], 0
    for rank in range(2):
        r = RenderCUDA(0, rank=rank, world_size=2)
        r.initialize(w, h)
        r.set_scene(scene)
        r.render_async(*args, True, 2)
        r.render_async(*args, False, 1)
        got += r.sync()[0].num_rays
        ranks.append(r)
    for src, r in enumerate(ranks):
        acc, img, ntiles = r.local_buffers()
        assert ntiles == 1
        ranks[0].assemble_rank(src, 2, acc, img)
    assert got == rays and (ranks[0].read_accum().view(np.uint32) == want.view(np.uint32)).all()
    assert (ranks[0].read_img() == want_img).all()
    # the same two ranks WITHOUT a gather: rank 0 exports its frame, rank 1 maps it, both resolve kernels write into it
    ranks, got = [], 0
    for rank in range(2):
        r = RenderCUDA(0, rank=rank, world_size=2)
        r.initialize(w, h)
        r.set_scene(scene)
        ranks.append(r)
    ranks[1].import_frame(ranks[0].export_frame())
    for r in ranks:
        r.render_async(*args, True, 2)
        r.render_async(*args, False, 1)
        got += r.sync()[0].num_rays
    assert got == rays and (ranks[0].read_accum().view(np.uint32) == want.view(np.uint32)).all()
    assert (ranks[0].read_img() == want_img).all()
    ranks[1].import_frame(None)
    with pytest.raises(ValueError):
        ranks[1].import_frame(b"short")
    # error behaviour of the frame-sharing entry points
    fresh = RenderCUDA(0, rank=1, world_size=2)
    with pytest.raises(RuntimeError, match="initialize"):
        fresh.export_frame()
    with pytest.raises(RuntimeError, match="same size"):
        ranks[0].share_frame_with(fresh)  # not initialized
    fresh.initialize(w + 64, h)
    with pytest.raises(RuntimeError, match="same size"):
        ranks[0].share_frame_with(fresh)
    with pytest.raises(RuntimeError, match="same renderer"):
        ranks[0].share_frame_with(ranks[0])
    # a resize undoes the sharing: both renderers are back to tile-local results that wait for a gather
    ranks[0].share_frame_with(ranks[1])
    for r in ranks:
        r.initialize(w, h)
        r.set_scene(scene)
    ranks[0].render(*args, True, True)
    ranks[1].render(*args, True, True)
    assert (ranks[0].read_accum() == 0).all()  # nothing has been written into rank 0's full frame
    # the torch.distributed wrapper of the same (distributed.PeerFrame), its collectives replaced by an in-process stand-in:
    # rank 0 exports + "broadcasts", rank 1 receives + imports, finish() is the barrier
    import torch
    import torch.distributed as dist

    from chameleonrt_b200.distributed import PeerFrame

    state = {"rank": 0, "wire": None, "barriers": 0}
    saved = {k: getattr(dist, k) for k in ("get_world_size", "get_rank", "get_backend", "broadcast", "barrier")}
    saved_stream = torch.cuda.current_stream

    def fake_broadcast(t, src, group=None):
        if state["rank"] == src:
            state["wire"] = t.clone()
        else:
            t.copy_(state["wire"])

    class _NoStream:
        def synchronize(self):
            pass

    try:
        dist.get_world_size, dist.get_rank, dist.get_backend = (lambda g=None: 2), (lambda g=None: state["rank"]), (lambda g=None: "gloo")
        dist.broadcast = fake_broadcast
        dist.barrier = lambda group=None: state.__setitem__("barriers", state["barriers"] + 1)
        torch.cuda.current_stream = lambda d=None: _NoStream()
        ranks, frames_ = [], []
        for rank in range(2):
            r = RenderCUDA(0, rank=rank, world_size=2)
            r.initialize(w, h)
            r.set_scene(scene)
            ranks.append(r)
            state["rank"] = rank
            frames_.append(PeerFrame(r))
        got = 0
        for rank, r in enumerate(ranks):
            state["rank"] = rank
            r.render_async(*args, True, 2)
            r.render_async(*args, False, 1)
            got += r.sync()[0].num_rays
        # (the emulation runs the ranks one after the other: the assembling rank's wait for the others' completion flags
        # comes after they have rendered; on GPUs the ranks run side by side and the wait kernel simply spins)
        for rank in (1, 0):
            state["rank"] = rank
            frames_[rank].submit()
            assert frames_[rank].finish() == (rank == 0)
        assert got == rays and state["barriers"] == 2  # (one per PeerFrame constructor: the frame is mapped everywhere)
        assert (ranks[0].read_accum().view(np.uint32) == want.view(np.uint32)).all() and (ranks[0].read_img() == want_img).all()
        ranks[1].import_frame(None)
    finally:
        for k, v in saved.items():
            setattr(dist, k, v)
        torch.cuda.current_stream = saved_stream
    # shadow rays far-first: on, and decided per scene from frames 1 and 2 — same image
    for mode in (1, 2):
        r = RenderCUDA(0, any_far_first=mode)
        r.initialize(w, h)
        r.set_scene(scene)
        assert sum(r.render(*args, f == 0, True).num_rays for f in range(3)) == rays
        assert (r.read_accum().view(np.uint32) == want.view(np.uint32)).all()


@pytest.mark.skipif(not FULL, reason="set CRT_SIMT_FULL=1 (several minutes under the emulation)")
def test_longer_gpu_tests_on_the_emulated_renderer(mods, gpu_tests):
    gpu_tests.test_sharded_ranks_with_frames_in_flight(mods)
    gpu_tests.test_shadow_ray_order_option_does_not_change_the_image(mods)
    gpu_tests.test_all_bsdf_lobes_and_textured_params(mods)
    gpu_tests.test_instanced_textured_gltf_class_scene(mods)
    for spp, frames, depth in [(1, 1, 5), (1, 3, 5), (4, 2, 5), (2, 2, 8), (1, 1, 1)]:
        gpu_tests.test_cornell_frames(mods, spp, frames, depth)


def test_drop_in_plugin_on_the_emulated_core(built, tmp_path):
    """`crt_headless cuda_simt <scene>`: the C++ plugin of backends/cuda (unchanged sources) linked against the
    emulated core — the reference's own loaders (OBJ, glTF with instances, .crts with every Disney parameter and an
    explicit light), RenderPlugin, RenderCUDA : RenderBackend, the C ABI and the kernels, without a GPU. Same checks
    as the B200's test_cuda_plugin_drop_in* tests."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
    import build as simt_build
    from test_reference_plugin import HEADLESS, _crts_case, run_headless

    if not os.path.exists(HEADLESS) or simt_build.build_plugin() is None:
        pytest.skip("needs oracle/_ref (built where /root/reference is)")
    from chameleonrt_b200.gltf_io import write_gltf
    from chameleonrt_b200.obj_io import write_obj
    from chameleonrt_b200.scenes import cornell_box, san_miguel_like

    scene, cam = cornell_box(spp=2)
    obj = write_obj(scene, str(tmp_path / "scene.obj"))
    a_gpu, v1, out = run_headless("cuda_simt", obj, cam, 96, 64, 2, 2, tmp_path)
    a_cpu, v2, _ = run_headless("oracle", obj, cam, 96, 64, 2, 2, tmp_path)
    assert v1 == v2 and "CUDA wavefront" in out
    assert_parity(a_gpu, a_cpu)
    # the plugins' second extra export (crt_<backend>_get_stats): per-stage times and counters of the last frame
    stats = [l for l in out.splitlines() if l.startswith("last frame: stage ms")]
    assert len(stats) == 1
    stage, counters = (list(map(float, part.split())) for part in stats[0][len("last frame: stage ms"):].split("| counters"))
    assert len(stage) == 7 and stage[6] > 0 and stage[2] > 0 and len(counters) == 8 and counters[0] > 0 and counters[1] > 0
    assert counters[2] == 2 + 3 * 5 + 1  # kernel launches of one depth-5 frame
    # the plugin's environment knob for the device BVH builders: same frame, bit for bit
    for builder in ("1", "2"):
        a_dev, _, _ = run_headless("cuda_simt", obj, cam, 96, 64, 2, 2, tmp_path, extra_env={"CRT_CUDA_BVH_BUILDER": builder})
        assert np.array_equal(a_dev.view(np.uint32), a_gpu.view(np.uint32))
    # CRT_CUDA_DEVICES: the plugin drives one renderer per GPU from its single host thread, tiles interleaved, every
    # renderer resolving into the first one's frame (crtc_share_frame) — same frame, bit for bit (the emulation has 8
    # "devices"); also two renderers on one device, the form a one-GPU box can run
    # (five frames: the first three after set_scene are rendered blocking — the shadow-ray order is chosen from their stage
    # times —, the rest fanned out with crtc_render_async)
    a_five, _, _ = run_headless("cuda_simt", obj, cam, 96, 64, 2, 5, tmp_path)
    for devices in ("0,1,2", "0,0"):
        a_multi, _, out_multi = run_headless("cuda_simt", obj, cam, 96, 64, 2, 5, tmp_path, extra_env={"CRT_CUDA_DEVICES": devices})
        assert np.array_equal(a_multi.view(np.uint32), a_five.view(np.uint32)), devices
    with pytest.raises(AssertionError, match="bvh_builder must be"):  # the knob does reach the core
        run_headless("cuda_simt", obj, cam, 96, 64, 2, 2, tmp_path, extra_env={"CRT_CUDA_BVH_BUILDER": "7"})
    if not FULL:
        return
    scene, cam = san_miguel_like(spp=2, scale=0.02, tex_size=64)
    gltf = write_gltf(scene, str(tmp_path / "scene.gltf"))
    a_gpu, v1, out = run_headless("cuda_simt", gltf, cam, 192, 108, 2, 2, tmp_path)
    a_cpu, v2, _ = run_headless("oracle", gltf, cam, 192, 108, 2, 2, tmp_path)
    assert v1 == v2 and "CUDA wavefront" in out
    assert_parity(a_gpu, a_cpu, min_frac=0.99, max_rel_l1=1e-2)
    scene, cam, crts = _crts_case(tmp_path)
    a_gpu, v1, out = run_headless("cuda_simt", crts, cam, 192, 128, 2, 2, tmp_path, depth=6)
    a_cpu, v2, _ = run_headless("oracle", crts, cam, 192, 128, 2, 2, tmp_path, depth=6)
    assert v1 == v2 and "CUDA wavefront" in out
    assert_parity(a_gpu, a_cpu, min_frac=0.99, max_rel_l1=1e-2)


def test_graft_entry_smoke_on_the_emulated_renderer(mods, capsys):
    """__graft_entry__.smoke() — what the driver runs on the B200 before the bench — dry-run on the emulation."""
    import __graft_entry__ as g

    g.smoke()
    assert "pixels within tol=1.00000" in capsys.readouterr().out


def test_device_bvh_builder_on_the_emulated_renderer(mods):
    """set_scene with bvh_builder="device" (bvh8_device.cuh: Morton keys, the hand-written radix sort and scans, Karras
    hierarchy, refit + collapse DP with inter-block atomics, level-wise BVH8 emission) executed under the SIMT
    emulation: frames bit-identical to the host-built tree's, and the edge cases of the GPU test."""
    from chameleonrt_b200.scenes import cornell_box, rungholt_like, san_miguel_like, sponza_like

    cases = [("cornell", lambda: cornell_box(spp=2), 40, 40, 2, 5),
             ("sponza", lambda: sponza_like(spp=1, detail=0.2, tex_size=32), 48, 32, 1, 8),
             ("instances", lambda: san_miguel_like(spp=1, scale=0.02, tex_size=32), 48, 32, 1, 5)]
    if FULL:
        cases.append(("voxels", lambda: rungholt_like(spp=1, scale=0.1), 48, 32, 1, 5))
    import test_z_new_gpu_paths as dev_tests

    dev_tests.test_device_built_bvh_renders_the_same_image(mods, cases)
    dev_tests.test_device_built_bvh_edge_cases(mods, big=4099 if FULL else 2500)
    dev_tests.test_device_set_scene_rejects_bad_input(mods)
    dev_tests.test_ploc_tail_kernel_builds_the_same_tree(mods, detail=0.2, size=(48, 32))


def test_triangle_pass_deferral_on_the_emulated_renderer(mods):
    """k_traverse<COUNT, DEFER> (option tri_pass_defer) in the whole renderer: same frames, same ray counts."""
    import test_z_new_gpu_paths as dev_tests

    dev_tests.test_triangle_pass_deferral_does_not_change_the_image(mods, size=(48, 32), detail=0.2)


def test_shade_queue_sort_on_the_emulated_renderer(mods):
    """The material-id bucketing of the shade queue (option shade_sort) through the real launch sequence."""
    import test_z_new_gpu_paths as dev_tests

    dev_tests.test_shade_queue_sort_does_not_change_the_image(mods, size=(48, 32), detail=0.2, frames=1, batch=2)


def test_triangle_count_limit_is_enforced_before_anything_is_allocated(mods):
    """k_traverse packs (owner lane, leaf-order triangle index) into 5 + 27 bits: a scene of 2^27 or more flattened triangles
    is refused by set_scene on both the host and the device path, from the instance / geometry counts alone (ADVICE r1:
    nothing enforced it, and ~13 GB of records fit a B200)."""
    from chameleonrt_b200.scene import DisneyMaterial, Geometry, Instance, Mesh, ParameterizedMesh, Scene, default_obj_light

    RenderCUDA = mods[0]
    n = 1 << 16
    verts = np.zeros((3, 3), np.float32)
    verts[1, 0] = verts[2, 1] = 1.0
    idx = np.tile(np.array([[0, 1, 2]], np.uint32), (n, 1))
    inst = [Instance(np.eye(4, dtype=np.float32), 0) for _ in range((1 << 27) // n)]  # 2048 x 65536 = 2^27 triangles
    scene = Scene(meshes=[Mesh([Geometry(verts, idx)])], parameterized_meshes=[ParameterizedMesh(0, [0])], instances=inst,
                  materials=[DisneyMaterial()], lights=[default_obj_light()])
    for builder in ("host", "device"):
        r = RenderCUDA(0, bvh_builder=builder)
        r.initialize(16, 16)
        with pytest.raises(RuntimeError, match="at most 2\\^27 - 1"):
            r.set_scene(scene)
    scene.instances = inst[:-1]  # one instance fewer is fine as far as the limit goes (not built here: 134 M triangles)


def test_round_2_options_on_the_emulated_renderer(mods):
    """stage_events = 0 (events only at frame start / end: the frame time stays, the per-stage times read 0, the image
    does not change — and the shadow-order trial of the first two frames still sees its traversal times), the host-buffer
    pinning knobs, bvh_top_smem (the shared-memory copy of the top of the tree: same image), and hw_textures, which the
    host emulation cannot provide: a textured scene's set_scene fails cleanly, an untextured one renders."""
    from chameleonrt_b200.scenes import cornell_box, sponza_like

    RenderCUDA = mods[0]
    scene, cam = cornell_box(spp=1)
    c = camera_for(cam)
    args = (c.eye(), c.dir(), c.up(), cam["fov_y"])
    out = {}
    for name, opts in (("default", {}), ("no_stage_events", {"stage_events": 0, "pin_host_buffers": 0, "pin_read_img": 1}),
                       ("top_smem", {"bvh_top_smem": 1}), ("hw_tex_untextured", {"hw_textures": 1})):
        r = RenderCUDA(0)
        for k, v in opts.items():
            r.set_option(k, v)
        r.initialize(48, 32)
        r.set_scene(scene)
        for f in range(4):
            st = r.render(*args, f == 0, True)
        out[name] = (r.read_accum(), r.read_img(r.img).copy(), st.num_rays, r.stage_times(), r.get_option("any_far_first_decision"))
    a0, i0, n0, t0, d0 = out["default"]
    for name, (a, i, n, t, d) in out.items():
        assert np.array_equal(a.view(np.uint32), a0.view(np.uint32)) and np.array_equal(i, i0) and n == n0, name
        assert t["frame"] > 0 and d in (0, 1), name
    assert out["default"][3]["traverse"] > 0 and out["no_stage_events"][3]["traverse"] == 0 and out["no_stage_events"][3]["shade"] == 0
    scene, _ = sponza_like(spp=1, detail=0.25, tex_size=16)
    r = RenderCUDA(0)
    r.set_option("hw_textures", 1)
    r.initialize(16, 16)
    with pytest.raises(RuntimeError, match="cudaMallocArray|not supported"):
        r.set_scene(scene)


def test_natively_loaded_crts_renders_the_same_frame(mods, tmp_path, size=(48, 32)):
    """.crts -> crtio_load_crts (geometry arrays used in place in the mapped file; several meshes, parameterized meshes and
    instances, every material parameter, textured scalars, an explicit light) -> crtc_set_scene -> the frame of the Python
    scene model of what Scene::load_crts builds (crts_io.crts_scene_view), bit for bit."""
    pytest.importorskip("PIL")
    from chameleonrt_b200 import scene_io
    from chameleonrt_b200.crts_io import crts_scene_view, write_crts
    from chameleonrt_b200.scene import QuadLight
    from helpers import synthetic_material_scene

    RenderCUDA = mods[0]
    scene, cam = synthetic_material_scene(spp=2)
    scene.lights = [QuadLight(emission=(12.0, 11.0, 9.0, 1.0), position=(0.5, 4.5, 0.5, 1.0), normal=(0.0, -1.0, 0.0),
                              v_x=(1.0, 0.0, 0.0), width=1.5, v_y=(0.0, 0.0, 1.0), height=1.0)]
    c = camera_for(cam)
    args = (c.eye(), c.dir(), c.up(), cam["fov_y"])
    loaded = scene_io.load_scene(write_crts(scene, str(tmp_path / "materials.crts")))
    frames = []
    for native in (False, True):
        r = RenderCUDA(0)
        r.initialize(*size)
        if native:
            r.set_scene_c(loaded.c_scene, samples_per_pixel=2)
        else:
            r.set_scene(crts_scene_view(scene))
        for f in range(2):
            st = r.render(*args, f == 0, True)
        frames.append((r.read_accum(), st.num_rays))
    assert frames[0][1] == frames[1][1] and np.array_equal(frames[0][0].view(np.uint32), frames[1][0].view(np.uint32))


def test_natively_loaded_gltf_renders_the_same_frame(mods, tmp_path, size=(48, 32), scale=0.02):
    """.gltf + .bin + PNG (the scene class of BASELINE configs 3 and 5: instances with non-identity transforms, textures) ->
    crtio_load_gltf -> crtc_set_scene -> the frame of the scene the file was written from, bit for bit."""
    pytest.importorskip("PIL")
    from chameleonrt_b200 import scene_io
    from chameleonrt_b200.gltf_io import write_gltf
    from chameleonrt_b200.scenes import san_miguel_like

    RenderCUDA = mods[0]
    scene, cam = san_miguel_like(spp=2, scale=scale, tex_size=32)
    c = camera_for(cam)
    args = (c.eye(), c.dir(), c.up(), cam["fov_y"])
    loaded = scene_io.load_scene(write_gltf(scene, str(tmp_path / "scene.gltf")))
    frames = []
    for native in (False, True):
        r = RenderCUDA(0)
        r.initialize(*size)
        if native:
            r.set_scene_c(loaded.c_scene, samples_per_pixel=2)
        else:
            r.set_scene(scene)
        for f in range(2):
            st = r.render(*args, f == 0, True)
        frames.append((r.read_accum(), st.num_rays))
    assert frames[0][1] == frames[1][1] and np.array_equal(frames[0][0].view(np.uint32), frames[1][0].view(np.uint32))


def test_natively_loaded_obj_renders_the_same_frame(mods, tmp_path, size=(48, 32), detail=0.25):
    """The scene-load row end to end: OBJ + MTL + PNG written by obj_io -> crtio_load_obj (native, parallel) ->
    crtc_set_scene on the native crt_scene_t (RenderCUDA.set_scene_c) -> the frame the Python scene model renders, bit for
    bit (on the B200 the same function runs from tests/test_z_new_gpu_paths.py)."""
    from chameleonrt_b200 import scene_io
    from chameleonrt_b200.obj_io import write_obj
    from chameleonrt_b200.scenes import sponza_like

    RenderCUDA = mods[0]
    scene, cam = sponza_like(spp=2, detail=detail, tex_size=32)
    c = camera_for(cam)
    args = (c.eye(), c.dir(), c.up(), cam["fov_y"])
    loaded = scene_io.load_obj(write_obj(scene, str(tmp_path / "scene.obj")))
    frames = []
    for native in (False, True):
        r = RenderCUDA(0)
        r.initialize(*size)
        if native:
            r.set_scene_c(loaded.c_scene, samples_per_pixel=2)
        else:
            r.set_scene(scene)
        for f in range(2):
            st = r.render(*args, f == 0, True)
        frames.append((r.read_accum(), st.num_rays))
    assert frames[0][1] == frames[1][1] and np.array_equal(frames[0][0].view(np.uint32), frames[1][0].view(np.uint32))
