"""The kernels and the renderer under AddressSanitizer + UndefinedBehaviorSanitizer (or ThreadSanitizer), on the CPU SIMT emulation
(tests/simt_emu: every CUDA thread an OS thread, shared memory and atomics real): a sanitized build of the emulated core renders a
few small frames through the C ABI with every option that changes the kernel chain. An out-of-bounds access to a queue, a
shared-memory array or a scene buffer, a signed overflow or a misaligned access in kernel code stops the run with a report.
What compute-sanitizer would check on the device, checked where no GPU is available.

    python scripts/simt_sanitize.py [--tsan]
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build(tsan: bool) -> str:
    from simt_emu import build as simt_build

    simt_build.build()  # (generates the translated sources under tests/simt_emu/_build)
    out = os.path.join(simt_build.OUT, "libcrt_cuda_core_simt_tsan.so" if tsan else "libcrt_cuda_core_simt_asan.so")
    csrc = simt_build.CSRC
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"))), "include")
    san = ["-fsanitize=thread"] if tsan else ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-pthread", "-march=x86-64-v3", "-ffp-contract=off", "-Wno-attributes",
                           "-fno-omit-frame-pointer"] + san + ["-I" + csrc, "-I" + cuda_inc, "-I" + os.path.join(ROOT, "include"), "-shared",
                           "-Wl,-Bsymbolic", "-o", out, os.path.join(simt_build.OUT, "crt_cuda_core_simt.cpp"),
                           os.path.join(simt_build.HERE, "cuda_emu.cpp"), os.path.join(csrc, "host_scene.cpp"), os.path.join(csrc, "bvh8_build.cpp")])
    return out


def render_all(lib: str):
    import numpy as np

    from chameleonrt_b200 import ArcballCamera, backend
    from chameleonrt_b200.scenes import cornell_box, sponza_like
    from helpers import synthetic_material_scene

    backend._LIB_PATH, backend._lib = lib, None
    cases = [("cornell", cornell_box(spp=1), {}), ("materials", synthetic_material_scene(spp=1), {}),
             ("sponza, device PLOC build, shade_sort", sponza_like(spp=1, detail=0.15, tex_size=16), {"bvh_builder": 1, "shade_sort": 2}),
             ("sponza, LBVH, far-first, no defer, top-of-tree in shared memory", sponza_like(spp=1, detail=0.15, tex_size=16),
              {"bvh_builder": 2, "any_far_first": 1, "tri_pass_defer": 0, "bvh_top_smem": 1}),
             ("cornell, tile shard 1 of 2 (130 x 70: six tiles)", cornell_box(spp=1), {"world_size": 2, "rank": 1, "size": (130, 70)})]
    for name, (scene, cam), options in cases:
        r = backend.RenderCUDA(0, max_depth=4)
        size = options.pop("size", (40, 24))
        for k, v in options.items():
            r.set_option(k, v)
        r.initialize(*size)
        r.set_scene(scene)
        c = ArcballCamera(cam["eye"], cam["center"], cam["up"])
        for f in range(2):
            st = r.render(c.eye(), c.dir(), c.up(), cam["fov_y"], f == 0, True)
        a = r.read_accum()
        assert np.isfinite(a).all() and st.num_rays > 0  # (the materials scene has one pixel at -1.7e9: the reference's own value)
        print(f"{name}: {st.num_rays} rays, mean {float(a.mean()):.4f}", flush=True)
        del r
    # frames in flight, the instrumented traversal kernels, stage events off, a resize, the traversal API
    scene, cam = cornell_box(spp=2)
    c = ArcballCamera(cam["eye"], cam["center"], cam["up"])
    args = (c.eye(), c.dir(), c.up(), cam["fov_y"])
    r = backend.RenderCUDA(0, max_depth=3)
    r.set_option("count_traversal", 1)
    r.set_option("stage_events", 0)
    r.initialize(33, 17)
    r.set_scene(scene)
    r.render_async(*args, True, 3)
    r.sync()
    r.initialize(70, 66)
    r.set_scene(scene)
    r.render(*args, True, True)
    from oracle.oracle import primary_rays

    rays = primary_rays(16, 12, c.eye(), c.dir(), c.up(), cam["fov_y"])
    hits = r.trace_closest(rays)
    occluded = r.trace_any(rays)
    print(f"async + counting + resize + traversal API: {r.counters()['closest_nodes_visited']} nodes visited, "
          f"{int(np.isfinite(hits[:, 0]).sum())} of {len(rays)} primary rays hit, {int(occluded.sum())} occluded", flush=True)
    # two renderers of one process sharing rank 0's frame (peer stores + completion flags)
    ranks = [backend.RenderCUDA(0, max_depth=3, rank=k, world_size=2) for k in range(2)]
    for rr in ranks:
        rr.initialize(130, 70)
        rr.set_scene(scene)
    ranks[1].import_frame(ranks[0].export_frame())
    for rr in ranks:  # (kernels run synchronously here: the order tests/test_simt_renderer.py uses)
        rr.render_async(*args, True, 2)
        rr.render_async(*args, False, 1)
        rr.sync()
    full = ranks[0].read_accum()
    lit = (full.reshape(-1, 3).sum(axis=1) > 0).reshape(full.shape[0], full.shape[1])
    assert np.isfinite(full).all() and lit[:, :64].mean() > 0.5 and lit[:, 64:128].mean() > 0.5  # tiles of both ranks arrived
    print(f"shared frame of two renderers: mean {float(full.mean()):.4f}, lit {lit.mean():.2f}", flush=True)


def adversarial_geometry():
    """NaN / infinite / huge / tiny coordinates, degenerate, identical, duplicated and collinear triangles, a single triangle:
    through set_scene (host builder, PLOC, LBVH) and a frame. Every builder must cope, and cast the same number of rays."""
    import numpy as np

    from chameleonrt_b200 import ArcballCamera, backend
    from chameleonrt_b200.scene import DisneyMaterial, Geometry, Instance, Mesh, ParameterizedMesh, Scene, default_obj_light

    rng = np.random.default_rng(0)

    def scene_of(verts, idx):
        g = Geometry(np.asarray(verts, np.float32), np.asarray(idx, np.uint32), None)
        return Scene(meshes=[Mesh([g])], parameterized_meshes=[ParameterizedMesh(0, [0])], instances=[Instance(np.eye(4, dtype=np.float32), 0)],
                     materials=[DisneyMaterial(base_color=(0.8, 0.8, 0.8))], textures=[], lights=[default_obj_light()], samples_per_pixel=1)

    v0, i0 = rng.normal(size=(60, 3)), rng.integers(0, 60, (100, 3))
    with_nan, with_inf = v0.copy(), v0.copy()
    with_nan[3] = np.nan
    with_inf[5], with_inf[6] = np.inf, -np.inf
    cases = {"nan vertex": (with_nan, i0), "infinite vertices": (with_inf, i0), "coordinates x 1e30": (v0 * 1e30, i0), "coordinates x 1e-30": (v0 * 1e-30, i0),
             "degenerate triangles": (v0, np.repeat(rng.integers(0, 60, (100, 1)), 3, axis=1)), "identical vertices": (np.zeros((60, 3)), i0),
             "single triangle": (v0[:3], [[0, 1, 2]]), "50 copies of two triangles": (v0, np.tile(i0[:2], (50, 1))),
             "collinear vertices": (np.outer(np.linspace(0, 1, 60), [1, 2, 3]), i0)}
    c = ArcballCamera((0, 0, 5), (0, 0, 0), (0, 1, 0))
    for name, (v, i) in cases.items():
        rays = []
        for builder in (0, 1, 2):
            r = backend.RenderCUDA(0, max_depth=3)
            r.set_option("bvh_builder", builder)
            r.initialize(24, 16)
            r.set_scene(scene_of(v, i))
            rays.append(r.render(c.eye(), c.dir(), c.up(), 45.0, True, True).num_rays)
        assert rays[0] == rays[1] == rays[2], (name, rays)
        print(f"adversarial geometry, {name}: {rays[0]} rays with every builder", flush=True)


if __name__ == "__main__":
    tsan = "--tsan" in sys.argv
    if os.environ.get("CRT_SANITIZED_CHILD") != "1":
        lib = build(tsan)
        runtime = subprocess.check_output(["g++", "-print-file-name=" + ("libtsan.so" if tsan else "libasan.so")], text=True).strip()
        env = dict(os.environ, CRT_SANITIZED_CHILD="1", CRT_SANITIZED_LIB=lib, LD_PRELOAD=runtime,
                   ASAN_OPTIONS="detect_leaks=0:abort_on_error=0:halt_on_error=1", TSAN_OPTIONS="halt_on_error=0:report_signal_unsafe=0")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))
    render_all(os.environ["CRT_SANITIZED_LIB"])
    adversarial_geometry()
    print("sanitized run finished")
