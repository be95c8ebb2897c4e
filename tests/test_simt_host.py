"""k_traverse — the persistent, warp-cooperative traversal kernel — executed on the host under a SIMT emulation
(TEST-ONLY libcrt_simt_hostcheck.so: one OS thread per CUDA thread, the 32 threads of a warp rendezvous at every
__ballot_sync / __shfl_*_sync / __syncwarp, block-shared memory and atomics are real). Its results must equal, ray by
ray and bit by bit, the single-ray host instantiation of bvh8_traverse.h (which the other tests pin against brute
force and the oracle): dynamic ray fetch and refill, the shared-memory short stack with local-memory spill, the
warp-pooled triangle tests merged by a 64-bit atomicMin, the merged shadow + closest launch, queue indirection,
partially filled warps, the far-first shadow order."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from chameleonrt_b200.scene import CScene
from helpers import HostCheck, camera_for

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


@pytest.fixture(scope="module")
def lib(built):
    lib = C.CDLL(os.path.join(ROOT, "chameleonrt_b200", "csrc", "libcrt_simt_hostcheck.so"))
    lib.crt_simt_create.restype = C.c_void_p
    lib.crt_simt_create.argtypes = [C.POINTER(CScene)]
    lib.crt_simt_destroy.argtypes = [C.c_void_p]
    lib.crt_simt_traverse.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p]
    return lib


def _rays(scene, cam, w, h, seed):
    from bvh_quality import cosine_bounce, shadow_rays
    from oracle import OracleBackend
    from oracle.oracle import primary_rays

    c = camera_for(cam)
    o = OracleBackend(fast=True)
    o.initialize(8, 8)
    o.set_scene(scene)
    rng = np.random.default_rng(seed)
    rays = primary_rays(w, h, c.eye(), c.dir(), c.up(), cam["fov_y"])
    hits, normals = o.trace_closest(rays, True)
    b1, p = cosine_bounce(rays, hits, normals, rng)
    return np.ascontiguousarray(np.concatenate([rays, b1]), np.float32), shadow_rays(p, scene.lights[0], rng)


def _run(lib, h, closest, shadow, sched, perm=None, blocks=1):
    out, vis = np.zeros((len(closest), 4), np.float32), np.zeros(max(1, len(shadow)), np.uint8)
    q = None if perm is None else np.ascontiguousarray(perm, np.uint32)
    lib.crt_simt_traverse(h, closest.ctypes.data, len(closest), None if q is None else q.ctypes.data, shadow.ctypes.data, len(shadow),
                          blocks, sched, out.ctypes.data, vis.ctypes.data)
    return out, vis[: len(shadow)]


@pytest.mark.parametrize("name", ["sponza", "cornell"])
def test_k_traverse_under_simt_emulation_equals_single_ray_traversal(lib, name):
    from chameleonrt_b200.scenes import cornell_box, sponza_like

    scene, cam = sponza_like(spp=1, detail=0.3, tex_size=16) if name == "sponza" else cornell_box()
    ms = scene.to_c()
    h = lib.crt_simt_create(C.byref(ms.c))
    assert h
    closest, shadow = _rays(scene, cam, 40, 24, 5)
    shadow[:, 3] = 1e-4  # the kernel starts shadow rays at kEpsilon
    hc = HostCheck(scene)
    want, _, _ = hc.trace(closest)
    occluded = hc.trace(shadow, any_hit=True)[0][:, 3].view(np.uint32) != 0xFFFFFFFF
    # default scheduling, far-first shadow rays, refill as soon as one lane idles, refill only when the warp is empty,
    # and the deferred-triangle-pass instantiations (option tri_pass_defer = 16 / 24), alone and with far-first / refill 1
    for sched in (4, 4 | 0x100, 1, 32, 4 | (16 << 16), 4 | (24 << 16), 1 | 0x100 | (16 << 16), 32 | (24 << 16)):
        got, vis = _run(lib, h, closest, shadow, sched)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), hex(sched)
        assert np.array_equal(vis == 0, occluded) and set(np.unique(vis)) <= {0, 1}
    # closest rays only / shadow rays only / a ragged handful (partially filled warp) / nothing at all
    got, _ = _run(lib, h, closest, shadow[:0], 4)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    _, vis = _run(lib, h, closest[:0], shadow, 4)
    assert np.array_equal(vis == 0, occluded)
    for sched in (4, 4 | (16 << 16)):
        got, vis = _run(lib, h, closest[:37], shadow[:5], sched)
        assert np.array_equal(got.view(np.uint32), want[:37].view(np.uint32)) and np.array_equal(vis == 0, occluded[:5])
    _run(lib, h, closest[:0], shadow[:0], 4)
    # queue indirection (the compacted queue of a later bounce): slots visited in a permuted order, two blocks
    perm = np.random.default_rng(1).permutation(len(closest)).astype(np.uint32)
    for sched in (4, 4 | (24 << 16)):
        got, vis = _run(lib, h, closest, shadow, sched, perm, blocks=2)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) and np.array_equal(vis == 0, occluded)
    lib.crt_simt_destroy(h)


@pytest.mark.parametrize("n,capacity,materials", [(0, 1, 3), (1, 1, 1), (31, 2048, 2), (2048, 2048, 5), (5000, 9000, 300),
                                                  (4097, 4097, 255)])
def test_shade_queue_sort_kernels_are_a_stable_sort_by_material_bucket(lib, n, capacity, materials):
    """k_queue_hist + k_queue_scatter (option shade_sort) under the SIMT emulation: the sorted queue is the stable
    sort of the input queue by bucket (material id, ids >= 254 share bucket 254, misses are bucket 255); entries
    beyond the device-side length are neither counted nor written."""
    rng = np.random.default_rng(n + 7 * materials)
    num_slots, num_tris = capacity + 5, 97
    tri_material = rng.integers(0, materials, num_tris).astype(np.uint32)
    hit_tri = rng.integers(0, num_tris, num_slots).astype(np.uint32)
    hit_tri[rng.random(num_slots) < 0.2] = 0xFFFFFFFF
    queue = rng.permutation(num_slots)[:capacity].astype(np.uint32)  # entries [n, capacity) are stale
    out = np.zeros(capacity, np.uint32)
    lib.crt_simt_sort_queue.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    lib.crt_simt_sort_queue.restype = None
    lib.crt_simt_last_error.restype = C.c_char_p
    lib.crt_simt_sort_queue(hit_tri.ctypes.data, num_slots, tri_material.ctypes.data, num_tris, queue.ctypes.data, n, capacity,
                            out.ctypes.data)
    assert not lib.crt_simt_last_error(), lib.crt_simt_last_error()
    live = queue[:n]
    tri = hit_tri[live]
    bucket = np.where(tri == 0xFFFFFFFF, 255, np.minimum(tri_material[np.minimum(tri, num_tris - 1)], 254))
    expect = live[np.argsort(bucket, kind="stable")]
    assert (out[:n] == expect).all()
    assert (out[n:] == 0xFFFFFFFF).all()

