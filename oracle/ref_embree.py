"""ctypes binding of oracle/_ref/libcrt_embree.so — the REFERENCE'S OWN Embree/ISPC backend.

TEST INFRASTRUCTURE. The library is compiled by oracle/ref_build/Makefile from
/root/reference/backends/embree/{render_embree.cpp,embree_utils.cpp,render_embree.ispc,*.ih} where
they lie (the ISPC kernels as scalar C++, Embree/TBB/GLM replaced by third_party/ stand-ins), so it
exists only where /root/reference was present at build time (this container; the GPU box gets the
prebuilt file). It is the pin for the oracle: tests compare the oracle's frames with this library's
frames, and tests/golden/ref_embree_*.npz hold frames generated from it (make_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from chameleonrt_b200.scene import CScene, RenderStats, Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcrt_embree.so")  # strict build (parity tests)
FAST_LIB_PATH = os.path.join(_HERE, "_ref", "libcrt_embree_fast.so")  # -O3, contraction allowed (timing)
_LIBS = {}


def available(fast: bool = False) -> bool:
    return os.path.exists(FAST_LIB_PATH if fast else LIB_PATH)


def load_ref_embree_lib(fast: bool = False) -> C.CDLL:
    path = FAST_LIB_PATH if fast else LIB_PATH
    if path in _LIBS:
        return _LIBS[path]
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is not built (needs /root/reference: make -C oracle/ref_build)")
    lib = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    lib.refembree_create.restype = C.c_void_p
    lib.refembree_destroy.argtypes = [C.c_void_p]
    lib.refembree_name.argtypes = [C.c_void_p]
    lib.refembree_name.restype = C.c_char_p
    lib.refembree_initialize.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.refembree_set_scene.argtypes = [C.c_void_p, C.POINTER(CScene)]
    lib.refembree_render.argtypes = [C.c_void_p, fp, fp, fp, C.c_float, C.c_int, fp]
    lib.refembree_render.restype = C.c_uint64
    lib.refembree_read_accum.argtypes = [C.c_void_p, C.c_void_p]
    lib.refembree_read_img.argtypes = [C.c_void_p, C.c_void_p]
    lib.refembree_read_ray_stats.argtypes = [C.c_void_p, C.c_void_p]
    # known-answer wrappers around the reference's own pure functions (ispc_cpp/epilogue.h); same
    # signatures as the oracle's oracle_kat_*
    vp = C.c_void_p
    lib.refispc_kat_rng.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
    lib.refispc_kat_disney_eval.argtypes = [vp] * 5
    lib.refispc_kat_disney_sample.argtypes = [vp] * 5
    lib.refispc_kat_light.argtypes = [vp] * 5
    lib.refispc_kat_texture.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp]
    lib.refispc_kat_texture_channel.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp]
    lib.refispc_kat_miss.argtypes = [vp, C.c_int, vp]
    lib.refispc_kat_ortho_basis.argtypes = [vp, vp]
    lib.refembree_set_max_path_depth.argtypes = [C.c_int]
    lib.refembree_get_max_path_depth.restype = C.c_int
    _LIBS[path] = lib
    return lib


def _vec3(v):
    a = np.ascontiguousarray(v, dtype=np.float32).reshape(3)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class RefEmbreeBackend:
    """RenderEmbree (backends/embree/render_embree.h:11-44) behind the RenderBackend surface.
    ``max_depth`` defaults to the reference's MAX_PATH_DEPTH = 5 (util.ih:10); it is a per-library
    global (the reference has it as a compile-time constant), set again before every render."""

    def __init__(self, max_depth: int = 5, fast: bool = False):
        self.lib = load_ref_embree_lib(fast)
        self.max_depth = max_depth
        self.h = C.c_void_p(self.lib.refembree_create())
        self.width = self.height = 0
        self.img = None

    def __del__(self):
        try:
            if self.h:
                self.lib.refembree_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def name(self) -> str:
        return self.lib.refembree_name(self.h).decode()

    def initialize(self, fb_width: int, fb_height: int) -> None:
        self.width, self.height = fb_width, fb_height
        self.lib.refembree_initialize(self.h, fb_width, fb_height)
        self.img = np.zeros((fb_height, fb_width), dtype=np.uint32)

    def set_scene(self, scene: Scene) -> None:
        ms = scene.to_c()
        self.lib.refembree_set_scene(self.h, C.byref(ms.c))

    def render(self, pos, dir, up, fovy, camera_changed, readback_framebuffer=True) -> RenderStats:
        _p, pp = _vec3(pos)
        _d, dp = _vec3(dir)
        _u, up_ = _vec3(up)
        ms = C.c_float(0)
        self.lib.refembree_set_max_path_depth(self.max_depth)
        rays = self.lib.refembree_render(self.h, pp, dp, up_, C.c_float(fovy), 1 if camera_changed else 0, C.byref(ms))
        self.lib.refembree_read_img(self.h, self.img.ctypes.data)
        t = float(ms.value)
        return RenderStats(t, rays / (t * 1e-3) if t > 0 else 0.0, int(rays))

    def read_accum(self) -> np.ndarray:
        out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        self.lib.refembree_read_accum(self.h, out.ctypes.data)
        return out

    def read_ray_stats(self) -> np.ndarray:
        out = np.zeros((self.height, self.width), dtype=np.uint16)
        self.lib.refembree_read_ray_stats(self.h, out.ctypes.data)
        return out
