"""ctypes binding of liboracle.so — the CPU restatement of the Embree/ISPC backend."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from chameleonrt_b200.scene import CRenderStats, CScene, RenderStats, Scene

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def load_oracle_lib(fast: bool = False) -> C.CDLL:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _LIBS:
        return _LIBS[name]
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    fp = C.POINTER(C.c_float)
    lib.oracle_create.restype = C.c_void_p
    lib.oracle_destroy.argtypes = [C.c_void_p]
    lib.oracle_set_options.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.oracle_initialize.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.oracle_set_pixel_window.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    lib.oracle_set_scene.argtypes = [C.c_void_p, C.POINTER(CScene)]
    lib.oracle_render.argtypes = [C.c_void_p, fp, fp, fp, C.c_float, C.c_int, C.POINTER(CRenderStats)]
    lib.oracle_read_img.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_read_accum.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_read_ray_stats.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.oracle_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.oracle_primary_rays.argtypes = [C.c_int, C.c_int, fp, fp, fp, C.c_float, C.c_uint32, C.c_uint32,
                                        C.c_uint32, C.c_void_p]
    lib.oracle_kat_rng.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.oracle_kat_camera.argtypes = [fp, fp, fp, C.c_float, C.c_int, C.c_int, C.c_void_p]
    lib.oracle_kat_disney_eval.argtypes = [C.c_void_p] * 5
    lib.oracle_kat_disney_sample.argtypes = [C.c_void_p] * 5
    lib.oracle_kat_light.argtypes = [C.c_void_p] * 5
    lib.oracle_kat_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_kat_miss.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_kat_ortho_basis.argtypes = [C.c_void_p, C.c_void_p]
    lib.oracle_kat_srgb8.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    lib.oracle_kat_tri.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.oracle_hardware_threads.restype = C.c_int
    _LIBS[name] = lib
    return lib


def _vec3(v):
    a = np.ascontiguousarray(v, dtype=np.float32).reshape(3)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class OracleBackend:
    """Same surface as RenderBackend (util/render_backend.h:12-32)."""

    def __init__(self, max_depth: int = 5, num_threads: int = 0, brute_force: bool = False, fast: bool = False):
        self.lib = load_oracle_lib(fast)
        self.h = C.c_void_p(self.lib.oracle_create())
        self.lib.oracle_set_options(self.h, max_depth, num_threads, 1 if brute_force else 0)
        self.width = self.height = 0
        self.samples_per_pixel = 1
        self.img = None

    def __del__(self):
        try:
            if self.h:
                self.lib.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def name(self) -> str:
        return "CPU oracle (Embree/ISPC backend restated; own BVH2)"

    def initialize(self, fb_width: int, fb_height: int) -> None:
        self.width, self.height = fb_width, fb_height
        self.lib.oracle_initialize(self.h, fb_width, fb_height)
        self.img = np.zeros((fb_height, fb_width), dtype=np.uint32)

    def set_scene(self, scene: Scene) -> None:
        ms = scene.to_c()
        self.samples_per_pixel = scene.samples_per_pixel
        self.lib.oracle_set_scene(self.h, C.byref(ms.c))

    def set_pixel_window(self, x0: int, y0: int, x1: int, y1: int) -> None:
        """Debugging aid: render only the pixels [x0, x1) x [y0, y1) from now on."""
        self.lib.oracle_set_pixel_window(self.h, x0, y0, x1, y1)

    def render(self, pos, dir, up, fovy, camera_changed, readback_framebuffer=True) -> RenderStats:
        _p, pp = _vec3(pos)
        _d, dp = _vec3(dir)
        _u, up_ = _vec3(up)
        st = CRenderStats()
        self.lib.oracle_render(self.h, pp, dp, up_, C.c_float(fovy), 1 if camera_changed else 0, C.byref(st))
        self.lib.oracle_read_img(self.h, self.img.ctypes.data)
        return RenderStats(st.render_time, st.rays_per_second, st.num_rays)

    def read_accum(self) -> np.ndarray:
        out = np.zeros((self.height, self.width, 3), dtype=np.float32)
        self.lib.oracle_read_accum(self.h, out.ctypes.data)
        return out

    def read_ray_stats(self) -> np.ndarray:
        out = np.zeros((self.height, self.width), dtype=np.uint16)
        self.lib.oracle_read_ray_stats(self.h, out.ctypes.data)
        return out

    def trace_closest(self, rays: np.ndarray, want_normals: bool = False):
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        n = rays.shape[0]
        hits = np.zeros((n, 4), dtype=np.float32)
        normals = np.zeros((n, 3), dtype=np.float32) if want_normals else None
        self.lib.oracle_trace_closest(self.h, rays.ctypes.data, n, hits.ctypes.data,
                                      normals.ctypes.data if want_normals else None)
        return (hits, normals) if want_normals else hits

    def trace_any(self, rays: np.ndarray) -> np.ndarray:
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
        occ = np.zeros(rays.shape[0], dtype=np.uint8)
        self.lib.oracle_trace_any(self.h, rays.ctypes.data, rays.shape[0], occ.ctypes.data)
        return occ


def primary_rays(w, h, pos, dir, up, fovy, frame_id=0, spp=1, s=0, fast=False) -> np.ndarray:
    lib = load_oracle_lib(fast)
    _p, pp = _vec3(pos)
    _d, dp = _vec3(dir)
    _u, up_ = _vec3(up)
    rays = np.zeros((h * w, 8), dtype=np.float32)
    lib.oracle_primary_rays(w, h, pp, dp, up_, C.c_float(fovy), frame_id, spp, s, rays.ctypes.data)
    return rays
