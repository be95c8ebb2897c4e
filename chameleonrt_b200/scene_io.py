"""ctypes binding of the native scene loader (include/crt_scene_io.h, chameleonrt_b200/csrc/libcrt_scene_io.so): the
parallel twin of the reference's ``Scene::load_obj`` (util/scene.cpp:94-228) and ``Scene::load_crts`` (:417-625) — same Scene,
bit for bit, from a memory-mapped file (an OBJ parsed on several threads; a .crts used in place, its images decoded
concurrently). ``LoadedScene.c_scene`` is the ``crt_scene_t`` pointer ``RenderCUDA.set_scene_c`` takes;
``LoadedScene.to_scene()`` copies it into the Python scene model (tests, the CPU oracle)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .scene import (CScene, DisneyMaterial, Geometry, Image, Instance, Mesh, ParameterizedMesh, QuadLight, Scene)

_LIB = None


def _f32_from_bits(bits: int) -> float:
    """The Python float holding the float32 with this bit pattern, built without a float load for denormals (a thread in
    denormals-are-zero mode would read a texture handle 0x80000000 | id as -0)."""
    import math

    if (bits >> 23) & 0xFF == 0:
        v = math.ldexp(bits & 0x7FFFFF, -149)
        return -v if bits & 0x80000000 else v
    return float(np.array([bits], np.uint32).view(np.float32)[0])


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcrt_scene_io.so")


def _lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(lib_path()):
            raise RuntimeError(f"{lib_path()} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        lib = C.CDLL(lib_path())
        for fn in (lib.crtio_load_obj, lib.crtio_load_crts, lib.crtio_load_gltf, lib.crtio_load):
            fn.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.crtio_load_mode.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.crtio_cameras.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_float))]
        lib.crtio_scene_view.restype = C.POINTER(CScene)
        lib.crtio_scene_view.argtypes = [C.c_void_p]
        lib.crtio_timings.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int]
        lib.crtio_warnings.restype = C.c_char_p
        lib.crtio_warnings.argtypes = [C.c_void_p]
        lib.crtio_free.argtypes = [C.c_void_p]
        lib.crtio_last_error.restype = C.c_char_p
        _LIB = lib
    return _LIB


class LoadedScene:
    """A scene owned by the native loader. Keep it alive while ``c_scene`` is in use."""

    def __init__(self, handle):
        self._h = handle
        self.c_scene = _lib().crtio_scene_view(handle)
        t = (C.c_double * 4)()
        _lib().crtio_timings(handle, t, 4)
        self.timings = dict(total_s=t[0], parse_s=t[1], remap_s=t[2], materials_textures_s=t[3])
        self.warnings = _lib().crtio_warnings(handle).decode()
        cams = C.POINTER(C.c_float)()
        n = _lib().crtio_cameras(handle, C.byref(cams))
        raw = np.ctypeslib.as_array(cams, (n, 10)).copy() if n else np.zeros((0, 10), np.float32)
        #: the file's cameras (util/camera.h): dicts of position / center / up (3 floats each) and fov_y
        self.cameras = [dict(position=c[0:3], center=c[3:6], up=c[6:9], fov_y=float(c[9])) for c in raw]

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib().crtio_free(self._h)
                self._h = None
        except Exception:
            pass

    def to_scene(self, spp: int = 1) -> Scene:
        """A copy in the Python scene model (numpy arrays), e.g. for the CPU oracle."""
        s = self.c_scene.contents
        meshes = []
        for m in range(s.num_meshes):
            geoms = []
            for g in range(s.meshes[m].num_geometries):
                cg = s.meshes[m].geometries[g]
                v = np.ctypeslib.as_array(cg.vertices, (cg.num_vertices, 3)).copy()
                uv = np.ctypeslib.as_array(cg.uvs, (cg.num_vertices, 2)).copy() if cg.uvs else None
                idx = np.ctypeslib.as_array(cg.indices, (cg.num_tris, 3)).copy()
                geoms.append(Geometry(v, idx, uv))
            meshes.append(Mesh(geoms))
        pms = []
        for i in range(s.num_parameterized_meshes):
            pm = s.parameterized_meshes[i]
            pms.append(ParameterizedMesh(int(pm.mesh_id), [int(pm.material_ids[k]) for k in range(pm.num_material_ids)]))
        mats = []
        for i in range(s.num_materials):
            w = np.ctypeslib.as_array(C.cast(C.pointer(s.materials[i]), C.POINTER(C.c_uint32)), (16,)).copy()
            f = [_f32_from_bits(int(x)) for x in w]  # (texture handles are bit patterns: denormal-safe, like scene.f32_bits)
            mats.append(DisneyMaterial(base_color=(f[0], f[1], f[2]), metallic=f[3], specular=f[4], roughness=f[5], specular_tint=f[6],
                                       anisotropy=f[7], sheen=f[8], sheen_tint=f[9], clearcoat=f[10], clearcoat_gloss=f[11], ior=f[12],
                                       specular_transmission=f[13]))
        texs = []
        for i in range(s.num_textures):
            im = s.textures[i]
            px = np.ctypeslib.as_array(im.data, (im.height, im.width, im.channels)).copy()
            texs.append(Image(f"tex{i}", px, int(im.color_space)))
        lights = []
        for i in range(s.num_lights):
            lf = [float(x) for x in np.ctypeslib.as_array(C.cast(C.pointer(s.lights[i]), C.POINTER(C.c_float)), (20,))]
            lights.append(QuadLight(emission=tuple(lf[0:4]), position=tuple(lf[4:8]), normal=tuple(lf[8:12]), v_x=tuple(lf[12:15]),
                                    width=lf[15], v_y=tuple(lf[16:19]), height=lf[19]))
        instances = []
        for i in range(s.num_instances):
            xf = np.ctypeslib.as_array(s.instances[i].transform, (16,)).copy().reshape(4, 4).T
            instances.append(Instance(xf.astype(np.float32), int(s.instances[i].parameterized_mesh_id)))
        return Scene(meshes=meshes, parameterized_meshes=pms, instances=instances, materials=mats, textures=texs, lights=lights,
                     samples_per_pixel=spp)


def _load(fn, path, threads) -> LoadedScene:
    h = C.c_void_p()
    if fn(os.fspath(path).encode(), threads, C.byref(h)) != 0:
        raise RuntimeError(_lib().crtio_last_error().decode())
    return LoadedScene(h)


def load_obj(path: str, threads: int = 0) -> LoadedScene:
    """``Scene::load_obj`` (util/scene.cpp:94-228), natively and in parallel. Raises RuntimeError like the reference throws."""
    return _load(_lib().crtio_load_obj, path, threads)


def load_crts(path: str, threads: int = 0) -> LoadedScene:
    """``Scene::load_crts`` (util/scene.cpp:417-625): geometry arrays used in place in the mapped file, images decoded in parallel."""
    return _load(_lib().crtio_load_crts, path, threads)


def load_gltf(path: str, threads: int = 0) -> LoadedScene:
    """``Scene::load_gltf`` (util/scene.cpp:230-415): .gltf / .glb, packed accessors used in place, the scene graph flattened."""
    return _load(_lib().crtio_load_gltf, path, threads)


def load_scene(path: str, threads: int = 0, white_diffuse: bool = False) -> LoadedScene:
    """``Scene::Scene`` (util/scene.cpp:49-67): the loader the file's extension names (obj, gltf, glb, crts). ``white_diffuse``:
    ``MaterialMode::WHITE_DIFFUSE`` (main.cpp's ``-mat-mode white_diffuse``) — no materials, one default material for everything."""
    h = C.c_void_p()
    if _lib().crtio_load_mode(os.fspath(path).encode(), threads, 1 if white_diffuse else 0, C.byref(h)) != 0:
        raise RuntimeError(_lib().crtio_last_error().decode())
    return LoadedScene(h)
