"""Backend-neutral scene model: the Python mirror of ChameleonRT's ``Scene``.

Names and field meaning follow the reference's ``util/`` headers so that host code and
tests read like the reference:

* ``Geometry``            util/mesh.h:6-12
* ``Mesh``                util/mesh.h:14-22
* ``ParameterizedMesh``   util/mesh.h:28-36
* ``Instance``            util/mesh.h:40-47 (``transform`` is a 4x4 object_to_world)
* ``DisneyMaterial``      util/material.h:29-46
* ``Image``               util/material.h:11-27
* ``QuadLight``           util/lights.h:6-18
* ``Camera``              util/camera.h:5-8
* ``Scene``               util/scene.h:23-32 (+ ``validate_materials`` scene.cpp:935-958,
  the generated OBJ light scene.cpp:218-227)

``Scene.to_c()`` marshals everything into the plain-C ``crt_scene_t`` of
``include/crt_scene.h``, which is what both the CUDA backend's C ABI and the CPU oracle take.
"""
from __future__ import annotations

import ctypes as C
import math
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

LINEAR = 0
SRGB = 1

TEXTURED_PARAM_MASK = 0x80000000


def textured_param(tex_id: int, channel: int = 0) -> float:
    """Encode a texture handle into a material float (util/texture_channel_mask.h:16-23)."""
    mask = TEXTURED_PARAM_MASK | ((channel & 0x3) << 29) | (tex_id & 0x1FFFFFFF)
    if (mask >> 23) & 0xFF == 0:
        # a float32 denormal: build the value by integer arithmetic, a float->double load would read
        # it as zero on a thread in denormals-are-zero mode (see f32_bits)
        return -math.ldexp(mask & 0x7FFFFF, -149)
    return struct.unpack("<f", struct.pack("<I", mask))[0]


def f32_bits(v: float) -> int:
    """IEEE-754 binary32 bit pattern of ``v`` (a Python float holding a float32 value), computed so that
    it does not depend on the thread's flush-to-zero / denormals-are-zero mode: texture handles with a
    small id are float32 DENORMALS (0x80000000 | id), and a double->float conversion under FTZ (the
    reference's Embree backend switches it on, render_embree.cpp:21-24) would flush them to -0."""
    v = float(v)
    a = abs(v)
    if a != 0.0 and a < 2.0 ** -126:  # float32 denormal: integer multiple of 2^-149, exact in double
        sign = 0x80000000 if math.copysign(1.0, v) < 0 else 0
        return sign | int(round(a * 2.0 ** 149))
    return int(np.float32(v).view(np.uint32))


@dataclass
class Geometry:
    vertices: np.ndarray  # (nv, 3) f32
    indices: np.ndarray  # (nt, 3) u32
    uvs: Optional[np.ndarray] = None  # (nv, 2) f32
    normals: Optional[np.ndarray] = None  # carried for fidelity; no kernel reads them

    def num_tris(self) -> int:
        return int(self.indices.shape[0])


@dataclass
class Mesh:
    geometries: List[Geometry] = field(default_factory=list)

    def num_tris(self) -> int:
        return sum(g.num_tris() for g in self.geometries)


@dataclass
class ParameterizedMesh:
    mesh_id: int
    material_ids: List[int]


@dataclass
class Instance:
    transform: np.ndarray  # (4,4) f32, math convention (row i, col j); object_to_world
    parameterized_mesh_id: int


@dataclass
class Image:
    name: str
    img: np.ndarray  # (h, w, channels) u8, row 0 first
    color_space: int = LINEAR

    @property
    def width(self) -> int:
        return int(self.img.shape[1])

    @property
    def height(self) -> int:
        return int(self.img.shape[0])

    @property
    def channels(self) -> int:
        return int(self.img.shape[2])


@dataclass
class DisneyMaterial:
    base_color: tuple = (0.9, 0.9, 0.9)
    metallic: float = 0.0
    specular: float = 0.0
    roughness: float = 1.0
    specular_tint: float = 0.0
    anisotropy: float = 0.0
    sheen: float = 0.0
    sheen_tint: float = 0.0
    clearcoat: float = 0.0
    clearcoat_gloss: float = 0.0
    ior: float = 1.5
    specular_transmission: float = 0.0

    def as_floats(self) -> List[float]:
        return [
            self.base_color[0], self.base_color[1], self.base_color[2], self.metallic,
            self.specular, self.roughness, self.specular_tint, self.anisotropy,
            self.sheen, self.sheen_tint, self.clearcoat, self.clearcoat_gloss,
            self.ior, self.specular_transmission, 0.0, 0.0,
        ]


@dataclass
class QuadLight:
    emission: tuple
    position: tuple
    normal: tuple
    v_x: tuple
    width: float
    v_y: tuple
    height: float

    def as_floats(self) -> List[float]:
        e, p, n = self.emission, self.position, self.normal
        return [
            e[0], e[1], e[2], e[3] if len(e) > 3 else e[0],
            p[0], p[1], p[2], p[3] if len(p) > 3 else 0.0,
            n[0], n[1], n[2], 0.0,
            self.v_x[0], self.v_x[1], self.v_x[2], self.width,
            self.v_y[0], self.v_y[1], self.v_y[2], self.height,
        ]


@dataclass
class Camera:
    position: tuple
    center: tuple
    up: tuple
    fov_y: float


def _f32(x):
    return np.float32(x)


def _normalize3(v):
    """glm::normalize in f32: v * (1/sqrt(dot(v,v)))."""
    v = np.asarray(v, dtype=np.float32)
    d = _f32(_f32(_f32(v[0] * v[0]) + _f32(v[1] * v[1])) + _f32(v[2] * v[2]))
    inv = _f32(1.0) / np.sqrt(d, dtype=np.float32)
    return (v * inv).astype(np.float32)


def _cross3(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return np.array(
        [
            _f32(a[1] * b[2]) - _f32(a[2] * b[1]),
            _f32(a[2] * b[0]) - _f32(a[0] * b[2]),
            _f32(a[0] * b[1]) - _f32(a[1] * b[0]),
        ],
        dtype=np.float32,
    )


def ortho_basis(n):
    """util/util.cpp:43-58."""
    n = np.asarray(n, dtype=np.float32)
    v_y = np.zeros(3, dtype=np.float32)
    if -0.6 < n[0] < 0.6:
        v_y[0] = 1.0
    elif -0.6 < n[1] < 0.6:
        v_y[1] = 1.0
    elif -0.6 < n[2] < 0.6:
        v_y[2] = 1.0
    else:
        v_y[0] = 1.0
    v_x = _normalize3(_cross3(v_y, n))
    v_y = _normalize3(_cross3(n, v_x))
    return v_x, v_y


def default_obj_light() -> QuadLight:
    """The light ``load_obj`` synthesises for every OBJ scene (util/scene.cpp:218-227)."""
    n = _normalize3(np.array([0.5, -0.8, -0.5], dtype=np.float32))
    pos = (np.float32(-10.0) * n).astype(np.float32)
    v_x, v_y = ortho_basis(n)
    return QuadLight(
        emission=(20.0, 20.0, 20.0, 20.0),
        position=(float(pos[0]), float(pos[1]), float(pos[2]), -0.0),
        normal=(float(n[0]), float(n[1]), float(n[2])),
        v_x=tuple(float(x) for x in v_x),
        width=5.0,
        v_y=tuple(float(x) for x in v_y),
        height=5.0,
    )


# ---------------------------------------------------------------------------------------
# ctypes mirror of include/crt_scene.h
# ---------------------------------------------------------------------------------------
class CGeometry(C.Structure):
    _fields_ = [
        ("vertices", C.POINTER(C.c_float)),
        ("uvs", C.POINTER(C.c_float)),
        ("indices", C.POINTER(C.c_uint32)),
        ("num_vertices", C.c_uint32),
        ("num_tris", C.c_uint32),
    ]


class CMesh(C.Structure):
    _fields_ = [("geometries", C.POINTER(CGeometry)), ("num_geometries", C.c_uint32)]


class CParameterizedMesh(C.Structure):
    _fields_ = [
        ("material_ids", C.POINTER(C.c_uint32)),
        ("num_material_ids", C.c_uint32),
        ("mesh_id", C.c_uint32),
    ]


class CInstance(C.Structure):
    _fields_ = [("transform", C.c_float * 16), ("parameterized_mesh_id", C.c_uint32)]


class CMaterial(C.Structure):
    _fields_ = [("p", C.c_float * 16)]


class CImage(C.Structure):
    _fields_ = [
        ("data", C.POINTER(C.c_uint8)),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("channels", C.c_int32),
        ("color_space", C.c_int32),
    ]


class CQuadLight(C.Structure):
    _fields_ = [("p", C.c_float * 20)]


class CScene(C.Structure):
    _fields_ = [
        ("meshes", C.POINTER(CMesh)),
        ("parameterized_meshes", C.POINTER(CParameterizedMesh)),
        ("instances", C.POINTER(CInstance)),
        ("materials", C.POINTER(CMaterial)),
        ("textures", C.POINTER(CImage)),
        ("lights", C.POINTER(CQuadLight)),
        ("num_meshes", C.c_uint32),
        ("num_parameterized_meshes", C.c_uint32),
        ("num_instances", C.c_uint32),
        ("num_materials", C.c_uint32),
        ("num_textures", C.c_uint32),
        ("num_lights", C.c_uint32),
        ("samples_per_pixel", C.c_uint32),
    ]


class CRenderStats(C.Structure):
    _fields_ = [
        ("render_time", C.c_float),
        ("rays_per_second", C.c_float),
        ("num_rays", C.c_uint64),
    ]


@dataclass
class RenderStats:
    """util/render_backend.h:7-10 (+ the ray count REPORT_RAY_STATS would sum)."""

    render_time: float = 0.0
    rays_per_second: float = 0.0
    num_rays: int = 0


class MarshalledScene:
    """Owns every buffer referenced by a ``crt_scene_t`` for as long as it is alive."""

    def __init__(self):
        self.keep = []
        self.c = CScene()


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@dataclass
class Scene:
    meshes: List[Mesh] = field(default_factory=list)
    parameterized_meshes: List[ParameterizedMesh] = field(default_factory=list)
    instances: List[Instance] = field(default_factory=list)
    materials: List[DisneyMaterial] = field(default_factory=list)
    textures: List[Image] = field(default_factory=list)
    lights: List[QuadLight] = field(default_factory=list)
    cameras: List[Camera] = field(default_factory=list)
    samples_per_pixel: int = 1

    def unique_tris(self) -> int:
        return sum(m.num_tris() for m in self.meshes)

    def total_tris(self) -> int:
        return sum(
            self.meshes[self.parameterized_meshes[i.parameterized_mesh_id].mesh_id].num_tris()
            for i in self.instances
        )

    def num_geometries(self) -> int:
        return sum(len(m.geometries) for m in self.meshes)

    def validate_materials(self) -> None:
        """util/scene.cpp:935-958: material id -1 -> an appended default DisneyMaterial."""
        need = any(m == 0xFFFFFFFF or m == -1 for pm in self.parameterized_meshes for m in pm.material_ids)
        if need:
            default_id = len(self.materials)
            self.materials.append(DisneyMaterial())
            for pm in self.parameterized_meshes:
                pm.material_ids = [default_id if (m == -1 or m == 0xFFFFFFFF) else m for m in pm.material_ids]

    def to_c(self) -> MarshalledScene:
        ms = MarshalledScene()
        keep = ms.keep
        cmeshes = (CMesh * max(1, len(self.meshes)))()
        for mi, mesh in enumerate(self.meshes):
            cgeoms = (CGeometry * max(1, len(mesh.geometries)))()
            for gi, g in enumerate(mesh.geometries):
                v = np.ascontiguousarray(g.vertices, dtype=np.float32).reshape(-1, 3)
                idx = np.ascontiguousarray(g.indices, dtype=np.uint32).reshape(-1, 3)
                keep += [v, idx]
                cgeoms[gi].vertices = _fptr(v)
                cgeoms[gi].indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
                cgeoms[gi].num_vertices = v.shape[0]
                cgeoms[gi].num_tris = idx.shape[0]
                if g.uvs is not None and len(g.uvs):
                    uv = np.ascontiguousarray(g.uvs, dtype=np.float32).reshape(-1, 2)
                    assert uv.shape[0] == v.shape[0]
                    keep.append(uv)
                    cgeoms[gi].uvs = _fptr(uv)
                else:
                    cgeoms[gi].uvs = None
            keep.append(cgeoms)
            cmeshes[mi].geometries = cgeoms
            cmeshes[mi].num_geometries = len(mesh.geometries)
        cpms = (CParameterizedMesh * max(1, len(self.parameterized_meshes)))()
        for i, pm in enumerate(self.parameterized_meshes):
            ids = np.array([m & 0xFFFFFFFF for m in pm.material_ids], dtype=np.uint32)
            keep.append(ids)
            cpms[i].material_ids = ids.ctypes.data_as(C.POINTER(C.c_uint32))
            cpms[i].num_material_ids = len(ids)
            cpms[i].mesh_id = pm.mesh_id
        cinst = (CInstance * max(1, len(self.instances)))()
        for i, inst in enumerate(self.instances):
            m = np.asarray(inst.transform, dtype=np.float32).reshape(4, 4)
            col_major = m.T.reshape(-1)  # glm::value_ptr layout
            for k in range(16):
                cinst[i].transform[k] = float(col_major[k])
            cinst[i].parameterized_mesh_id = inst.parameterized_mesh_id
        cmats = (CMaterial * max(1, len(self.materials)))()
        for i, mat in enumerate(self.materials):
            # texture handles are bit patterns (negative denormals / NaNs as floats): copy bitwise
            raw = np.array([f32_bits(x) for x in mat.as_floats()], dtype=np.uint32)
            C.memmove(C.addressof(cmats[i]), raw.ctypes.data, 64)
        ctex = (CImage * max(1, len(self.textures)))()
        for i, t in enumerate(self.textures):
            a = np.ascontiguousarray(t.img, dtype=np.uint8)
            keep.append(a)
            ctex[i].data = a.ctypes.data_as(C.POINTER(C.c_uint8))
            ctex[i].width = t.width
            ctex[i].height = t.height
            ctex[i].channels = t.channels
            ctex[i].color_space = t.color_space
        clights = (CQuadLight * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            raw = np.array(l.as_floats(), dtype=np.float32)
            C.memmove(C.addressof(clights[i]), raw.ctypes.data, 80)
        keep += [cmeshes, cpms, cinst, cmats, ctex, clights]
        c = ms.c
        c.meshes = cmeshes
        c.parameterized_meshes = cpms
        c.instances = cinst
        c.materials = cmats
        c.textures = ctex
        c.lights = clights
        c.num_meshes = len(self.meshes)
        c.num_parameterized_meshes = len(self.parameterized_meshes)
        c.num_instances = len(self.instances)
        c.num_materials = len(self.materials)
        c.num_textures = len(self.textures)
        c.num_lights = len(self.lights)
        c.samples_per_pixel = self.samples_per_pixel
        return ms
