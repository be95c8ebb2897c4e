"""chameleonrt_b200 — a B200-native wavefront path tracer behind ChameleonRT's RenderBackend API.

Only the per-frame render path is here (BASELINE.json north_star / SURVEY.md §8):
``csrc/`` holds the sm_100a CUDA kernels, the host BVH8 builder and the C ABI
(``include/crt_cuda.h``); ``backend.RenderCUDA`` mirrors ``RenderBackend``; ``scene`` mirrors the
reference's ``Scene`` model; ``scenes`` generates the benchmark stand-ins.
"""
from .backend import RenderCUDA, load_lib  # noqa: F401
from .camera import ArcballCamera  # noqa: F401
from .scene import (DisneyMaterial, Geometry, Image, Instance, Mesh, ParameterizedMesh, QuadLight,  # noqa: F401
                    RenderStats, Scene)
