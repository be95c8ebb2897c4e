import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """Native code is built once per session (nvcc cross-compiles on the CPU box)."""
    import __graft_entry__ as g

    csrc = os.path.join(ROOT, "chameleonrt_b200", "csrc")
    need = [os.path.join(csrc, "libcrt_cuda_core.so"), os.path.join(csrc, "libcrt_scene_io.so"), os.path.join(csrc, "libcrt_bvh8_hostcheck.so"),
            os.path.join(csrc, "libcrt_shade_hostcheck.so"), os.path.join(csrc, "libcrt_wavefront_hostcheck.so"),
            os.path.join(csrc, "libcrt_wavefront_hostcheck_powf.so"), os.path.join(csrc, "libcrt_simt_hostcheck.so"),
            os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    return True


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "oracle_kat.npz"))
