"""Builds tests/simt_emu/_build/libcrt_cuda_core_simt.so: the PRODUCT's renderer object and C ABI
(chameleonrt_b200/csrc/crt_cuda_core.cu, kernels.cuh, ...) compiled for the host, with every kernel launch executed
under the SIMT environment of chameleonrt_b200/csrc/simt_env.h and the CUDA runtime replaced by cuda_emu.cpp.
TEST INFRASTRUCTURE: the only edit made to the product source is mechanical — each
    kernel<<<grid, block, 0, stream>>>(args);
becomes
    simt::launch(grid, block, [&] { kernel(args); });
(g++ cannot parse the chevrons). Nothing in chameleonrt_b200/ loads this library; tests/test_simt_renderer.py
points a RenderCUDA at it explicitly."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "chameleonrt_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libcrt_cuda_core_simt.so")


def translate(src: str) -> str:
    out, pos = [], 0
    pat = re.compile(r"(crt::k_\w+(?:<\w+>)?)<<<([^,]+),\s*([^,]+),\s*0,\s*stream>>>\(")
    n = 0
    while True:
        m = pat.search(src, pos)
        if not m:
            break
        out.append(src[pos:m.start()])
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        assert src[i] == ";", src[m.start():i + 1]
        out.append(f"simt::launch({m.group(2).strip()}, {m.group(3).strip()}, [&] {{ {m.group(1)}({src[m.end():i - 1]}); }});")
        pos = i + 1
        n += 1
    out.append(src[pos:])
    assert n == 7, f"expected 7 kernel launches in crt_cuda_core.cu, found {n}"
    return '#include "simt_env.h"\n' + "".join(out)


def build(force: bool = False) -> str:
    srcs = [os.path.join(CSRC, f) for f in ("crt_cuda_core.cu", "kernels.cuh", "shade_math.cuh", "bvh8_traverse.h", "simt_env.h",
                                            "host_scene.cpp", "bvh8_build.cpp")] + [os.path.join(HERE, "cuda_emu.cpp"), __file__]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    gen = os.path.join(OUT, "crt_cuda_core_simt.cpp")
    with open(gen, "w") as f:
        f.write(translate(open(os.path.join(CSRC, "crt_cuda_core.cu")).read().replace('"../../include/crt_cuda.h"',
                                                                                          f'"{ROOT}/include/crt_cuda.h"')))
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"))), "include")
    subprocess.check_call(["make", "-s", "-C", CSRC, "host_scene.o", "bvh8_build.o"])
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-pthread", "-march=x86-64-v3", "-ffp-contract=off", "-Wno-attributes",
                           "-I" + CSRC, "-I" + cuda_inc, "-shared", "-o", LIB, gen, os.path.join(HERE, "cuda_emu.cpp"),
                           os.path.join(CSRC, "host_scene.o"), os.path.join(CSRC, "bvh8_build.o")])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
