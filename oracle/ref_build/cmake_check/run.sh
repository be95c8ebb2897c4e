#!/bin/bash
# Configures and builds backends/cuda/CMakeLists.txt inside the stand-in superproject (see CMakeLists.txt here). No GPU needed.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
BUILD="${1:-$HERE/../../_ref/cmake_check}"
cmake -S "$HERE" -B "$BUILD" -DCMAKE_BUILD_TYPE=Release -DCMAKE_CUDA_COMPILER="${NVCC:-/usr/local/cuda/bin/nvcc}" > "$BUILD.configure.log" 2>&1 || { tail -30 "$BUILD.configure.log"; exit 1; }
cmake --build "$BUILD" -j 8 > "$BUILD.build.log" 2>&1 || { tail -40 "$BUILD.build.log"; exit 1; }
ls -la "$BUILD"/libcrt_cuda.so "$BUILD"/backends/cuda/libcrt_scene_native.a
