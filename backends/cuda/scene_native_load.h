// scene_native_load.h — the native scene loader (include/crt_scene_io.h) behind the reference's own Scene type.
//
// main.cpp:186 builds `Scene scene(scene_file, material_mode);` (util/scene.cpp:49-67: tinyobjloader / tinygltf / the .crts
// reader, all single-threaded). A maintainer who wants the parallel / in-place loaders in ./chameleonrt replaces that line by
//
//     Scene scene = crt_cuda::load_scene_native(scene_file, material_mode);
//
// and links the static library crt_scene_native of this directory's CMakeLists.txt (scene_native_load.cpp +
// chameleonrt_b200/csrc/scene_io.cpp, zlib) into the application. The Scene that comes back
// is, array for array, the one the constructor builds (oracle/ref_build builds both and tests/test_scene_io.py compares them).
#pragma once

#include <string>
#include "scene.h"

namespace crt_cuda {

// threads: 0 = all hardware threads. Throws std::runtime_error like the Scene constructor.
Scene load_scene_native(const std::string &fname, MaterialMode material_mode, int threads = 0);

}
