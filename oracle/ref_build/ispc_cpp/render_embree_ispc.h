// render_embree_ispc.h — OWN stand-in for the header the ISPC compiler would generate from
// backends/embree/render_embree.ispc (its two `export` functions, render_embree.ispc:199 and :356),
// included by the reference's render_embree.cpp:14. The definitions come from the same .ispc file
// compiled as scalar C++ (ispc_cpp/prologue.h).
#pragma once
#include <cstdint>

namespace ispc {
void trace_rays(void *scene, void *tile, const void *view_params);
void tile_to_uint8(void *tile, uint8_t *fb);
}  // namespace ispc
