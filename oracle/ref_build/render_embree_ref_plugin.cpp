// render_embree_ref_plugin.cpp — libcrt_embree.so for the headless twin (TEST INFRASTRUCTURE).
// The renderer is the REFERENCE'S OWN Embree backend, compiled from the sources where they lie:
// backends/embree/render_embree.cpp + embree_utils.cpp (host side) and render_embree.ispc + *.ih
// (kernel side, as scalar C++: ispc_cpp/prologue.h), against third_party/{embree_stub,tbb_stub,miniglm}.
// This file only replaces backends/embree/render_embree_plugin.cpp:7-27 (which needs SDL + GLDisplay)
// and adds two exports next to the plugin API:
//   crt_embree_read_accum   the float framebuffer (RenderEmbree::tiles, render_embree.h:26) as w*h*3
//   refembree_*             a C API over RenderEmbree for ctypes (tests/, tests/golden/make_golden.py):
//                           takes the same crt_scene_t the CUDA backend and the oracle take
#include <SDL.h>
#include <xmmintrin.h>
#include <cstring>
#include <numeric>
#include <stdexcept>
#include <vector>
#include "imgui.h"
#include "render_embree.h"
#include "render_plugin.h"
#include "scene.h"
#include "../../include/crt_scene.h"

namespace {

void read_accum(const RenderEmbree &r, float *rgb_out)
{
    // inverse of the tile layout of render_embree.cpp:176-195
    const uint32_t w = r.fb_dims.x, h = r.fb_dims.y, ts = r.tile_size.x;
    const uint32_t ntx = w / ts + (w % ts != 0 ? 1 : 0);
    for (uint32_t y = 0; y < h; ++y) {
        for (uint32_t x = 0; x < w; ++x) {
            const uint32_t tx = x / ts, ty = y / ts;
            const uint32_t tw = std::min(ts, w - tx * ts);
            const std::vector<float> &tile = r.tiles[ty * ntx + tx];
            const uint32_t px = ((y - ty * ts) * tw + (x - tx * ts)) * 3;
            std::memcpy(rgb_out + (static_cast<size_t>(y) * w + x) * 3, tile.data() + px, 3 * sizeof(float));
        }
    }
}

struct NullDisplay : Display {
    std::string gpu_brand() override { return "none"; }
    std::string name() override { return "null"; }
    void resize(const int, const int) override {}
    void new_frame() override {}
    void display(RenderBackend *) override {}
};

// RenderEmbree's constructor switches the calling thread to flush-to-zero / denormals-are-zero
// (render_embree.cpp:21-24) and the threads its parallel loops start inherit that. Behind the ctypes API
// the caller is a Python interpreter thread, so the mode is confined to the reference's own calls.
struct FtzDazScope {
    unsigned saved;
    explicit FtzDazScope(bool enable = true) : saved(_mm_getcsr())
    {
        if (enable) {
            _mm_setcsr(saved | 0x8040);  // FTZ (bit 15) | DAZ (bit 6)
        }
    }
    ~FtzDazScope() { _mm_setcsr(saved); }
};

Scene scene_from_c(const crt_scene_t *c)
{
    Scene s;
    for (uint32_t m = 0; m < c->num_meshes; ++m) {
        std::vector<Geometry> geoms;
        for (uint32_t g = 0; g < c->meshes[m].num_geometries; ++g) {
            const crt_geometry_t &cg = c->meshes[m].geometries[g];
            Geometry geom;
            geom.vertices.resize(cg.num_vertices);
            std::memcpy(geom.vertices.data(), cg.vertices, sizeof(float) * 3 * cg.num_vertices);
            if (cg.uvs) {
                geom.uvs.resize(cg.num_vertices);
                std::memcpy(geom.uvs.data(), cg.uvs, sizeof(float) * 2 * cg.num_vertices);
            }
            geom.indices.resize(cg.num_tris);
            std::memcpy(geom.indices.data(), cg.indices, sizeof(uint32_t) * 3 * cg.num_tris);
            geoms.push_back(geom);
        }
        s.meshes.emplace_back(geoms);
    }
    for (uint32_t i = 0; i < c->num_parameterized_meshes; ++i) {
        const crt_parameterized_mesh_t &pm = c->parameterized_meshes[i];
        s.parameterized_meshes.emplace_back(pm.mesh_id,
                                            std::vector<uint32_t>(pm.material_ids, pm.material_ids + pm.num_material_ids));
    }
    for (uint32_t i = 0; i < c->num_instances; ++i) {
        glm::mat4 t;
        std::memcpy(&t[0][0], c->instances[i].transform, sizeof(float) * 16);
        s.instances.emplace_back(t, c->instances[i].parameterized_mesh_id);
    }
    static_assert(sizeof(DisneyMaterial) == sizeof(crt_material_t), "material layout");
    s.materials.resize(c->num_materials);
    if (c->num_materials) {
        std::memcpy(s.materials.data(), c->materials, sizeof(DisneyMaterial) * c->num_materials);
    }
    for (uint32_t i = 0; i < c->num_textures; ++i) {
        const crt_image_t &t = c->textures[i];
        s.textures.emplace_back(t.data, t.width, t.height, t.channels, "tex" + std::to_string(i),
                                t.color_space == CRT_COLOR_SPACE_SRGB ? SRGB : LINEAR);
    }
    static_assert(sizeof(QuadLight) == sizeof(crt_quad_light_t), "light layout");
    s.lights.resize(c->num_lights);
    if (c->num_lights) {
        std::memcpy(s.lights.data(), c->lights, sizeof(QuadLight) * c->num_lights);
    }
    s.samples_per_pixel = c->samples_per_pixel;
    return s;
}

}  // namespace

extern "C" int crt_embree_read_accum(RenderBackend *backend, float *rgb_out)
{
    RenderEmbree *r = dynamic_cast<RenderEmbree *>(backend);
    if (!r) {
        return 1;
    }
    read_accum(*r, rgb_out);
    return 0;
}

extern "C" {
void *refembree_create()
{
    FtzDazScope scope(false);  // the constructor sets the mode; restore the caller's afterwards
    return new RenderEmbree();
}
void refembree_destroy(void *r) { delete static_cast<RenderEmbree *>(r); }
const char *refembree_name(void *r)
{
    static std::string n;
    n = static_cast<RenderEmbree *>(r)->name();
    return n.c_str();
}
void refembree_initialize(void *r, int w, int h) { static_cast<RenderEmbree *>(r)->initialize(w, h); }
void refembree_set_scene(void *r, const crt_scene_t *c)
{
    FtzDazScope scope;
    static_cast<RenderEmbree *>(r)->set_scene(scene_from_c(c));
}
// returns the number of rays of the frame (REPORT_RAY_STATS: render_embree.cpp:197-203)
uint64_t refembree_render(void *r, const float *pos, const float *dir, const float *up, float fovy, int camera_changed,
                          float *render_time_ms)
{
    FtzDazScope scope;
    RenderEmbree *e = static_cast<RenderEmbree *>(r);
    const RenderStats st = e->render(glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]),
                                     glm::vec3(up[0], up[1], up[2]), fovy, camera_changed != 0, true);
    if (render_time_ms) {
        *render_time_ms = st.render_time;
    }
    uint64_t total = 0;
#ifdef REPORT_RAY_STATS
    total = std::accumulate(e->num_rays.begin(), e->num_rays.end(), uint64_t(0));
#endif
    return total;
}
void refembree_read_accum(void *r, float *rgb_out) { read_accum(*static_cast<RenderEmbree *>(r), rgb_out); }
void refembree_read_img(void *r, uint32_t *out)
{
    RenderEmbree *e = static_cast<RenderEmbree *>(r);
    std::memcpy(out, e->img.data(), e->img.size() * sizeof(uint32_t));
}
// per-pixel ray counts of the last frame (Tile::ray_stats), as w*h uint16
void refembree_read_ray_stats(void *r, uint16_t *out)
{
    RenderEmbree *e = static_cast<RenderEmbree *>(r);
    const uint32_t w = e->fb_dims.x, h = e->fb_dims.y, ts = e->tile_size.x;
    const uint32_t ntx = w / ts + (w % ts != 0 ? 1 : 0);
    for (uint32_t y = 0; y < h; ++y) {
        for (uint32_t x = 0; x < w; ++x) {
            const uint32_t tx = x / ts, ty = y / ts;
            const uint32_t tw = std::min(ts, w - tx * ts);
            out[static_cast<size_t>(y) * w + x] = e->ray_stats[ty * ntx + tx][(y - ty * ts) * tw + (x - tx * ts)];
        }
    }
}
}

uint32_t get_sdl_window_flags() { return 0; }
void set_imgui_context(ImGuiContext *context) { ImGui::SetCurrentContext(context); }
std::unique_ptr<Display> make_display(SDL_Window *) { return std::make_unique<NullDisplay>(); }
std::unique_ptr<RenderBackend> make_renderer(Display *) { return std::make_unique<RenderEmbree>(); }

POPULATE_PLUGIN_FUNCTIONS(get_sdl_window_flags, set_imgui_context, make_display, make_renderer)
