// render_oracle_plugin.cpp — libcrt_oracle.so: the CPU oracle behind ChameleonRT's plugin API
// (TEST INFRASTRUCTURE). Lets the headless driver run `crt_headless oracle <scene>` beside
// `crt_headless cuda <scene>` on a scene loaded by the reference's own loaders.
#include <SDL.h>
#include <cstring>
#include <stdexcept>
#include <vector>
#include "imgui.h"
#include "render_backend.h"
#include "render_plugin.h"
#include "scene.h"
#include "../../include/crt_scene.h"

extern "C" {
void *oracle_create();
void oracle_destroy(void *);
void oracle_set_options(void *, int max_depth, int num_threads, int brute_force);
void oracle_initialize(void *, int, int);
void oracle_set_scene(void *, const crt_scene_t *);
void oracle_render(void *, const float *, const float *, const float *, float, int, crt_render_stats_t *);
void oracle_read_img(void *, uint32_t *);
void oracle_read_accum(void *, float *);
}

struct RenderOracle : RenderBackend {
    void *o;
    int w = 0, h = 0;
    crt_render_stats_t last{};  // the last frame's stats, for crt_oracle_get_stats
    RenderOracle() : o(oracle_create())
    {
        const char *d = std::getenv("CRT_CUDA_MAX_DEPTH");
        oracle_set_options(o, d ? std::atoi(d) : 5, 0, 0);
    }
    ~RenderOracle() override { oracle_destroy(o); }
    std::string name() override { return "CPU oracle (Embree/ISPC backend restated)"; }
    void initialize(const int fb_width, const int fb_height) override
    {
        w = fb_width;
        h = fb_height;
        img.resize(static_cast<size_t>(w) * h);
        oracle_initialize(o, w, h);
    }
    void set_scene(const Scene &scene) override
    {
        samples_per_pixel = scene.samples_per_pixel;
        std::vector<std::vector<crt_geometry_t>> geometries(scene.meshes.size());
        std::vector<crt_mesh_t> meshes(scene.meshes.size());
        for (size_t m = 0; m < scene.meshes.size(); ++m) {
            for (const auto &g : scene.meshes[m].geometries) {
                crt_geometry_t cg;
                cg.vertices = reinterpret_cast<const float *>(g.vertices.data());
                cg.uvs = g.uvs.empty() ? nullptr : reinterpret_cast<const float *>(g.uvs.data());
                cg.indices = reinterpret_cast<const uint32_t *>(g.indices.data());
                cg.num_vertices = static_cast<uint32_t>(g.vertices.size());
                cg.num_tris = static_cast<uint32_t>(g.indices.size());
                geometries[m].push_back(cg);
            }
            meshes[m].geometries = geometries[m].data();
            meshes[m].num_geometries = static_cast<uint32_t>(geometries[m].size());
        }
        std::vector<crt_parameterized_mesh_t> pms;
        for (const auto &pm : scene.parameterized_meshes) {
            pms.push_back(crt_parameterized_mesh_t{pm.material_ids.data(), static_cast<uint32_t>(pm.material_ids.size()),
                                                   static_cast<uint32_t>(pm.mesh_id)});
        }
        std::vector<crt_instance_t> instances;
        for (const auto &inst : scene.instances) {
            crt_instance_t c;
            std::memcpy(c.transform, &inst.transform[0][0], sizeof(c.transform));
            c.parameterized_mesh_id = static_cast<uint32_t>(inst.parameterized_mesh_id);
            instances.push_back(c);
        }
        std::vector<crt_image_t> textures;
        for (const auto &t : scene.textures) {
            textures.push_back(crt_image_t{t.img.data(), t.width, t.height, t.channels,
                                           t.color_space == SRGB ? CRT_COLOR_SPACE_SRGB : CRT_COLOR_SPACE_LINEAR});
        }
        crt_scene_t c;
        c.meshes = meshes.data();
        c.parameterized_meshes = pms.data();
        c.instances = instances.data();
        c.materials = reinterpret_cast<const crt_material_t *>(scene.materials.data());
        c.textures = textures.data();
        c.lights = reinterpret_cast<const crt_quad_light_t *>(scene.lights.data());
        c.num_meshes = static_cast<uint32_t>(meshes.size());
        c.num_parameterized_meshes = static_cast<uint32_t>(pms.size());
        c.num_instances = static_cast<uint32_t>(instances.size());
        c.num_materials = static_cast<uint32_t>(scene.materials.size());
        c.num_textures = static_cast<uint32_t>(textures.size());
        c.num_lights = static_cast<uint32_t>(scene.lights.size());
        c.samples_per_pixel = scene.samples_per_pixel;
        oracle_set_scene(o, &c);
    }
    RenderStats render(const glm::vec3 &pos, const glm::vec3 &dir, const glm::vec3 &up, const float fovy,
                       const bool camera_changed, const bool) override
    {
        crt_render_stats_t s;
        oracle_render(o, &pos.x, &dir.x, &up.x, fovy, camera_changed ? 1 : 0, &s);
        last = s;
        oracle_read_img(o, img.data());
        RenderStats stats;
        stats.render_time = s.render_time;
        stats.rays_per_second = s.rays_per_second;
        return stats;
    }
};

// the twin of crt_cuda_get_stats: a CPU frame has one stage (entry 6, whole frame) and one ray count (entry 0: closest-hit
// + occlusion rays together); everything else reads 0
extern "C" int crt_oracle_get_stats(RenderBackend *backend, float *stage_ms, int num_stages, uint64_t *counters, int num_counters)
{
    RenderOracle *r = dynamic_cast<RenderOracle *>(backend);
    if (!r) {
        return 1;
    }
    for (int i = 0; stage_ms && i < num_stages; ++i) {
        stage_ms[i] = i == 6 ? r->last.render_time : 0.f;
    }
    for (int i = 0; counters && i < num_counters; ++i) {
        counters[i] = i == 0 ? r->last.num_rays : 0;
    }
    return 0;
}

extern "C" int crt_oracle_read_accum(RenderBackend *backend, float *rgb_out)
{
    RenderOracle *r = dynamic_cast<RenderOracle *>(backend);
    if (!r) {
        return 1;
    }
    oracle_read_accum(r->o, rgb_out);
    return 0;
}

struct NullDisplay : Display {
    std::string gpu_brand() override { return "none"; }
    std::string name() override { return "null"; }
    void resize(const int, const int) override {}
    void new_frame() override {}
    void display(RenderBackend *) override {}
};

uint32_t get_sdl_window_flags() { return 0; }
void set_imgui_context(ImGuiContext *context) { ImGui::SetCurrentContext(context); }
std::unique_ptr<Display> make_display(SDL_Window *) { return std::make_unique<NullDisplay>(); }
std::unique_ptr<RenderBackend> make_renderer(Display *) { return std::make_unique<RenderOracle>(); }

POPULATE_PLUGIN_FUNCTIONS(get_sdl_window_flags, set_imgui_context, make_display, make_renderer)
