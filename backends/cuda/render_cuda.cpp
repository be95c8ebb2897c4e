// render_cuda.cpp — RenderBackend -> C ABI adapter (include/crt_cuda.h).
//
// set_scene borrows the Scene's own vectors: glm::vec3/uvec3/vec2 arrays are tightly packed
// float/uint32 triples, DisneyMaterial (util/material.h:29-46) and QuadLight
// (util/lights.h:6-18) have exactly the crt_material_t / crt_quad_light_t layout, so nothing is
// copied on this side; the core library copies what it needs (the reference destroys its Scene
// right after set_scene, main.cpp:185-214).
#include "render_cuda.h"
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>
#include "crt_cuda.h"
#include "scene.h"

static_assert(sizeof(DisneyMaterial) == sizeof(crt_material_t), "DisneyMaterial layout");
static_assert(sizeof(QuadLight) == sizeof(crt_quad_light_t), "QuadLight layout");
static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::uvec3) == 12 && sizeof(glm::vec2) == 8 &&
                  sizeof(glm::mat4) == 64,
              "GLM types must be tightly packed");

namespace {
void check(int rc)
{
    if (rc != 0) {
        // the reference's error behaviour: std::runtime_error propagating to the app
        throw std::runtime_error(std::string("crt_cuda: ") + crtc_last_error());
    }
}

long env_or(const char *name, long fallback)
{
    const char *v = std::getenv(name);
    return v && *v ? std::strtol(v, nullptr, 10) : fallback;
}
}

RenderCUDA::RenderCUDA()
{
    // The plugin API has no option channel: knobs come from the environment.
    check(crtc_create(&renderer, static_cast<int>(env_or("CRT_CUDA_DEVICE", 0))));
    check(crtc_set_option(renderer, "max_depth", env_or("CRT_CUDA_MAX_DEPTH", 5)));
    check(crtc_set_option(renderer, "bvh_threads", env_or("CRT_CUDA_BVH_THREADS", 0)));
    check(crtc_set_option(renderer, "bvh_builder", env_or("CRT_CUDA_BVH_BUILDER", 0)));  // 1 / 2 = build on the device (crt_cuda.h)
    check(crtc_set_option(renderer, "any_far_first", env_or("CRT_CUDA_ANY_FAR_FIRST", 2)));  // crt_cuda.h; 2 = per scene
}

RenderCUDA::~RenderCUDA()
{
    crtc_destroy(renderer);
}

std::string RenderCUDA::name()
{
    return crtc_name();
}

void RenderCUDA::initialize(const int fb_width, const int fb_height)
{
    fb_dims = glm::ivec2(fb_width, fb_height);
    img.resize(static_cast<size_t>(fb_width) * fb_height);
    check(crtc_initialize(renderer, fb_width, fb_height));
}

void RenderCUDA::set_scene(const Scene &scene)
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
        crt_parameterized_mesh_t c;
        c.material_ids = pm.material_ids.data();
        c.num_material_ids = static_cast<uint32_t>(pm.material_ids.size());
        c.mesh_id = static_cast<uint32_t>(pm.mesh_id);
        pms.push_back(c);
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
        crt_image_t c;
        c.data = t.img.data();
        c.width = t.width;
        c.height = t.height;
        c.channels = t.channels;
        c.color_space = t.color_space == SRGB ? CRT_COLOR_SPACE_SRGB : CRT_COLOR_SPACE_LINEAR;
        textures.push_back(c);
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
    check(crtc_set_scene(renderer, &c));
}

RenderStats RenderCUDA::render(const glm::vec3 &pos,
                               const glm::vec3 &dir,
                               const glm::vec3 &up,
                               const float fovy,
                               const bool camera_changed,
                               const bool readback_framebuffer)
{
    crt_render_stats_t s;
    // with a non-native display the app reads `img` every frame (gldisplay.cpp:111-123)
    const bool readback = readback_framebuffer || !native_display;
    check(crtc_render(renderer,
                      &pos.x,
                      &dir.x,
                      &up.x,
                      fovy,
                      camera_changed ? 1 : 0,
                      readback ? 1 : 0,
                      img.data(),
                      &s));
    RenderStats stats;
    stats.render_time = s.render_time;
    stats.rays_per_second = s.rays_per_second;
    return stats;
}

void RenderCUDA::read_accum(float *rgb_out)
{
    check(crtc_read_accum(renderer, rgb_out));
}

extern "C" int crt_cuda_read_accum(RenderBackend *backend, float *rgb_out)
{
    RenderCUDA *r = dynamic_cast<RenderCUDA *>(backend);
    if (!r) {
        return 1;
    }
    r->read_accum(rgb_out);
    return 0;
}
