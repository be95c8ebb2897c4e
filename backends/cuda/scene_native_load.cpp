// scene_native_load.cpp — see scene_native_load.h. Copies the loader's crt_scene_t view into the reference's Scene
// (util/scene.h:23-32); the element layouts are the same (glm::vec3 / glm::uvec3 / glm::vec2 arrays, DisneyMaterial 64 B,
// QuadLight 80 B, glm::mat4 column major), so every array is one block copy.
#include "scene_native_load.h"

#include <cstring>
#include <memory>
#include <stdexcept>
#include "crt_scene_io.h"

static_assert(sizeof(DisneyMaterial) == sizeof(crt_material_t), "DisneyMaterial layout");
static_assert(sizeof(QuadLight) == sizeof(crt_quad_light_t), "QuadLight layout");
static_assert(sizeof(Camera) == sizeof(crtio_camera_t), "Camera layout");
static_assert(sizeof(glm::vec3) == 12 && sizeof(glm::uvec3) == 12 && sizeof(glm::vec2) == 8 && sizeof(glm::mat4) == 64, "glm layouts");

namespace crt_cuda {

Scene load_scene_native(const std::string &fname, MaterialMode material_mode, int threads)
{
    crtio_scene *handle = nullptr;
    const int mode = material_mode == MaterialMode::WHITE_DIFFUSE ? CRTIO_MATERIALS_WHITE_DIFFUSE : CRTIO_MATERIALS_DEFAULT;
    if (crtio_load_mode(fname.c_str(), threads, mode, &handle) != 0) {
        throw std::runtime_error(crtio_last_error());
    }
    const std::unique_ptr<crtio_scene, void (*)(crtio_scene *)> owner(handle, crtio_free);
    const crt_scene_t &v = *crtio_scene_view(handle);
    Scene scene;
    scene.material_mode = material_mode;
    scene.meshes.resize(v.num_meshes);
    for (uint32_t m = 0; m < v.num_meshes; ++m) {
        scene.meshes[m].geometries.resize(v.meshes[m].num_geometries);
        for (uint32_t g = 0; g < v.meshes[m].num_geometries; ++g) {
            const crt_geometry_t &src = v.meshes[m].geometries[g];
            Geometry &dst = scene.meshes[m].geometries[g];
            dst.vertices.resize(src.num_vertices);
            dst.indices.resize(src.num_tris);
            if (src.num_vertices) {
                std::memcpy(dst.vertices.data(), src.vertices, sizeof(glm::vec3) * src.num_vertices);
            }
            if (src.num_tris) {
                std::memcpy(dst.indices.data(), src.indices, sizeof(glm::uvec3) * src.num_tris);
            }
            if (src.uvs && src.num_vertices) {
                dst.uvs.resize(src.num_vertices);
                std::memcpy(dst.uvs.data(), src.uvs, sizeof(glm::vec2) * src.num_vertices);
            }
        }
    }
    for (uint32_t i = 0; i < v.num_parameterized_meshes; ++i) {
        const crt_parameterized_mesh_t &pm = v.parameterized_meshes[i];
        scene.parameterized_meshes.emplace_back(pm.mesh_id, std::vector<uint32_t>(pm.material_ids, pm.material_ids + pm.num_material_ids));
    }
    scene.instances.resize(v.num_instances);
    for (uint32_t i = 0; i < v.num_instances; ++i) {
        std::memcpy(&scene.instances[i].transform, v.instances[i].transform, sizeof(glm::mat4));
        scene.instances[i].parameterized_mesh_id = v.instances[i].parameterized_mesh_id;
    }
    scene.materials.resize(v.num_materials);
    if (v.num_materials) {
        std::memcpy(static_cast<void *>(scene.materials.data()), v.materials, sizeof(DisneyMaterial) * v.num_materials);
    }
    for (uint32_t i = 0; i < v.num_textures; ++i) {
        const crt_image_t &t = v.textures[i];
        scene.textures.emplace_back(t.data, t.width, t.height, t.channels, crtio_texture_name(handle, i),
                                    t.color_space == CRT_COLOR_SPACE_SRGB ? SRGB : LINEAR);
    }
    scene.lights.resize(v.num_lights);
    if (v.num_lights) {
        std::memcpy(static_cast<void *>(scene.lights.data()), v.lights, sizeof(QuadLight) * v.num_lights);
    }
    const crtio_camera_t *cameras = nullptr;
    const int num_cameras = crtio_cameras(handle, &cameras);
    scene.cameras.resize((size_t)num_cameras);
    if (num_cameras) {
        std::memcpy(static_cast<void *>(scene.cameras.data()), cameras, sizeof(Camera) * (size_t)num_cameras);
    }
    return scene;
}

}
