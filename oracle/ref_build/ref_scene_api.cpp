// ref_scene_api.cpp — TEST INFRASTRUCTURE: the reference's own Scene loader (util/scene.cpp, util/tiny_obj_loader.h, stb_image;
// compiled from /root/reference by this directory's Makefile) behind a C API, so that tests/test_scene_io.py can compare what
// Scene::load_obj / Scene::load_crts build with what the product's native loader (chameleonrt_b200/csrc/scene_io.cpp) builds, array by array.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <exception>
#include <string>
#include "scene.h"
#ifdef CRT_REFSCENE_NATIVE  // the same C API over backends/cuda/scene_native_load.cpp: the shim a maintainer would call from main.cpp
#include "scene_native_load.h"
#endif

static_assert(sizeof(DisneyMaterial) == 64 && sizeof(QuadLight) == 80, "layouts");

namespace {
std::string g_err;
}

extern "C" {

void *refscene_load(const char *path, double *seconds)
{
    try {
        const auto t0 = std::chrono::steady_clock::now();
        Scene *s = new Scene(path, MaterialMode::DEFAULT);
        if (seconds) {
            *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        return s;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void *refscene_load_mode(const char *path, int white_diffuse, double *seconds)
{
    try {
        const auto t0 = std::chrono::steady_clock::now();
#ifdef CRT_REFSCENE_NATIVE
        Scene *s = new Scene(crt_cuda::load_scene_native(path, white_diffuse ? MaterialMode::WHITE_DIFFUSE : MaterialMode::DEFAULT));
#else
        Scene *s = new Scene(path, white_diffuse ? MaterialMode::WHITE_DIFFUSE : MaterialMode::DEFAULT);
#endif
        if (seconds) {
            *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        return s;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
const char *refscene_error()
{
    return g_err.c_str();
}
void refscene_free(void *p)
{
    delete static_cast<Scene *>(p);
}
// [0] meshes [1] geometries of mesh 0 [2] parameterized meshes [3] instances [4] materials [5] textures [6] lights
void refscene_counts(void *p, uint32_t *out)
{
    const Scene &s = *static_cast<Scene *>(p);
    out[0] = (uint32_t)s.meshes.size();
    out[1] = s.meshes.empty() ? 0u : (uint32_t)s.meshes[0].geometries.size();
    out[2] = (uint32_t)s.parameterized_meshes.size();
    out[3] = (uint32_t)s.instances.size();
    out[4] = (uint32_t)s.materials.size();
    out[5] = (uint32_t)s.textures.size();
    out[6] = (uint32_t)s.lights.size();
}
void refscene_geometry(void *p, uint32_t g, const float **verts, uint32_t *nv, const float **uvs, uint32_t *nuv, const uint32_t **idx,
                       uint32_t *ntris)
{
    const Geometry &geom = static_cast<Scene *>(p)->meshes[0].geometries[g];
    *verts = reinterpret_cast<const float *>(geom.vertices.data());
    *nv = (uint32_t)geom.vertices.size();
    *uvs = reinterpret_cast<const float *>(geom.uvs.data());
    *nuv = (uint32_t)geom.uvs.size();
    *idx = reinterpret_cast<const uint32_t *>(geom.indices.data());
    *ntris = (uint32_t)geom.indices.size();
}
const uint32_t *refscene_material_ids(void *p, uint32_t *n, uint32_t *mesh_id)
{
    const ParameterizedMesh &pm = static_cast<Scene *>(p)->parameterized_meshes[0];
    *n = (uint32_t)pm.material_ids.size();
    *mesh_id = (uint32_t)pm.mesh_id;
    return pm.material_ids.data();
}
const float *refscene_materials(void *p)
{
    return reinterpret_cast<const float *>(static_cast<Scene *>(p)->materials.data());
}
const float *refscene_light(void *p, uint32_t i)
{
    return reinterpret_cast<const float *>(&static_cast<Scene *>(p)->lights[i]);
}
const float *refscene_instance(void *p, uint32_t i, uint32_t *pm)
{
    const Instance &inst = static_cast<Scene *>(p)->instances[i];
    *pm = (uint32_t)inst.parameterized_mesh_id;
    return reinterpret_cast<const float *>(&inst.transform);
}
// ---- scenes of several meshes / parameterized meshes (.crts) ----
void refscene_mesh_geometry(void *p, uint32_t mesh, uint32_t g, const float **verts, uint32_t *nv, const float **uvs, uint32_t *nuv,
                            const uint32_t **idx, uint32_t *ntris, uint32_t *num_geometries)
{
    const Mesh &m = static_cast<Scene *>(p)->meshes[mesh];
    *num_geometries = (uint32_t)m.geometries.size();
    const Geometry &geom = m.geometries[g];
    *verts = reinterpret_cast<const float *>(geom.vertices.data());
    *nv = (uint32_t)geom.vertices.size();
    *uvs = reinterpret_cast<const float *>(geom.uvs.data());
    *nuv = (uint32_t)geom.uvs.size();
    *idx = reinterpret_cast<const uint32_t *>(geom.indices.data());
    *ntris = (uint32_t)geom.indices.size();
}
const uint32_t *refscene_parameterized_mesh(void *p, uint32_t i, uint32_t *n, uint32_t *mesh_id)
{
    const ParameterizedMesh &pm = static_cast<Scene *>(p)->parameterized_meshes[i];
    *n = (uint32_t)pm.material_ids.size();
    *mesh_id = (uint32_t)pm.mesh_id;
    return pm.material_ids.data();
}
// Camera: position, center, up, fov_y = 10 floats
const float *refscene_cameras(void *p, uint32_t *n)
{
    static_assert(sizeof(Camera) == 40, "layout");
    const Scene &s = *static_cast<Scene *>(p);
    *n = (uint32_t)s.cameras.size();
    return reinterpret_cast<const float *>(s.cameras.data());
}
const uint8_t *refscene_texture(void *p, uint32_t i, int *w, int *h, int *channels, int *color_space)
{
    const Image &im = static_cast<Scene *>(p)->textures[i];
    *w = im.width;
    *h = im.height;
    *channels = im.channels;
    *color_space = (int)im.color_space;
    return im.img.data();
}
const char *refscene_texture_name(void *p, uint32_t i)
{
    return static_cast<Scene *>(p)->textures[i].name.c_str();
}
}
