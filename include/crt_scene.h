/* crt_scene.h — plain-C view of ChameleonRT's backend-neutral scene model.
 *
 * This is the data a backend receives in RenderBackend::set_scene(const Scene&)
 * (reference util/render_backend.h:23), restated as C structs of pointers+sizes so it
 * can cross a C ABI (ctypes / dlopen / cgo-style FFI). Field meaning, order and units
 * are exactly the reference's:
 *
 *   crt_geometry_t            <- Geometry            util/mesh.h:6-12   (normals are never
 *                                                    read by any kernel, SURVEY App.A #8,
 *                                                    so they are not carried)
 *   crt_mesh_t                <- Mesh                util/mesh.h:14-22
 *   crt_parameterized_mesh_t  <- ParameterizedMesh   util/mesh.h:28-36
 *   crt_instance_t            <- Instance            util/mesh.h:40-47  (glm::mat4, column major)
 *   crt_material_t            <- DisneyMaterial      util/material.h:29-46 (16 f32, 64 B)
 *   crt_image_t               <- Image               util/material.h:11-27
 *   crt_quad_light_t          <- QuadLight           util/lights.h:6-18   (20 f32, 80 B)
 *   crt_scene_t               <- Scene               util/scene.h:23-32
 *
 * All pointers are borrowed for the duration of the call that takes the scene; the
 * callee copies what it needs (the reference destroys its Scene right after set_scene,
 * main.cpp:185-214).
 */
#ifndef CRT_SCENE_H
#define CRT_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crt_geometry_t {
    const float *vertices;   /* 3 * num_vertices, object space (glm::vec3 AoS) */
    const float *uvs;        /* 2 * num_vertices or NULL (Geometry::uvs empty) */
    const uint32_t *indices; /* 3 * num_tris (glm::uvec3 AoS) */
    uint32_t num_vertices;
    uint32_t num_tris;
} crt_geometry_t;

typedef struct crt_mesh_t {
    const crt_geometry_t *geometries;
    uint32_t num_geometries;
} crt_mesh_t;

typedef struct crt_parameterized_mesh_t {
    const uint32_t *material_ids; /* one per geometry of the mesh */
    uint32_t num_material_ids;
    uint32_t mesh_id;
} crt_parameterized_mesh_t;

typedef struct crt_instance_t {
    float transform[16]; /* object_to_world, column major (glm::value_ptr layout) */
    uint32_t parameterized_mesh_id;
} crt_instance_t;

/* DisneyMaterial, util/material.h:29-46. A float with the sign bit set is a texture
 * handle (util/texture_channel_mask.h:16-23): bits[30:29]=channel, bits[28:0]=texture id. */
typedef struct crt_material_t {
    float base_color[3];
    float metallic;
    float specular;
    float roughness;
    float specular_tint;
    float anisotropy;
    float sheen;
    float sheen_tint;
    float clearcoat;
    float clearcoat_gloss;
    float ior;
    float specular_transmission;
    float pad[2];
} crt_material_t;

enum { CRT_COLOR_SPACE_LINEAR = 0, CRT_COLOR_SPACE_SRGB = 1 }; /* util/material.h:9 */

typedef struct crt_image_t {
    const uint8_t *data; /* width*height*channels, row-major, row 0 first */
    int32_t width;
    int32_t height;
    int32_t channels;
    int32_t color_space;
} crt_image_t;

typedef struct crt_quad_light_t {
    float emission[4];
    float position[4];
    float normal[4];
    float v_x[3];
    float width;
    float v_y[3];
    float height;
} crt_quad_light_t;

typedef struct crt_scene_t {
    const crt_mesh_t *meshes;
    const crt_parameterized_mesh_t *parameterized_meshes;
    const crt_instance_t *instances;
    const crt_material_t *materials;
    const crt_image_t *textures;
    const crt_quad_light_t *lights;
    uint32_t num_meshes;
    uint32_t num_parameterized_meshes;
    uint32_t num_instances;
    uint32_t num_materials;
    uint32_t num_textures;
    uint32_t num_lights;
    uint32_t samples_per_pixel; /* Scene::samples_per_pixel, util/scene.h:31 */
} crt_scene_t;

/* RenderStats, util/render_backend.h:7-10, plus what REPORT_RAY_STATS would sum
 * (backends/embree/render_embree.cpp:197-211). */
typedef struct crt_render_stats_t {
    float render_time;     /* ms */
    float rays_per_second; /* total rays / (render_time * 1e-3) */
    uint64_t num_rays;     /* closest-hit casts + occlusion casts this frame */
} crt_render_stats_t;

#ifdef __cplusplus
}
#endif

#endif
