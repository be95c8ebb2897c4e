/* crt_scene_io.h — C ABI of the native scene loader (scene-load throughput and formats).
 *
 * The step in front of RenderBackend::set_scene is the reference's Scene constructor (util/scene.cpp:49-67), which picks
 * Scene::load_obj, load_gltf or load_crts by the file extension. All three run on one thread and copy everything they read.
 * The loaders here build THE SAME Scene — every array, handle, light and camera, bit for bit (tests/test_scene_io.py compares
 * with the reference's own loader compiled from its sources, on hand-written, generated and randomly fuzzed files):
 *
 *   OBJ   (Scene::load_obj, :94-228: tinyobjloader, then one hash map per shape that turns (position, normal, texcoord) index
 *         triples into single indices) from a memory-mapped file parsed in parallel: lines are classified and counted per
 *         chunk, prefix sums place every chunk's vertices / faces, the chunks are parsed into the global arrays concurrently
 *         (numbers with tinyobjloader's own decimal-to-double rule, so that the bits match), shapes are delimited from the
 *         g / o / usemtl events, faces of more than three corners are cut by tinyobjloader's ear clipping restated in float
 *         (tiny_obj_loader.h:1107-1310: same triangles, same order, also where it gives up on a degenerate polygon; faces of
 *         fewer than three corners are skipped), and the shapes are remapped concurrently, one open-addressing table each.
 *         MTL: newmtl / Kd / Ns / map_Kd with its texture options (other statements are ignored as the reference ignores them).
 *   .crts (Scene::load_crts, :417-625: the reference's binary format) with the geometry arrays left where they are in the
 *         mapped file, lights, cameras, every DisneyMaterial parameter with its texture handles.
 *   glTF  (Scene::load_gltf, :230-415, through tinygltf and util/flatten_gltf.cpp): .gltf + buffers in files or data: URIs,
 *         or .glb; packed accessors in place, the node hierarchy flattened in glm's float arithmetic.
 *
 * Textures — the reference decodes them with stb_image to four components (flipped vertically for OBJ and .crts,
 * util/material.cpp:5-17; not for glTF) — are decoded concurrently, to the bytes stb_image returns: PNG of every colour type
 * and bit depth, Adam7 interlacing and tRNS transparency included; TGA true-colour / grey / colour-mapped, raw or run-length
 * encoded, with stb_image's reading of 15/16-bit pixels; BMP with palettes, 16 / 24 / 32 bits and bit-field masks; JPEG (chameleonrt_b200/csrc/jpeg_decode.h) baseline and progressive,
 * grey or three components, any sampling factors, restart intervals, with stb_image's integer inverse DCT, chroma upsampling
 * and YCbCr -> RGB conversion.
 *
 * The result is a crt_scene_t (include/crt_scene.h) owned by the handle: pass crtio_scene_view(h) to crtc_set_scene.
 * backends/cuda/scene_native_load.cpp fills the reference's own Scene struct from it (the swap for main.cpp:186). */
#ifndef CRT_SCENE_IO_H
#define CRT_SCENE_IO_H

#include "crt_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crtio_scene crtio_scene;

/* Camera, util/camera.h:5-8 (what Scene::cameras holds; the application picks one with -camera, main.cpp). */
typedef struct crtio_camera_t {
    float position[3];
    float center[3];
    float up[3];
    float fov_y;
} crtio_camera_t;

/* Scene::load_obj (util/scene.cpp:94-228). threads: 0 = all hardware threads. Returns 0 and *out on success; otherwise a
 * non-zero code and crtio_last_error() (the reference throws std::runtime_error). */
int crtio_load_obj(const char *path, int threads, crtio_scene **out);
/* Scene::load_crts (util/scene.cpp:417-625): the reference's own binary format (a uint64 header size, a JSON header, a data
 * block of buffer views) — one geometry per mesh, MESH objects that instance a (mesh, material) pair under a matrix, LIGHT
 * objects (quad lights, frame = the object's matrix), CAMERA objects, every DisneyMaterial parameter with optional texture
 * handles, images as embedded PNG / JPEG / TGA / BMP files. The geometry arrays are NOT copied: the file stays mapped for the
 * lifetime of the handle and crt_geometry_t points into it (arrays that are not 4-byte aligned in the file are copied). */
int crtio_load_crts(const char *path, int threads, crtio_scene **out);
/* Scene::load_gltf (util/scene.cpp:230-415, which reads the file through tinygltf and flattens the scene graph with
 * util/flatten_gltf.cpp): glTF 2.0 as .gltf (buffers in external files or data: URIs) or .glb. A glTF mesh becomes a mesh plus
 * its parameterized mesh (primitives = geometries, POSITION / TEXCOORD_0 as floats, 16- or 32-bit indices), the nodes of the
 * default scene that carry a mesh become instances (node transforms composed down the hierarchy in glm's float arithmetic),
 * pbrMetallicRoughness becomes base colour / metallic / roughness with texture handles (base colour textures sRGB,
 * metallic-roughness textures linear, B and G channels), the light is generated. Accessors that are tightly packed and
 * aligned are used in place in the mapped buffer; interleaved or 16-bit ones are gathered.
 * Not read (an error): sparse accessors, non-indexed or non-triangle primitives, extensions that move data (Draco, meshopt). */
int crtio_load_gltf(const char *path, int threads, crtio_scene **out);
/* Scene::Scene (util/scene.cpp:49-67): the loader is chosen by the file extension (obj, gltf, glb, crts). */
int crtio_load(const char *path, int threads, crtio_scene **out);
/* Scene::Scene with a MaterialMode (util/scene.h:21, main.cpp's -mat-mode): CRTIO_MATERIALS_WHITE_DIFFUSE reads no materials
 * (and, for glTF, no images) and gives every geometry the default DisneyMaterial. */
enum { CRTIO_MATERIALS_DEFAULT = 0, CRTIO_MATERIALS_WHITE_DIFFUSE = 1 };
int crtio_load_mode(const char *path, int threads, int material_mode, crtio_scene **out);
/* The cameras of the file (CAMERA objects of a .crts; none for an OBJ): returns how many, *out = the array. */
int crtio_cameras(const crtio_scene *s, const crtio_camera_t **out);
/* The loaded scene as the plain-C view crtc_set_scene takes; valid until crtio_free. samples_per_pixel is 1 (the
 * application sets it from its command line, main.cpp:186). */
const crt_scene_t *crtio_scene_view(const crtio_scene *s);
/* Image::name of texture i (OBJ: the map_Kd string; .crts / glTF: the image's "name"); "" past the end. */
const char *crtio_texture_name(const crtio_scene *s, uint32_t i);
/* Wall-clock seconds of the phases of the load: [0] total [1] parse (OBJ: mmap + both passes + polygons; .crts: the header)
 * [2] index remap (OBJ only) [3] materials + textures. Returns the number of entries written. */
int crtio_timings(const crtio_scene *s, double *out, int n);
/* Warnings the reference would print (per-face material ids, missing material file, ...), one per line. */
const char *crtio_warnings(const crtio_scene *s);
void crtio_free(crtio_scene *s);
const char *crtio_last_error(void);

#ifdef __cplusplus
}
#endif

#endif
