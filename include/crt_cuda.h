/* crt_cuda.h — C ABI of libcrt_cuda_core.so, the B200-native wavefront path tracer that sits
 * behind ChameleonRT's RenderBackend interface.
 *
 * The reference's plugin boundary is C++ (a vtable and std::vector cross it:
 * util/render_backend.h:12-32, util/render_plugin.h:23-63). This header is the plain-C
 * surface underneath: one entry point per RenderBackend member, plain pointers and sizes, no
 * C++ or torch types. backends/cuda/render_cuda.cpp (the drop-in `crt_cuda` plugin a
 * ChameleonRT maintainer builds) and chameleonrt_b200/backend.py (ctypes) both bind exactly
 * these symbols; INTEGRATION.md shows the binding.
 *
 *   RenderBackend member (util/render_backend.h)        ->  C entry point
 *   ---------------------------------------------------------------------------------------
 *   constructor / make_renderer (render_plugin.h:41)    ->  crtc_create
 *   virtual ~RenderBackend()                 :16        ->  crtc_destroy
 *   std::string name()                       :18        ->  crtc_name
 *   void initialize(fb_width, fb_height)     :20        ->  crtc_initialize
 *   void set_scene(const Scene&)             :23        ->  crtc_set_scene
 *   RenderStats render(pos, dir, up, fovy,
 *        camera_changed, readback_framebuffer) :26-31   ->  crtc_render
 *   std::vector<uint32_t> img                :13        ->  the `img` out-parameter of crtc_render
 *   uint32_t samples_per_pixel               :14        ->  crt_scene_t::samples_per_pixel
 *
 * Error behaviour: the reference throws std::runtime_error and nothing catches it
 * (SURVEY.md §5). Here every call returns 0 on success and non-zero on failure, with the
 * message available from crtc_last_error(); the C++ plugin rethrows it as
 * std::runtime_error, the Python binding raises RuntimeError. There is NO CPU fallback: if
 * no CUDA device is usable crtc_create fails.
 *
 * Threading: like the reference, all calls on one renderer come from one thread.
 */
#ifndef CRT_CUDA_H
#define CRT_CUDA_H

#include <stdint.h>

#include "crt_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crtc_renderer crtc_renderer;

/* Message of the last failed call on this thread ("" if none). */
const char *crtc_last_error(void);

/* RenderBackend::name(). */
const char *crtc_name(void);

/* Creates a renderer on CUDA device `device` (cudaSetDevice ordinal). */
int crtc_create(crtc_renderer **out, int device);
void crtc_destroy(crtc_renderer *r);

/* Options the plugin API has no channel for (the reference hard-codes them; the plugin reads
 * CRT_CUDA_* environment variables and forwards them here). Must be set before
 * crtc_initialize / crtc_set_scene.
 *   "max_depth"   path depth, reference MAX_PATH_DEPTH = 5 (backends/embree/util.ih:10)
 *   "rank", "world_size"  image-tile sharding: this renderer owns the 64x64 tiles with
 *                 tile_id % world_size == rank (tile_id as in render_embree.cpp:178-180)
 *   "bvh_threads" host threads for the BVH8 build (0 = all)
 *   "tri_pass_defer" 0, 16 (default) or 24: scheduling variant of the traversal kernel — a warp's pooled
 *                 triangle pass waits until that many (ray, triangle) pairs are pending or no lane can descend.
 *                 Never changes a result. Measured on B200 (round 2): 16 shortens the frame by 2.8 % / 1.8 % / 6.4 % on
 *                 the C2 / C3 / C4 workloads, 24 by slightly less.
 *   "shade_sort"  0 (default), 1 or 2: the queue of paths to shade is bucketed by the material id of each path's hit
 *                 (one stable counting-sort pass on the device, 256 buckets) before the shading kernel runs, so that a
 *                 warp unpacks one material and runs the same BSDF lobes: 1 = from the first bounce on (primary hits
 *                 keep their screen order), 2 = every bounce. Never changes a result. Measured on B200 (round 2): the
 *                 frame gets 10 % LONGER (the sort passes and the scattered path-state accesses cost more than the
 *                 coherence buys); kept as an option, off.
 *   "hw_textures" 0 (default): textures are filtered in software with the reference's own arithmetic (texture2d.ih:
 *                 float->int truncation of texel coordinates, wrap by modulo, bilinear weights in float, over the 8-bit
 *                 LINEARISED texels render_embree.cpp:96-103 produces) — bit-level parity with the Embree backend.
 *                 1: one cudaTextureObject_t per texture (wrap, linear, normalised coordinates, sRGB decode in the texture
 *                 unit), as the reference's OptiX backend does (backends/optix/optix_utils.cpp:60-85): 8-bit fixed-point
 *                 filter weights and no re-quantisation of the linearised texels, hence a looser parity with the Embree
 *                 path (dark sRGB texels are not crushed to 0). Takes effect at the next crtc_set_scene.
 *   "bvh_top_smem" 0 (default) / 1: the traversal kernel reads the first 73 BVH nodes (root + two levels) from a shared-memory
 *                 copy — BASELINE.json's north_star names this staging; measured on B200 it costs 4-5 % (those nodes are the
 *                 L1's hottest lines anyway and the copy takes 47 KB of L1 per SM), so it is off. Never changes a result.
 *   "stage_events" 1 (default): a CUDA event after every launch feeds crtc_get_stage_times; 0: events only at the start
 *                 and the end of a frame ([6], the frame time, stays; the other stage times read 0) — 28 event records cost
 *                 ~0.07 ms per frame, which matters when a GPU renders a 1/8 shard in 2 ms. May be changed between frames.
 *   "pin_host_buffers" 1 (default): the caller's img buffer is page-locked on first readback (cudaHostRegister);
 *   "pin_read_img" 1: crtc_read_img page-locks its destination too (a frame loop reading into one buffer). Default 0.
 *   "bvh_builder" where crtc_set_scene builds the BVH8: 0 = on the host (binned SAH, the default); on the device
 *                 (chameleonrt_b200/csrc/bvh8_device.cuh): 1 = PLOC (mutual nearest neighbours in Morton order),
 *                 2 = LBVH (Karras), both followed by the host builder's 8-wide collapse — a much shorter set_scene
 *                 on scenes of millions of triangles for a somewhat slower tree. Rendered frames are bit-identical
 *                 in every case (closest hits break ties on the primitive id).
 *                 Developer knobs of the PLOC builder: "bvh_ploc_radius" (1..32, default 16: clusters searched on
 *                 either side) and "bvh_ploc_tail" (default 1: the last rounds run in a single block).
 *   "refill_idle" scheduling knob of the persistent traversal kernels (how many idle lanes
 *                 trigger a refill from the ray queue); the default is tuned
 *   "any_far_first" 1 = shadow (any-hit) rays visit the children of a BVH node farthest-first instead of
 *                 nearest-first. The result of an occlusion query does not depend on the order; the work does
 *                 (fewer node steps when occluders sit near the light's end of the segment, more when they sit
 *                 near the surface). 2 = decide per scene (the default): the second crtc_render frame after
 *                 crtc_set_scene runs far-first, the third near-first, and far-first is kept only if its traversal
 *                 stage was at least 3 % faster. 0 = off.
 *   "count_traversal" 1 = instrumented traversal kernels that count node visits and triangle
 *                 tests (for the algorithmic-byte figure; slower, off by default)
 */
int crtc_set_option(crtc_renderer *r, const char *key, int64_t value);

/* Reads an option back. Readable keys: every key of crtc_set_option, plus "any_far_first_decision" = the shadow-ray order
 * frames are rendered with from now on (0 near-first, 1 far-first, -1 = mode 2 has not decided yet), "bvh_build_rounds" =
 * the number of PLOC rounds of the last device build, and "bvh_builder_fallbacks" = device builds that gave up on their
 * input and were redone by the host builder. */
int crtc_get_option(crtc_renderer *r, const char *key, int64_t *value);

/* Use an existing CUDA stream (cudaStream_t) for all work of this renderer; NULL = the
 * renderer's own stream. Lets a host that already owns a stream (e.g. PyTorch's current
 * stream) order and time the work with its own events. */
int crtc_set_stream(crtc_renderer *r, void *cuda_stream);

/* RenderBackend::initialize: (re)allocates the framebuffer state and resets accumulation. */
int crtc_initialize(crtc_renderer *r, int fb_width, int fb_height);

/* RenderBackend::set_scene: flattens instances, builds the BVH8 on the host, uploads
 * geometry / materials / textures / lights, resets accumulation. */
int crtc_set_scene(crtc_renderer *r, const crt_scene_t *scene);

/* RenderBackend::render. pos/dir/up: 3 floats each; fovy in degrees. If
 * readback_framebuffer != 0 and img != NULL, img receives fb_width*fb_height RGBA8 pixels
 * (row-major, y down — RenderBackend::img). stats may be NULL. With world_size > 1 only the
 * tiles this rank owns are rendered and written; the rest of img is left untouched. */
int crtc_render(crtc_renderer *r, const float *pos, const float *dir, const float *up, float fovy,
                int camera_changed, int readback_framebuffer, uint32_t *img, crt_render_stats_t *stats);

/* Throughput variant of render for frame loops that do not need each frame's result on the host:
 * enqueues `num_frames` (>= 1) consecutive frames on the renderer's stream and returns immediately,
 * so the host can queue more work (or a gather) while they run. The frames of one call are
 * rendered as a single wavefront — num_frames * spp samples per pixel in flight together, seeded
 * and folded into the running mean exactly as num_frames separate calls would be (bit-identical
 * result) — which keeps a GPU that owns only part of the image busy. crtc_sync waits for all such frames and returns
 * their totals: `total` (render_time = sum of per-frame device times, num_rays = sum),
 * `stage_ms_sum` (7 floats, as crtc_get_stage_times), `counters_sum` (8, as crtc_get_counters),
 * `num_frames`; any of them may be NULL. The accumulation semantics are those of crtc_render. */
int crtc_render_async(crtc_renderer *r, const float *pos, const float *dir, const float *up, float fovy,
                      int camera_changed, uint32_t num_frames);
int crtc_sync(crtc_renderer *r, crt_render_stats_t *total, float *stage_ms_sum, uint64_t *counters_sum,
              uint32_t *num_frames);

/* The accumulated float framebuffer (what parity compares): fb_width*fb_height*3 floats,
 * row-major RGB. The reference keeps it backend-private and tile-major
 * (render_embree.h:26, render_embree.ispc:345); this is the extra export SURVEY.md §8b asks
 * for. */
int crtc_read_accum(crtc_renderer *r, float *rgb_out);

/* Per-stage device times of the last frame, in ms (CUDA events on the renderer's stream):
 * [0] raygen [1] traversal of the primary rays [2] shade [3] traversal of bounce b's shadow rays
 * together with bounce b+1's continuation rays (one launch per bounce) [4] NEE resolve
 * [5] resolve+tonemap [6] whole frame. Returns the number of entries written (<= n). */
int crtc_get_stage_times(crtc_renderer *r, float *ms_out, int n);

/* Counters of the last frame: [0] closest-hit rays [1] occlusion rays [2] kernel launches
 * [3] BVH nodes visited and [4] triangles tested by closest-hit traversal [5] paths started
 * [6] nodes visited and [7] triangles tested by any-hit traversal ([3],[4],[6],[7] only with
 * count_traversal=1). */
int crtc_get_counters(crtc_renderer *r, uint64_t *out, int n);

/* Scene/BVH facts: [0] triangles [1] BVH8 nodes [2] BVH depth [3] build ms [4] node bytes
 * [5] triangle bytes; the last crtc_set_scene in phases, ms: [6] flattening (host wall clock: host work, or uploads +
 * k_flatten), and for the device builders (CUDA events) [7] Morton keys + sort [8] binary tree (PLOC / LBVH) [9] BVH8
 * emission [10] record packing ([10] is host wall clock for the host builder, [7]-[9] are 0 there); [11] PLOC rounds. */
int crtc_get_scene_info(crtc_renderer *r, double *out, int n);

/* Kernel-level access for parity tests and micro-benchmarks: trace a batch of rays against
 * the current scene with the same traversal kernels render() uses. rays: n*8 floats
 * {ox,oy,oz,tnear,dx,dy,dz,tfar} in HOST memory; hits: n*4 floats {t,u,v,bits(flat prim id or
 * 0xffffffff)}; occluded: n bytes. */
int crtc_trace_closest(crtc_renderer *r, const float *rays, uint64_t n, float *hits);
int crtc_trace_any(crtc_renderer *r, const float *rays, uint64_t n, uint8_t *occluded);
/* Same kernels on DEVICE-resident rays, repeated `iters` times, returning the mean kernel time
 * in ms (CUDA events). For the roofline measurement of the traversal kernel in isolation. */
int crtc_bench_trace(crtc_renderer *r, const float *rays_host, uint64_t n, int any_hit, int iters,
                     float *ms_out);

/* Multi-GPU: device pointers of this rank's tile-local buffers, for the frame-end gather over
 * NVLink (torch.distributed / NCCL moves them; this library never touches the network).
 * accum: num_local_tiles*4096*3 floats; img: num_local_tiles*4096 RGBA8. */
int crtc_local_buffers(crtc_renderer *r, void **accum_dev, void **img_dev, uint32_t *num_local_tiles);
/* On the assembling rank: scatter one rank's gathered tile-local buffers (DEVICE pointers) into
 * the full row-major frame of this renderer. */
int crtc_assemble_rank(crtc_renderer *r, int src_rank, int world_size, const void *accum_dev,
                       const void *img_dev);
/* Multi-GPU without a gather (one process per GPU): the assembling rank exports its full-frame buffers
 * (handles_out: 128 bytes = two cudaIpcMemHandle_t, accum then img) and keeps writing its own tiles there; every
 * other rank imports them (ship the 128 bytes with any transport, e.g. torch.distributed.broadcast) and from then
 * on its frame-end resolve stores each of its pixels straight into the assembling rank's frame over NVLink, next to
 * its tile-local copy. Tile ownership is disjoint: nothing is reduced, no copy kernel runs, no collective library is
 * involved: the frame ends with a few flag words that travel through the same mapping — every rank's resolve kernel
 * publishes "frame s stored" there (system-scope fence + release store by its last block), the assembling rank's
 * stream waits for all of them before anything reads the frame (crtc_frame_wait; crtc_read_img / crtc_read_accum do it
 * themselves), and a rank does not store frame s + 1 before the assembling rank has started its own frame s + 1.
 * Every rank must therefore enqueue the same sequence of crtc_render / crtc_render_async calls; a flag that does not
 * arrive within 20 s becomes an error of the next synchronising call, not a hang.
 * crtc_initialize undoes both; crtc_import_frame(r, NULL) unmaps. */
int crtc_export_frame(crtc_renderer *r, void *handles_out);
int crtc_import_frame(crtc_renderer *r, const void *handles);
/* On the assembling rank of a shared frame (exported or crtc_share_frame'd): orders the renderer's stream after every
 * rank's stores of the last enqueued frame (a 1-warp kernel that spins on the flags). A no-op on other renderers. */
int crtc_frame_wait(crtc_renderer *r);
/* The same within ONE process (one host thread driving a renderer per GPU, as backends/cuda does for
 * CRT_CUDA_DEVICES): from now on `src` resolves its tiles into `dst`'s full frame; peer access between the two
 * devices is enabled if they differ. Both must be initialized with the same size; crtc_initialize undoes it. */
int crtc_share_frame(crtc_renderer *dst, crtc_renderer *src);
/* Presentation without a host round trip: copies the assembled RGBA8 frame, device to device, into
 * a cudaArray_t (passed as void*: this header needs no CUDA types) — the mapped array of an OpenGL texture registered with
 * cudaGraphicsGLRegisterImage, as the reference's OptiX backend presents (backends/optix/render_optix.cpp:410-426). The
 * array must live on this renderer's device and be fb_width x fb_height RGBA8. Returns when the copy has completed. */
int crtc_copy_img_to_array(crtc_renderer *r, void *cuda_array);
/* Read the assembled full frame (after crtc_assemble_rank for every rank, after every rank's peer-written
 * frame has completed, or after a world_size==1 render). */
int crtc_read_img(crtc_renderer *r, uint32_t *img);

#ifdef __cplusplus
}
#endif

#endif
