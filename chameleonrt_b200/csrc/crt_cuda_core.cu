// crt_cuda_core.cu — the renderer object behind the C ABI of include/crt_cuda.h.
//
// Host-side role = RenderEmbree / RenderOptiX (reference backends/embree/render_embree.cpp,
// backends/optix/render_optix.cpp): own the device scene, compute the camera basis, launch the
// frame, time it, read the framebuffer back. Everything per-frame runs on the GPU as the
// kernel chain of kernels.cuh; there is no CPU fallback anywhere in this file.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/crt_cuda.h"
#include "bvh8.h"
#include "cuda_host_utils.h"
#include "host_scene.h"
#include "kernels.cuh"
#include "scene_device_build.cuh"

namespace {

thread_local std::string g_last_error;

using crt_host::ArenaBuf;
using crt_host::DeviceArena;
using crt_host::DeviceBuffer;

enum Stage { kStRaygen = 0, kStPrimary, kStShade, kStTraverse, kStNee, kStResolve, kStFrame, kNumStages };

// What one enqueued frame leaves behind for later collection: its stage events and a pinned copy
// of its device counters. Several frames may be in flight (crtc_render_async), so every frame has
// its own record.
struct FrameRecord {
    std::vector<cudaEvent_t> events;
    std::vector<int> event_stage;  // stage that ENDS at event i (i >= 1)
    size_t num_events = 0;
    uint32_t *h_counters = nullptr;  // pinned
    unsigned long long *h_trav = nullptr;
    uint32_t launches = 0;
    uint32_t frames = 1;  // frames rendered by this record's launch sequence

    FrameRecord()
    {
        CUDA_CHECK(cudaMallocHost(&h_counters, crt::kNumCounters * sizeof(uint32_t)));
        CUDA_CHECK(cudaMallocHost(&h_trav, 4 * sizeof(unsigned long long)));
    }
    ~FrameRecord()
    {
        for (auto e : events) {
            cudaEventDestroy(e);
        }
        cudaFreeHost(h_counters);
        cudaFreeHost(h_trav);
    }
    bool all_stages = true;  // false: only the first and the last mark of a frame are recorded (option "stage_events" = 0)
    void mark(cudaStream_t s, int stage_ended)
    {
        if (!all_stages && stage_ended != -1 && stage_ended != kStResolve) {
            return;
        }
        if (num_events >= events.size()) {
            cudaEvent_t e;
            CUDA_CHECK(cudaEventCreate(&e));
            events.push_back(e);
            event_stage.push_back(0);
        }
        event_stage[num_events] = stage_ended;
        CUDA_CHECK(cudaEventRecord(events[num_events++], s));
    }
};

}  // namespace

struct crtc_renderer {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;

    // options
    int max_depth = 5;
    int rank = 0, world_size = 1;
    int bvh_threads = 0;
    int bvh_builder = 0;  // 0 = host (binned SAH, bvh8_build.cpp); on the device (bvh8_device.cuh): 1 = PLOC, 2 = LBVH
    int build_rounds = 0;
    int ploc_radius = crt::kPlocRadius;
    bool ploc_tail = true;  // the last rounds of PLOC in one block (k_ploc_tail); off only to test that the tree is the same
    bool count_traversal = false;
    int refill_idle = crt::kRefillIdle;  // idle lanes that trigger a refill of the traversal warps
    bool bvh_top_smem = false;           // option "bvh_top_smem": the first 73 BVH nodes read from a shared-memory copy (measured: no gain)
    int tri_pass_defer = 16;             // 0 / 16 (default) / 24: pooled pairs a triangle pass waits for (kernels.cuh)
    // The shade queue bucketed by material id before k_shade (k_queue_hist / k_queue_scatter): 0 = off (default),
    // 1 = from the first bounce on (primary hits keep their screen order), 2 = every shade launch. Same image.
    int shade_sort = 0;
    // Shadow rays visit the children of a node farthest-first: 0 = no, 1 = yes, 2 = auto (default) — frame 1
    // after set_scene is rendered far-first, frame 2 near-first, and far-first is kept from frame 3 on only if its
    // traversal stage was at least 3 % faster (blocking render() calls only; the image is the same either way).
    // Far-first is tried on the EARLIER, possibly colder frame on purpose: whatever warm-up effect is left works
    // against the switch, so a wrong decision can only mean staying with near-first.
    int any_far_first = 2;
    int auto_frames = 0;            // blocking frames rendered since set_scene (auto mode)
    float auto_trav_ms[2] = {0.f, 0.f};
    bool auto_decided = false, auto_choice = false;
    bool frame_far_first = false;   // the order used by the frame being enqueued

    // `blocking`: the frame comes from render(), whose stage times feed auto_tune_step(). Frames enqueued with
    // crtc_render_async never advance the trial, so while it is undecided they run near-first (the order the
    // instrumented counting pass of bench.py assumes for "undecided").
    bool pick_far_first(bool blocking) const
    {
        if (any_far_first != 2) {
            return any_far_first == 1;
        }
        return auto_decided ? auto_choice : (blocking && auto_frames == 1);
    }
    // after a blocking frame's stage times are known
    void auto_tune_step()
    {
        if (any_far_first != 2 || auto_decided) {
            return;
        }
        if (auto_frames == 1 || auto_frames == 2) {
            auto_trav_ms[auto_frames - 1] = stage_ms[kStTraverse];  // [0] far-first (frame 1), [1] near-first (frame 2)
        }
        if (auto_frames == 2) {
            auto_decided = true;
            auto_choice = auto_trav_ms[0] < 0.97f * auto_trav_ms[1];  // switch only for a clear gain
        }
        ++auto_frames;
    }

    // framebuffer layout
    int fb_w = 0, fb_h = 0;
    uint32_t ntx = 0, nty = 0;
    std::vector<uint32_t> local_tiles;
    uint32_t npx_local = 0;
    uint32_t frame_id = 0;
    uint32_t spp = 1;
    bool have_scene = false;

    // device scene
    DeviceBuffer<float4> d_nodes, d_tris, d_shade, d_materials, d_lights;
    DeviceBuffer<uint32_t> d_texels;
    DeviceBuffer<crt::DevTex> d_tex;
    // option "hw_textures": one CUDA array + texture object per scene texture (wrap, linear, normalised coordinates, sRGB
    // decode in the texture unit — the reference's OptiX backend, backends/optix/optix_utils.cpp:60-85)
    bool hw_textures = false;
    std::vector<cudaArray_t> tex_arrays;
    std::vector<cudaTextureObject_t> tex_objects;
    DeviceBuffer<unsigned long long> d_tex_objects;
    void release_hw_textures()
    {
        for (cudaTextureObject_t t : tex_objects) {
            cudaDestroyTextureObject(t);
        }
        for (cudaArray_t a : tex_arrays) {
            cudaFreeArray(a);
        }
        tex_objects.clear();
        tex_arrays.clear();
    }
    void build_hw_textures(const crt_scene_t *scene)
    {
        release_hw_textures();
        if (!hw_textures || scene->num_textures == 0) {
            d_tex_objects.release();
            return;
        }
        std::vector<uchar4> rgba;
        for (uint32_t i = 0; i < scene->num_textures; ++i) {
            const crt_image_t &im = scene->textures[i];  // (shape checked by convert_shading_inputs / flatten_scene)
            const size_t n = (size_t)im.width * im.height;
            rgba.assign(n, make_uchar4(0, 0, 0, 0));  // channels the image lacks read as 0 (texture2d.ih:13-27)
            for (size_t px = 0; px < n; ++px) {
                unsigned char c[4] = {0, 0, 0, 0};
                for (int k = 0; k < im.channels; ++k) {
                    c[k] = im.data[px * im.channels + k];
                }
                rgba[px] = make_uchar4(c[0], c[1], c[2], c[3]);
            }
            const cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
            cudaArray_t arr = nullptr;
            CUDA_CHECK(cudaMallocArray(&arr, &fmt, (size_t)im.width, (size_t)im.height));
            tex_arrays.push_back(arr);
            CUDA_CHECK(cudaMemcpy2DToArrayAsync(arr, 0, 0, rgba.data(), (size_t)im.width * 4, (size_t)im.width * 4, (size_t)im.height,
                                                cudaMemcpyHostToDevice, stream));
            CUDA_CHECK(cudaStreamSynchronize(stream));  // `rgba` is reused
            cudaResourceDesc res = {};
            res.resType = cudaResourceTypeArray;
            res.res.array.array = arr;
            cudaTextureDesc td = {};
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeWrap;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeNormalizedFloat;
            td.sRGB = im.color_space == CRT_COLOR_SPACE_SRGB ? 1 : 0;
            td.normalizedCoords = 1;
            cudaTextureObject_t obj = 0;
            CUDA_CHECK(cudaCreateTextureObject(&obj, &res, &td, nullptr));
            tex_objects.push_back(obj);
        }
        static_assert(sizeof(cudaTextureObject_t) == sizeof(unsigned long long), "texture handles are 64-bit");
        d_tex_objects.upload(reinterpret_cast<const unsigned long long *>(tex_objects.data()), tex_objects.size(), stream);
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    uint32_t num_lights = 0;
    std::vector<uint32_t> leaf_flat_ids;  // host copy: leaf-order triangle -> flattened prim id
    double scene_info[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // crtc_get_scene_info
    double phase_ms[5] = {0, 0, 0, 0, 0};  // last set_scene: flatten, keys + sort, binary tree, BVH8 emission, packing

    // path state
    size_t path_capacity = 0;
    DeviceBuffer<float4> d_ray_o, d_ray_d, d_hit, d_thr_rng, d_radiance, d_nee_T, d_nee_l1, d_nee_l2, d_sray_o,
        d_sray_d;
    DeviceBuffer<uint8_t> d_vis;
    DeviceBuffer<uint32_t> d_queue0, d_queue1, d_counters;
    DeviceBuffer<uint32_t> d_queue_sorted, d_sort_hist, d_sort_scratch;  // option shade_sort
    DeviceBuffer<unsigned long long> d_trav_counters;

    // framebuffers
    DeviceBuffer<uint32_t> d_tile_ids;
    DeviceBuffer<float> d_accum_local, d_accum_full;
    DeviceBuffer<uint32_t> d_img_local, d_img_full;

    // frames: records of enqueued-but-not-collected frames, and a pool of reusable ones
    std::vector<std::unique_ptr<FrameRecord>> in_flight, record_pool;
    float stage_ms[kNumStages] = {0};   // last collected frame
    uint64_t counters_out[8] = {0};     // last collected frame

    // Peer-written frames (multi-GPU without a gather): the assembling rank exports its full-frame buffers as CUDA
    // IPC handles, the other ranks map them and their k_resolve stores each pixel there as well (st.global to a
    // peer address over NVLink). Tile ownership is disjoint, so there is nothing to reduce.
    bool frame_exported = false;
    // Completion flags of a shared frame (kernels.cuh "frame completion flags"): share_seq counts the wavefronts enqueued
    // since the frame was shared (every rank enqueues the same sequence of crtc_render / crtc_render_async calls).
    uint32_t share_seq = 0;
    DeviceBuffer<uint32_t> d_done_counter;
    uint32_t *h_sync_err = nullptr, *d_sync_err = nullptr;  // one mapped host word: a flag wait that timed out
    uint32_t *sync_words() { return d_img_full.ptr + (size_t)fb_w * fb_h; }
    uint32_t *peer_sync_words() { return peer_img_full + (size_t)fb_w * fb_h; }
    void ensure_sync_state()
    {
        if (!d_done_counter.ptr) {
            d_done_counter.alloc(1);
            CUDA_CHECK(cudaMemsetAsync(d_done_counter.ptr, 0, sizeof(uint32_t), stream));
        }
        if (!h_sync_err) {
            CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&h_sync_err), sizeof(uint32_t), cudaHostAllocMapped));
            *h_sync_err = 0u;
            CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&d_sync_err), h_sync_err, 0));
        }
    }
    void reset_share_sequence()
    {
        share_seq = 0;
        if (d_img_full.ptr && fb_w) {
            CUDA_CHECK(cudaMemsetAsync(sync_words(), 0, crt::kSyncWords * sizeof(uint32_t), stream));
            CUDA_CHECK(cudaStreamSynchronize(stream));
        }
    }
    void check_sync_error()
    {
        if (h_sync_err && *h_sync_err) {
            const uint32_t who = *h_sync_err - 1u;
            *h_sync_err = 0u;
            throw std::runtime_error("shared frame: rank " + std::to_string(who) + "'s completion flag did not arrive within 20 s "
                                     "(every rank must enqueue the same sequence of frames)");
        }
    }
    // On the assembling rank, stream-ordered: everything enqueued afterwards sees the frame of the last enqueued
    // wavefront complete, every rank's tiles included.
    void frame_wait()
    {
        if (!frame_exported || world_size < 2 || share_seq == 0) {
            return;
        }
        make_current();
        ensure_sync_state();
        crt::k_wait_words<<<1, 32, 0, stream>>>(sync_words(), (uint32_t)world_size, (uint32_t)rank, share_seq, d_sync_err);
        CUDA_CHECK(cudaGetLastError());
    }
    float *peer_accum_full = nullptr;
    uint32_t *peer_img_full = nullptr;
    bool peer_is_ipc = false;  // mapped through cudaIpcOpenMemHandle (another process) or a plain peer pointer (this one)

    void close_peer_frame()
    {
        if (peer_accum_full && peer_is_ipc) {
            cudaIpcCloseMemHandle(peer_accum_full);
            cudaIpcCloseMemHandle(peer_img_full);
        }
        peer_accum_full = nullptr;
        peer_img_full = nullptr;
    }
    // The in-process form (one host thread driving several renderers, as the ChameleonRT plugin does): this renderer
    // resolves its tiles into `dst`'s full frame; different devices need peer access, which NVSwitch gives every pair.
    void share_frame_of(crtc_renderer *dst)
    {
        if (dst == this) {
            throw std::runtime_error("crtc_share_frame: source and destination are the same renderer");
        }
        if (fb_w == 0 || dst->fb_w != fb_w || dst->fb_h != fb_h) {
            throw std::runtime_error("crtc_share_frame: both renderers must be initialized with the same size");
        }
        make_current();
        close_peer_frame();
        if (dst->device != device) {
            int can = 0;
            CUDA_CHECK(cudaDeviceCanAccessPeer(&can, device, dst->device));
            if (!can) {
                throw std::runtime_error("crtc_share_frame: device " + std::to_string(device) + " cannot access device " +
                                         std::to_string(dst->device));
            }
            const cudaError_t err = cudaDeviceEnablePeerAccess(dst->device, 0);
            if (err == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();  // not an error: clear it
            } else {
                CUDA_CHECK(err);
            }
        }
        peer_accum_full = dst->d_accum_full.ptr;
        peer_img_full = dst->d_img_full.ptr;
        peer_is_ipc = false;
        share_seq = 0;
        if (!dst->frame_exported) {
            dst->make_current();
            dst->reset_share_sequence();
            make_current();
        }
        dst->frame_exported = true;
    }
    void export_frame(void *handles_out)
    {
        if (fb_w == 0) {
            throw std::runtime_error("crtc_export_frame: initialize() has not been called");
        }
        make_current();
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two 64-byte handles");
        cudaIpcMemHandle_t h[2];
        CUDA_CHECK(cudaIpcGetMemHandle(&h[0], d_accum_full.ptr));
        CUDA_CHECK(cudaIpcGetMemHandle(&h[1], d_img_full.ptr));
        std::memcpy(handles_out, h, sizeof(h));
        reset_share_sequence();
        frame_exported = true;
    }
    void import_frame(const void *handles)
    {
        make_current();
        close_peer_frame();
        if (!handles) {
            return;
        }
        cudaIpcMemHandle_t h[2];
        std::memcpy(h, handles, sizeof(h));
        void *a = nullptr, *i = nullptr;
        CUDA_CHECK(cudaIpcOpenMemHandle(&a, h[0], cudaIpcMemLazyEnablePeerAccess));
        const cudaError_t err = cudaIpcOpenMemHandle(&i, h[1], cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            cudaIpcCloseMemHandle(a);
            CUDA_CHECK(err);
        }
        peer_accum_full = static_cast<float *>(a);
        peer_img_full = static_cast<uint32_t *>(i);
        peer_is_ipc = true;
        share_seq = 0;
    }

    // The caller's `img` (RenderBackend::img: a std::vector / numpy array that lives as long as the framebuffer size) is
    // page-locked on first use so that the frame-end readback is one DMA transfer instead of a staged pageable copy
    // (3.7 MB at 1280x720: ~0.35 ms -> ~0.15 ms per frame). Registration failures are not errors: the copy then stays pageable.
    void *pinned_img = nullptr;
    size_t pinned_bytes = 0;
    void unpin_img()
    {
        if (pinned_img) {
            cudaHostUnregister(pinned_img);
            cudaGetLastError();
            pinned_img = nullptr;
            pinned_bytes = 0;
        }
    }
    void pin_img(void *img, size_t bytes)
    {
        if (img == pinned_img && bytes == pinned_bytes) {
            return;
        }
        unpin_img();
        if (pin_host_buffers && cudaHostRegister(img, bytes, cudaHostRegisterDefault) == cudaSuccess) {
            pinned_img = img;
            pinned_bytes = bytes;
        } else {
            cudaGetLastError();
        }
    }
    // Option "stage_events": 1 (default) = a CUDA event after every launch, for the per-stage times of
    // crtc_get_stage_times; 0 = only at the start and the end of a frame (the stage times then read 0 except [6], the
    // whole frame; with a 1/8 shard per GPU the 27 event records are a measurable part of a 2 ms frame).
    bool stage_events = true;
    bool pin_host_buffers = true;  // option "pin_host_buffers"
    bool pin_repeated_reads = false;  // option "pin_read_img": crtc_read_img page-locks its destination too (a frame loop)

    ~crtc_renderer()
    {
        unpin_img();
        release_hw_textures();
        close_peer_frame();
        in_flight.clear();
        record_pool.clear();
        if (own_stream) {
            cudaStreamDestroy(own_stream);
        }
        if (h_sync_err) {
            cudaFreeHost(h_sync_err);
        }
    }

    void make_current() { CUDA_CHECK(cudaSetDevice(device)); }

    // persistent traversal grid: every SM filled with as many blocks as fit
    unsigned trav_grid = 0;
    int num_sms = 0;
    void launch_traverse(const crt::DeviceScene &sc, const crt::PathState &ps, const uint32_t *queue,
                         const uint32_t *count_closest, const uint32_t *count_any, uint32_t *work_counter)
    {
        if (trav_grid == 0) {
            int sms = 0, per_sm = 0;
            CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, crt::k_traverse<false>, crt::kTravBlock, 0));
            trav_grid = (unsigned)(sms * std::max(1, per_sm));
        }
        const int sched = (refill_idle & 0xff) | (frame_far_first ? 0x100 : 0);
        if (bvh_top_smem && !count_traversal && tri_pass_defer == 16) {  // experiment (kernels.cuh: TOP), default scheduling only
            crt::k_traverse<false, 16, true><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched,
                                                                                         (uint32_t)scene_info[1]);
            return;
        }
        // (one instantiation per variant, so that the default kernel's code does not change with the options)
        const int variant = (count_traversal ? 1 : 0) | (tri_pass_defer == 16 ? 2 : (tri_pass_defer == 24 ? 4 : 0));
        switch (variant) {
        case 0:
            crt::k_traverse<false><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched);
            break;
        case 1:
            crt::k_traverse<true><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched);
            break;
        case 2:
            crt::k_traverse<false, 16><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched);
            break;
        case 3:
            crt::k_traverse<true, 16><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched);
            break;
        case 4:
            crt::k_traverse<false, 24><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched);
            break;
        default:
            crt::k_traverse<true, 24><<<trav_grid, crt::kTravBlock, 0, stream>>>(sc, ps, queue, count_closest, count_any, work_counter, sched);
            break;
        }
    }

    // exclusive scan of n uint32 in place; `scratch` holds the tile sums of every recursion level. Returns its launches.
    int scan_u32(uint32_t *data, uint32_t n, uint32_t *scratch)
    {
        const unsigned tiles = (n + crt::kBuildTile - 1) / crt::kBuildTile;
        crt::k_scan_tile<uint32_t><<<tiles, crt::kBuildBlock, 0, stream>>>(data, data, n, tiles > 1 ? scratch : nullptr);
        if (tiles <= 1) {
            return 1;
        }
        const int below = scan_u32(scratch, tiles, scratch + tiles);
        crt::k_scan_add<uint32_t><<<tiles, crt::kBuildBlock, 0, stream>>>(data, n, scratch);
        return below + 2;
    }

    // Option shade_sort: queue `queue_in` of bounce `bounce` (length on the device, capacity `npaths`) bucketed by the
    // material of each path's hit into d_queue_sorted. Returns the number of launches.
    int sort_shade_queue(const crt::DeviceScene &sc, const crt::PathState &ps, const uint32_t *queue_in, int bounce, size_t npaths)
    {
        const uint32_t num_tiles = (uint32_t)((npaths + crt::kSortTile - 1) / crt::kSortTile);
        const uint32_t hist_items = 256u * num_tiles;
        size_t scratch_items = 1, left = hist_items;
        while (left > (size_t)crt::kBuildTile) {
            left = (left + crt::kBuildTile - 1) / crt::kBuildTile;
            scratch_items += left;
        }
        if (d_queue_sorted.count < npaths || d_sort_hist.count < hist_items || d_sort_scratch.count < scratch_items) {
            CUDA_CHECK(cudaStreamSynchronize(stream));  // (first frame with the option on, or a larger batch)
            d_queue_sorted.alloc(npaths);
            d_sort_hist.alloc(hist_items);
            d_sort_scratch.alloc(scratch_items);
        }
        crt::k_queue_hist<256><<<num_tiles, crt::kSortBlock, 0, stream>>>(sc, ps, queue_in, bounce, d_sort_hist.ptr, num_tiles);
        const int scans = scan_u32(d_sort_hist.ptr, hist_items, d_sort_scratch.ptr);
        crt::k_queue_scatter<256><<<num_tiles, crt::kSortBlock, 0, stream>>>(sc, ps, queue_in, d_queue_sorted.ptr, bounce, d_sort_hist.ptr, num_tiles);
        return scans + 2;
    }

    crt::DeviceScene device_scene() const
    {
        crt::DeviceScene sc;
        sc.nodes = d_nodes.ptr;
        sc.tris = d_tris.ptr;
        sc.shade = d_shade.ptr;
        sc.materials = d_materials.ptr;
        sc.lights = d_lights.ptr;
        sc.texels = d_texels.ptr;
        sc.tex = d_tex.ptr;
        sc.tex_objects = d_tex_objects.ptr;
        sc.num_lights = num_lights;
        sc.float_one = 0x3F800000u;
        return sc;
    }

    crt::FrameLayout frame_layout() const
    {
        crt::FrameLayout f;
        f.fb_w = fb_w;
        f.fb_h = fb_h;
        f.ntx = ntx;
        f.npx_local = npx_local;
        f.spp = spp;
        f.frames = 1;
        f.tile_ids = d_tile_ids.ptr;
        return f;
    }

    crt::PathState path_state()
    {
        crt::PathState ps;
        ps.ray_o = d_ray_o.ptr;
        ps.ray_d = d_ray_d.ptr;
        ps.hit = d_hit.ptr;
        ps.thr_rng = d_thr_rng.ptr;
        ps.radiance = d_radiance.ptr;
        ps.nee_T = d_nee_T.ptr;
        ps.nee_l1 = d_nee_l1.ptr;
        ps.nee_l2 = d_nee_l2.ptr;
        ps.sray_o = d_sray_o.ptr;
        ps.sray_d = d_sray_d.ptr;
        ps.vis = d_vis.ptr;
        ps.queue[0] = d_queue0.ptr;
        ps.queue[1] = d_queue1.ptr;
        ps.counters = d_counters.ptr;
        ps.trav_counters = d_trav_counters.ptr;
        return ps;
    }

    void ensure_path_buffers(size_t npaths)
    {
        if (npaths <= path_capacity && d_counters.ptr) {
            return;
        }
        d_ray_o.alloc(npaths);
        d_ray_d.alloc(npaths);
        d_hit.alloc(npaths);
        d_thr_rng.alloc(npaths);
        d_radiance.alloc(npaths);
        d_nee_T.alloc(npaths);
        d_nee_l1.alloc(npaths);
        d_nee_l2.alloc(npaths);
        d_sray_o.alloc(2 * npaths);
        d_sray_d.alloc(2 * npaths);
        d_vis.alloc(2 * npaths);
        d_queue0.alloc(npaths);
        d_queue1.alloc(npaths);
        d_counters.alloc(crt::kNumCounters);
        d_trav_counters.alloc(4);
        path_capacity = npaths;
    }

    void initialize(int w, int h)
    {
        if (w <= 0 || h <= 0) {
            throw std::runtime_error("initialize: framebuffer dimensions must be positive");
        }
        make_current();
        CUDA_CHECK(cudaStreamSynchronize(stream));
        unpin_img();             // (the caller's img buffer is reallocated with the new size)
        close_peer_frame();      // the frame's shape (and on the exporting rank, its buffers) change:
        frame_exported = false;  // crtc_export_frame / crtc_import_frame must be called again
        frame_id = 0;
        fb_w = w;
        fb_h = h;
        // render_embree.cpp:43-44
        ntx = w / crt::kTile + (w % crt::kTile != 0 ? 1 : 0);
        nty = h / crt::kTile + (h % crt::kTile != 0 ? 1 : 0);
        local_tiles.clear();
        for (uint32_t t = 0; t < ntx * nty; ++t) {
            if ((int)(t % (uint32_t)world_size) == rank) {
                local_tiles.push_back(t);
            }
        }
        npx_local = (uint32_t)local_tiles.size() * crt::kTilePixels;
        assemble_tiles.clear();
        d_tile_ids.upload(local_tiles.data(), local_tiles.size(), stream);
        d_accum_local.alloc((size_t)npx_local * 3);
        d_img_local.alloc(npx_local);
        d_accum_full.alloc((size_t)w * h * 3);
        d_img_full.alloc((size_t)w * h + crt::kSyncWords);  // + the frame completion flags (kernels.cuh)
        if (npx_local) {
            CUDA_CHECK(cudaMemsetAsync(d_accum_local.ptr, 0, (size_t)npx_local * 3 * sizeof(float), stream));
            CUDA_CHECK(cudaMemsetAsync(d_img_local.ptr, 0, (size_t)npx_local * 4, stream));
        }
        CUDA_CHECK(cudaMemsetAsync(d_accum_full.ptr, 0, (size_t)w * h * 3 * sizeof(float), stream));
        CUDA_CHECK(cudaMemsetAsync(d_img_full.ptr, 0, ((size_t)w * h + crt::kSyncWords) * 4, stream));
        share_seq = 0;
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    static void check_depth(uint32_t depth) { crt_host::check_bvh_depth(depth); }
    // k_traverse packs a pooled (owner lane, leaf-order triangle index) pair into one 32-bit slot: 5 + 27 bits
    static void check_triangle_count(size_t n)
    {
        if (n >= (size_t)crt::kMaxTriangles) {
            throw std::runtime_error("set_scene: " + std::to_string(n) + " triangles; this backend holds at most 2^27 - 1 (" +
                                     std::to_string(crt::kMaxTriangles - 1) + ") per scene");
        }
    }
    int builder_fallbacks = 0;  // device builds that fell back to the host builder (crtc_get_option "bvh_builder_fallbacks")

    void set_scene(const crt_scene_t *scene)
    {
        make_current();
        frame_id = 0;
        auto_frames = 0;
        auto_decided = false;
        crt::HostScene hs;
        uint32_t bvh_nodes = 0, bvh_depth = 0;
        double bvh_ms = 0.0;
        size_t num_tris = 0;
        // (the plan is cheap — reference checks and counts — and gives the triangle total before anything is allocated)
        crt::FlattenPlan plan;
        crt::plan_flatten(scene, plan);
        check_triangle_count(plan.total_tris);
        bool built_on_device = false;
        if (bvh_builder != 0 && plan.total_tris > 0) {
            check_triangle_count(plan.total_tris);
            // set_scene on the device: only the references are checked and the materials / textures converted on the host
            crt::convert_shading_inputs(scene, hs, bvh_threads);
            crt_host::DeviceSceneBuild job;
            job.stream = stream;
            job.device = device;
            job.builder = bvh_builder;
            job.ploc_radius = ploc_radius;
            job.ploc_tail = ploc_tail;
            job.d_nodes = &d_nodes;
            job.d_tris = &d_tris;
            job.d_shade = &d_shade;
            job.leaf_flat_ids = &leaf_flat_ids;
            try {
                job.run(scene, plan);
                built_on_device = true;
            } catch (const crt_host::DeviceBuildFailure &e) {
                // the builder gave up on this input (no PLOC progress, a tree deeper than the traversal stack, ...): the
                // host builder takes over; scene errors and CUDA errors are not caught here
                std::fprintf(stderr, "crt_cuda: %s; building the BVH on the host instead\n", e.what());
                CUDA_CHECK(cudaStreamSynchronize(stream));
                hs = crt::HostScene();
                builder_fallbacks++;
            }
            if (built_on_device) {
                num_tris = plan.total_tris;
                bvh_nodes = job.num_nodes;
                bvh_depth = job.depth;
                bvh_ms = job.build_ms;
                build_rounds = job.rounds;
                for (int i = 0; i < 5; ++i) {
                    phase_ms[i] = job.phase_ms[i];
                }
            }
        }
        if (!built_on_device) {
            const auto t0 = std::chrono::steady_clock::now();
            crt::flatten_scene(scene, hs, bvh_threads);
            phase_ms[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            phase_ms[1] = phase_ms[2] = phase_ms[3] = 0.0;  // the host builder reports one figure: scene_info[3]
            num_tris = hs.num_tris();
            check_triangle_count(num_tris);
            crt::Bvh8 bvh;
            crt::build_bvh8(hs.tri_verts.data(), hs.num_tris(), bvh_threads, bvh);
            check_depth(bvh.max_depth);
            std::vector<float> tri_records;
            std::vector<crt::TriShade> shade;
            const auto t_pack = std::chrono::steady_clock::now();
            crt::pack_triangles(hs, bvh, tri_records, shade, bvh_threads);
            phase_ms[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pack).count();
            leaf_flat_ids.resize(shade.size());
            for (size_t i = 0; i < shade.size(); ++i) {
                leaf_flat_ids[i] = shade[i].flat_id;
            }
            static_assert(sizeof(crt::Bvh8Node) == 5 * sizeof(float4), "node = 5 float4");
            d_nodes.upload(reinterpret_cast<const float4 *>(bvh.nodes.data()), bvh.nodes.size() * 5, stream);
            // keep at least one (degenerate) record so empty scenes have valid pointers
            if (tri_records.empty()) {
                tri_records.assign(12, 0.f);
                shade.resize(1);
                std::memset(&shade[0], 0, sizeof(crt::TriShade));
            }
            d_tris.upload(reinterpret_cast<const float4 *>(tri_records.data()), tri_records.size() / 4, stream);
            d_shade.upload(reinterpret_cast<const float4 *>(shade.data()), shade.size() * 3, stream);
            CUDA_CHECK(cudaStreamSynchronize(stream));  // the staging vectors go out of scope
            bvh_nodes = (uint32_t)bvh.nodes.size();
            bvh_depth = bvh.max_depth;
            bvh_ms = bvh.build_seconds * 1e3;
        }
        static_assert(sizeof(crt_material_t) == 64 && sizeof(crt_quad_light_t) == 80, "layouts");
        if (hs.materials.empty()) {
            hs.materials.resize(1);
            std::memset(&hs.materials[0], 0, sizeof(crt_material_t));
        }
        d_materials.upload(reinterpret_cast<const float4 *>(hs.materials.data()), hs.materials.size() * 4, stream);
        d_lights.upload(reinterpret_cast<const float4 *>(hs.lights.data()), hs.lights.size() * 5, stream);
        num_lights = (uint32_t)hs.lights.size();
        if (hs.texels.empty()) {
            hs.texels.assign(1, 0u);
        }
        d_texels.upload(hs.texels.data(), hs.texels.size(), stream);
        std::vector<crt::DevTex> tex(std::max<size_t>(1, hs.tex_desc.size()));
        for (size_t i = 0; i < hs.tex_desc.size(); ++i) {
            tex[i] = crt::DevTex{hs.tex_desc[i].offset, hs.tex_desc[i].width, hs.tex_desc[i].height, 0};
        }
        d_tex.upload(tex.data(), tex.size(), stream);
        CUDA_CHECK(cudaStreamSynchronize(stream));
        build_hw_textures(scene);
        spp = std::max<uint32_t>(1u, hs.samples_per_pixel);
        have_scene = true;
        scene_info[0] = (double)num_tris;
        scene_info[1] = (double)bvh_nodes;
        scene_info[2] = (double)bvh_depth;
        scene_info[3] = bvh_ms;
        scene_info[4] = (double)bvh_nodes * 80.0;
        scene_info[5] = (double)num_tris * 48.0;
        for (int i = 0; i < 5; ++i) {
            scene_info[6 + i] = phase_ms[i];
        }
        scene_info[11] = (bvh_builder == 1 && built_on_device) ? (double)build_rounds : 0.0;
        if (npx_local) {
            CUDA_CHECK(cudaMemsetAsync(d_accum_local.ptr, 0, (size_t)npx_local * 3 * sizeof(float), stream));
        }
    }

    static float3 glm_normalize(float3 v)
    {
        // glm::normalize(v) = v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x)
        const float inv = 1.f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        return make_float3(v.x * inv, v.y * inv, v.z * inv);
    }
    static float3 cross3(float3 a, float3 b)
    {
        return make_float3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
    }

    // render_embree.cpp:149-159 / render_optix.cpp:454-460
    crt::ViewParams view_params(const float *pos, const float *dir_, const float *up_, float fovy) const
    {
        const float3 dir = make_float3(dir_[0], dir_[1], dir_[2]);
        const float3 up = make_float3(up_[0], up_[1], up_[2]);
        const float plane_y = 2.f * std::tan((0.5f * fovy) * 0.01745329251994329576923690768489f);
        const float plane_x = plane_y * static_cast<float>(fb_w) / static_cast<float>(fb_h);
        crt::ViewParams v;
        v.pos = make_float3(pos[0], pos[1], pos[2]);
        const float3 du = glm_normalize(cross3(dir, up));
        v.dir_du = make_float3(du.x * plane_x, du.y * plane_x, du.z * plane_x);
        const float3 dvn = glm_normalize(cross3(v.dir_du, dir));
        v.dir_dv = make_float3(-dvn.x * plane_y, -dvn.y * plane_y, -dvn.z * plane_y);
        v.dir_top_left = make_float3(dir.x - 0.5f * v.dir_du.x - 0.5f * v.dir_dv.x,
                                     dir.y - 0.5f * v.dir_du.y - 0.5f * v.dir_dv.y,
                                     dir.z - 0.5f * v.dir_du.z - 0.5f * v.dir_dv.z);
        v.frame_id = frame_id;
        return v;
    }

    // Enqueues one frame on the stream (no host synchronisation) and returns its record.
    // `num_frames` consecutive frames are rendered as ONE wavefront (their samples are in flight
    // together) and folded into the running mean in order — bit-identical to num_frames calls.
    FrameRecord &enqueue_frame(const float *pos, const float *dir, const float *up, float fovy, bool camera_changed,
                               uint32_t num_frames = 1, bool blocking = false)
    {
        if (num_frames < 1 || num_frames > 64) {
            throw std::runtime_error("render: num_frames must be in [1, 64]");
        }
        if (fb_w == 0) {
            throw std::runtime_error("render: initialize() has not been called");
        }
        if (!have_scene) {
            throw std::runtime_error("render: set_scene() has not been called");
        }
        make_current();
        if (camera_changed) {
            frame_id = 0;  // render_embree.cpp:145-147
        }
        const crt::ViewParams view = view_params(pos, dir, up, fovy);
        const size_t npaths = (size_t)npx_local * spp * num_frames;
        if (npaths >= 0x7fffffffull) {
            throw std::runtime_error("render: more than 2^31 paths in flight on one device");
        }
        ensure_path_buffers(npaths);
        frame_far_first = pick_far_first(blocking);
        const crt::DeviceScene sc = device_scene();
        crt::FrameLayout fl = frame_layout();
        fl.frames = num_frames;
        const crt::PathState ps = path_state();
        if (record_pool.empty()) {
            record_pool.emplace_back(new FrameRecord());
        }
        in_flight.push_back(std::move(record_pool.back()));
        record_pool.pop_back();
        FrameRecord &rec = *in_flight.back();
        rec.num_events = 0;
        rec.launches = 0;
        rec.all_stages = stage_events || (any_far_first == 2 && !auto_decided);  // (the shadow-order trial reads the traversal stage)

        const bool shared_dst = frame_exported && world_size > 1, shared_src = peer_accum_full != nullptr && !frame_exported;
        if (shared_dst || shared_src) {
            ensure_sync_state();
        }
        rec.mark(stream, -1);
        if (shared_dst && share_seq > 0) {
            // starting the next wavefront: the assembling rank is done with the previous frame, peers may overwrite it
            crt::k_set_word<<<1, 32, 0, stream>>>(sync_words() + crt::kSyncConsumed, share_seq);
        }
        CUDA_CHECK(cudaMemsetAsync(d_counters.ptr, 0, crt::kNumCounters * sizeof(uint32_t), stream));
        if (count_traversal) {
            CUDA_CHECK(cudaMemsetAsync(d_trav_counters.ptr, 0, 4 * sizeof(unsigned long long), stream));
        }
        if (npaths) {
            const unsigned g256 = (unsigned)((npaths + 255) / 256);
            // k_nee_resolve strides over its queue: two waves of resident blocks cover any queue length
            if (num_sms == 0) {
                CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
            }
            const unsigned g128 = (unsigned)((npaths + 127) / 128);
            const unsigned g_nee = (unsigned)std::min<size_t>((npaths + 255) / 256, (size_t)num_sms * 8 * 2);
            crt::k_raygen<<<g256, 256, 0, stream>>>(view, fl, ps);
            rec.mark(stream, kStRaygen);
            // closest hit of the primary rays, then per bounce: shade -> one traversal launch for this
            // bounce's shadow rays AND the next bounce's continuation rays -> NEE resolve
            launch_traverse(sc, ps, ps.queue[0], ps.counters + crt::kCntQueue, nullptr, ps.counters + crt::kCntWork);
            rec.mark(stream, kStPrimary);
            rec.launches += 2;
            for (int b = 0; b < max_depth; ++b) {
                const uint32_t *qin = ps.queue[b & 1];
                uint32_t *qout = ps.queue[(b + 1) & 1];
                if (shade_sort == 2 || (shade_sort == 1 && b > 0)) {
                    // (its time counts as shading: the next mark closes the stage)
                    rec.launches += sort_shade_queue(sc, ps, qin, b, npaths);
                    qin = d_queue_sorted.ptr;
                }
                if (sc.tex_objects) {
                    crt::k_shade<true><<<g128, 128, 0, stream>>>(sc, ps, qin, qout, b, max_depth);
                } else {
                    crt::k_shade<false><<<g128, 128, 0, stream>>>(sc, ps, qin, qout, b, max_depth);
                }
                rec.mark(stream, kStShade);
                const bool last = b + 1 == max_depth;
                launch_traverse(sc, ps, qout, last ? nullptr : ps.counters + crt::kCntQueue + b + 1,
                                ps.counters + crt::kCntShadow + b, ps.counters + crt::kCntWork + 1 + b);
                rec.mark(stream, kStTraverse);
                crt::k_nee_resolve<<<g_nee, 256, 0, stream>>>(ps, qin, b);
                rec.mark(stream, kStNee);
                rec.launches += 3;
            }
            const unsigned gpx = (unsigned)((npx_local + 255) / 256);
            // where the resolved pixels also go as a full row-major frame: this renderer's own full frame (single
            // GPU, or the assembling rank of a peer-written frame), the assembling rank's full frame over NVLink
            // (crtc_import_frame), or nowhere (tiles are gathered afterwards: crtc_local_buffers)
            float *full_accum = nullptr;
            uint32_t *full_img = nullptr;
            if (world_size == 1 || frame_exported) {
                full_accum = d_accum_full.ptr;
                full_img = d_img_full.ptr;
            } else if (peer_accum_full) {
                full_accum = peer_accum_full;
                full_img = peer_img_full;
            }
            uint32_t *arrive = nullptr;
            if (shared_src) {
                if (share_seq > 0) {  // not before the assembling rank is done with the frame this one overwrites
                    crt::k_wait_words<<<1, 32, 0, stream>>>(peer_sync_words() + crt::kSyncConsumed, 1u, 0xffffffffu, share_seq, d_sync_err);
                }
                arrive = peer_sync_words() + rank;
            }
            crt::k_resolve<<<gpx, 256, 0, stream>>>(fl, ps, frame_id, d_accum_local.ptr, d_img_local.ptr, full_accum, full_img,
                                                   d_done_counter.ptr, arrive, share_seq + 1u);
            rec.launches += 1;
            rec.mark(stream, kStResolve);
        } else if (shared_src) {
            // a rank that owns no tile of this frame (more ranks than tiles) still has to say "done"
            crt::k_set_word<<<1, 32, 0, stream>>>(peer_sync_words() + rank, share_seq + 1u);
        }
        CUDA_CHECK(cudaMemcpyAsync(rec.h_counters, d_counters.ptr, crt::kNumCounters * sizeof(uint32_t),
                                   cudaMemcpyDeviceToHost, stream));
        if (count_traversal) {
            CUDA_CHECK(cudaMemcpyAsync(rec.h_trav, d_trav_counters.ptr, 4 * sizeof(unsigned long long),
                                       cudaMemcpyDeviceToHost, stream));
        }
        CUDA_CHECK(cudaGetLastError());
        frame_id += num_frames;
        rec.frames = num_frames;
        if (shared_dst || shared_src) {
            ++share_seq;
        }
        return rec;
    }

    // After a stream synchronisation: folds every in-flight frame's events and counters into
    // `stage_sum` / `counter_sum` (both optional) and leaves the last frame's values in
    // stage_ms / counters_out. Returns the number of frames collected.
    uint32_t collect(float *stage_sum, uint64_t *counter_sum)
    {
        uint32_t n = 0;
        for (auto &recp : in_flight) {
            FrameRecord &rec = *recp;
            for (int s = 0; s < kNumStages; ++s) {
                stage_ms[s] = 0.f;
            }
            for (size_t i = 1; rec.all_stages && i < rec.num_events; ++i) {
                float ms = 0.f;
                CUDA_CHECK(cudaEventElapsedTime(&ms, rec.events[i - 1], rec.events[i]));
                stage_ms[rec.event_stage[i]] += ms;
            }
            if (rec.num_events > 1) {
                CUDA_CHECK(cudaEventElapsedTime(&stage_ms[kStFrame], rec.events[0], rec.events[rec.num_events - 1]));
            }
            uint64_t closest = 0, shadow = 0;
            for (int b = 0; b < max_depth; ++b) {
                closest += rec.h_counters[crt::kCntQueue + b];
                shadow += rec.h_counters[crt::kCntShadow + b];
            }
            counters_out[0] = closest;
            counters_out[1] = shadow;
            counters_out[2] = rec.launches;
            counters_out[3] = count_traversal ? rec.h_trav[0] : 0;
            counters_out[4] = count_traversal ? rec.h_trav[1] : 0;
            counters_out[5] = rec.h_counters[crt::kCntQueue];
            counters_out[6] = count_traversal ? rec.h_trav[2] : 0;
            counters_out[7] = count_traversal ? rec.h_trav[3] : 0;
            for (int s = 0; stage_sum && s < kNumStages; ++s) {
                stage_sum[s] += stage_ms[s];
            }
            for (int k = 0; counter_sum && k < 8; ++k) {
                counter_sum[k] += counters_out[k];
            }
            n += rec.frames;
        }
        for (auto &recp : in_flight) {
            record_pool.push_back(std::move(recp));
        }
        in_flight.clear();
        return n;
    }

    // RenderBackend::render: enqueue, wait, collect (+ optional readback of img).
    void render(const float *pos, const float *dir, const float *up, float fovy, bool camera_changed, bool readback,
                uint32_t *img, crt_render_stats_t *stats)
    {
        enqueue_frame(pos, dir, up, fovy, camera_changed, 1, true);
        if (readback && img && world_size == 1) {
            pin_img(img, (size_t)fb_w * fb_h * 4);
            CUDA_CHECK(cudaMemcpyAsync(img, d_img_full.ptr, (size_t)fb_w * fb_h * 4, cudaMemcpyDeviceToHost, stream));
        }
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        check_sync_error();
        collect(nullptr, nullptr);
        auto_tune_step();
        if (stats) {
            const uint64_t rays = counters_out[0] + counters_out[1];
            stats->render_time = stage_ms[kStFrame];
            stats->num_rays = rays;
            stats->rays_per_second =
                stage_ms[kStFrame] > 0.f ? (float)((double)rays / (stage_ms[kStFrame] * 1.0e-3)) : 0.f;
        }
    }

    // crtc_sync: waits for every frame enqueued with crtc_render_async and returns their totals.
    uint32_t sync_frames(crt_render_stats_t *total, float *stage_sum, uint64_t *counter_sum)
    {
        make_current();
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        check_sync_error();
        float ssum[kNumStages] = {0};
        uint64_t csum[8] = {0};
        const uint32_t n = collect(ssum, csum);
        if (total) {
            total->render_time = ssum[kStFrame];
            total->num_rays = csum[0] + csum[1];
            total->rays_per_second = ssum[kStFrame] > 0.f ? (float)((double)total->num_rays / (ssum[kStFrame] * 1.0e-3)) : 0.f;
        }
        for (int s = 0; stage_sum && s < kNumStages; ++s) {
            stage_sum[s] = ssum[s];
        }
        for (int k = 0; counter_sum && k < 8; ++k) {
            counter_sum[k] = csum[k];
        }
        return n;
    }

    // ---- kernel-level access ----
    void upload_rays(const float *rays, uint64_t n)
    {
        ensure_path_buffers(std::max<size_t>(n, path_capacity));
        std::vector<float4> o(n), d(n);
        for (uint64_t i = 0; i < n; ++i) {
            o[i] = make_float4(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2], rays[8 * i + 3]);
            d[i] = make_float4(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6], rays[8 * i + 7]);
        }
        CUDA_CHECK(cudaMemcpyAsync(d_ray_o.ptr, o.data(), n * sizeof(float4), cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaMemcpyAsync(d_ray_d.ptr, d.data(), n * sizeof(float4), cudaMemcpyHostToDevice, stream));
        const uint32_t cnt = (uint32_t)n;
        CUDA_CHECK(cudaMemcpyAsync(d_counters.ptr, &cnt, 4, cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    void launch_closest(uint64_t)
    {
        CUDA_CHECK(cudaMemsetAsync(d_counters.ptr + crt::kCntWork, 0, sizeof(uint32_t), stream));
        launch_traverse(device_scene(), path_state(), nullptr, d_counters.ptr, nullptr, d_counters.ptr + crt::kCntWork);
    }

    // shadow-ray layout for the any-hit kernel: sray_o = org|tfar, sray_d = dir|bits(index)
    void stage_any(uint64_t n, const float *rays)
    {
        std::vector<float4> o(n), d(n);
        for (uint64_t i = 0; i < n; ++i) {
            o[i] = make_float4(rays[8 * i], rays[8 * i + 1], rays[8 * i + 2], rays[8 * i + 7]);
            uint32_t idx = (uint32_t)i;
            float fi;
            std::memcpy(&fi, &idx, 4);
            d[i] = make_float4(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6], fi);
        }
        CUDA_CHECK(cudaMemcpyAsync(d_sray_o.ptr, o.data(), n * sizeof(float4), cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaMemcpyAsync(d_sray_d.ptr, d.data(), n * sizeof(float4), cudaMemcpyHostToDevice, stream));
        const uint32_t cnt = (uint32_t)n;
        CUDA_CHECK(cudaMemcpyAsync(d_counters.ptr, &cnt, 4, cudaMemcpyHostToDevice, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
    }

    void launch_any(uint64_t)
    {
        CUDA_CHECK(cudaMemsetAsync(d_counters.ptr + crt::kCntWork, 0, sizeof(uint32_t), stream));
        launch_traverse(device_scene(), path_state(), nullptr, nullptr, d_counters.ptr, d_counters.ptr + crt::kCntWork);
    }

    void require_scene() const
    {
        if (!have_scene) {
            throw std::runtime_error("set_scene() has not been called");
        }
    }

    void trace_closest(const float *rays, uint64_t n, float *hits)
    {
        require_scene();
        make_current();
        if (n == 0) {
            return;
        }
        if (n > 0x3fffffffull) {
            throw std::runtime_error("trace_closest: too many rays in one batch");
        }
        upload_rays(rays, n);
        launch_closest(n);
        std::vector<float4> h(n);
        CUDA_CHECK(cudaMemcpyAsync(h.data(), d_hit.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        for (uint64_t i = 0; i < n; ++i) {
            uint32_t tri;
            std::memcpy(&tri, &h[i].w, 4);
            const uint32_t flat = tri == crt::kMiss ? crt::kMiss : leaf_flat_ids[tri];
            hits[4 * i] = h[i].x;
            hits[4 * i + 1] = h[i].y;
            hits[4 * i + 2] = h[i].z;
            std::memcpy(&hits[4 * i + 3], &flat, 4);
        }
    }

    void trace_any(const float *rays, uint64_t n, uint8_t *occluded)
    {
        require_scene();
        make_current();
        if (n == 0) {
            return;
        }
        if (n > 0x3fffffffull) {
            throw std::runtime_error("trace_any: too many rays in one batch");
        }
        ensure_path_buffers(std::max<size_t>((n + 1) / 2, path_capacity));
        // the any-hit kernel uses tnear = EPSILON like the reference's shadow rays; a batch ray
        // with another tnear is moved along its direction so the interval is preserved
        std::vector<float> adj(rays, rays + 8 * n);
        for (uint64_t i = 0; i < n; ++i) {
            const float shift = adj[8 * i + 3] - crt::kEpsilon;
            if (shift != 0.f) {
                for (int k = 0; k < 3; ++k) {
                    adj[8 * i + k] += shift * adj[8 * i + 4 + k];
                }
                adj[8 * i + 7] -= shift;
            }
        }
        stage_any(n, adj.data());
        launch_any(n);
        std::vector<uint8_t> vis(n);
        CUDA_CHECK(cudaMemcpyAsync(vis.data(), d_vis.ptr, n, cudaMemcpyDeviceToHost, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        for (uint64_t i = 0; i < n; ++i) {
            occluded[i] = vis[i] ? 0 : 1;
        }
    }

    float bench_trace(const float *rays, uint64_t n, bool any_hit, int iters)
    {
        require_scene();
        make_current();
        if (n == 0 || iters <= 0) {
            return 0.f;
        }
        if (any_hit) {
            ensure_path_buffers(std::max<size_t>((n + 1) / 2, path_capacity));
            stage_any(n, rays);
        } else {
            upload_rays(rays, n);
        }
        cudaEvent_t e0, e1;
        CUDA_CHECK(cudaEventCreate(&e0));
        CUDA_CHECK(cudaEventCreate(&e1));
        for (int w = 0; w < 3; ++w) {
            any_hit ? launch_any(n) : launch_closest(n);
        }
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaEventRecord(e0, stream));
        for (int i = 0; i < iters; ++i) {
            any_hit ? launch_any(n) : launch_closest(n);
        }
        CUDA_CHECK(cudaEventRecord(e1, stream));
        CUDA_CHECK(cudaStreamSynchronize(stream));
        CUDA_CHECK(cudaGetLastError());
        float ms = 0.f;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        return ms / iters;
    }

    // tile-id lists of the other ranks, cached on the device (keyed by world_size * 4096 + rank)
    std::vector<std::pair<int, std::unique_ptr<DeviceBuffer<uint32_t>>>> assemble_tiles;

    // Stream-ordered (no host synchronisation): the caller orders it after the gather.
    void assemble_rank(int src_rank, int ws, const void *accum_dev, const void *img_dev)
    {
        make_current();
        const int key = ws * 4096 + src_rank;
        DeviceBuffer<uint32_t> *d_tiles = nullptr;
        for (auto &e : assemble_tiles) {
            if (e.first == key) {
                d_tiles = e.second.get();
            }
        }
        if (!d_tiles) {
            std::vector<uint32_t> tiles;
            for (uint32_t t = 0; t < ntx * nty; ++t) {
                if ((int)(t % (uint32_t)ws) == src_rank) {
                    tiles.push_back(t);
                }
            }
            assemble_tiles.emplace_back(key, std::make_unique<DeviceBuffer<uint32_t>>());
            d_tiles = assemble_tiles.back().second.get();
            d_tiles->upload(tiles.data(), tiles.size(), stream);
            CUDA_CHECK(cudaStreamSynchronize(stream));  // `tiles` is a host temporary
        }
        if (d_tiles->count == 0) {
            return;
        }
        crt::FrameLayout f = frame_layout();
        f.tile_ids = d_tiles->ptr;
        f.npx_local = (uint32_t)d_tiles->count * crt::kTilePixels;
        const unsigned g = (unsigned)((f.npx_local + 255) / 256);
        crt::k_assemble<<<g, 256, 0, stream>>>(f, static_cast<const float *>(accum_dev),
                                              static_cast<const uint32_t *>(img_dev), d_accum_full.ptr, d_img_full.ptr);
        CUDA_CHECK(cudaGetLastError());
    }
};

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
static void require_renderer(const crtc_renderer *r)
{
    if (!r) {
        throw std::runtime_error("renderer handle is NULL");
    }
}

#define CRTC_TRY(body)                      \
    try {                                   \
        body;                               \
        return 0;                           \
    } catch (const std::exception &e) {     \
        g_last_error = e.what();            \
        return 1;                           \
    } catch (...) {                         \
        g_last_error = "unknown exception"; \
        return 1;                           \
    }

extern "C" {

const char *crtc_last_error(void)
{
    return g_last_error.c_str();
}

const char *crtc_name(void)
{
    return "CUDA wavefront (B200, BVH8)";
}

int crtc_create(crtc_renderer **out, int device)
{
    CRTC_TRY({
        if (!out) {
            throw std::runtime_error("crtc_create: out is NULL");
        }
        *out = nullptr;
        int n = 0;
        cudaError_t err = cudaGetDeviceCount(&n);
        if (err != cudaSuccess || n == 0) {
            throw std::runtime_error(std::string("crtc_create: no usable CUDA device (") +
                                     (err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0") +
                                     "); this backend has no CPU fallback");
        }
        if (device < 0 || device >= n) {
            throw std::runtime_error("crtc_create: device ordinal out of range");
        }
        std::unique_ptr<crtc_renderer> r(new crtc_renderer());
        r->device = device;
        r->make_current();
        CUDA_CHECK(cudaStreamCreateWithFlags(&r->own_stream, cudaStreamNonBlocking));
        r->stream = r->own_stream;
        *out = r.release();
    })
}

void crtc_destroy(crtc_renderer *r)
{
    if (r) {
        cudaSetDevice(r->device);
        delete r;
    }
}

int crtc_set_option(crtc_renderer *r, const char *key, int64_t value)
{
    CRTC_TRY({ require_renderer(r);
        const std::string k = key ? key : "";
        if (k == "max_depth") {
            if (value < 1 || value > crt::kMaxDepthSupported) {
                throw std::runtime_error("max_depth must be in [1, 16]");
            }
            r->max_depth = (int)value;
        } else if (k == "rank") {
            r->rank = (int)value;
        } else if (k == "world_size") {
            if (value < 1) {
                throw std::runtime_error("world_size must be >= 1");
            }
            r->world_size = (int)value;
        } else if (k == "bvh_threads") {
            r->bvh_threads = (int)value;
        } else if (k == "tri_pass_defer") {
            if (value != 0 && value != 16 && value != 24) {
                throw std::runtime_error("tri_pass_defer must be 0, 16 or 24");
            }
            r->tri_pass_defer = (int)value;
        } else if (k == "shade_sort") {
            if (value < 0 || value > 2) {
                throw std::runtime_error("shade_sort must be 0, 1 or 2");
            }
            r->shade_sort = (int)value;
        } else if (k == "hw_textures") {
            r->hw_textures = value != 0;  // takes effect at the next crtc_set_scene
        } else if (k == "stage_events") {
            r->stage_events = value != 0;
        } else if (k == "bvh_top_smem") {
            r->bvh_top_smem = value != 0;
        } else if (k == "pin_host_buffers") {
            r->pin_host_buffers = value != 0;
            if (!r->pin_host_buffers) {
                r->unpin_img();
            }
        } else if (k == "pin_read_img") {
            r->pin_repeated_reads = value != 0;
        } else if (k == "bvh_ploc_tail") {
            r->ploc_tail = value != 0;
        } else if (k == "bvh_ploc_radius") {
            if (value < 1 || value > crt::kPlocMaxRadius) {
                throw std::runtime_error("bvh_ploc_radius must be in [1, 32]");
            }
            r->ploc_radius = (int)value;
        } else if (k == "bvh_builder") {
            if (value < 0 || value > 2) {
                throw std::runtime_error("bvh_builder must be 0 (host), 1 (device, PLOC) or 2 (device, LBVH)");
            }
            r->bvh_builder = (int)value;
        } else if (k == "count_traversal") {
            r->count_traversal = value != 0;
        } else if (k == "refill_idle") {
            r->refill_idle = (int)std::min<int64_t>(std::max<int64_t>(value, 1), 32);
        } else if (k == "any_far_first") {
            if (value < 0 || value > 2) {
                throw std::runtime_error("any_far_first must be 0 (off), 1 (on) or 2 (auto)");
            }
            r->any_far_first = (int)value;
            r->auto_frames = 0;
            r->auto_decided = false;
        } else {
            throw std::runtime_error("unknown option '" + k + "'");
        }
        // (rank may be set before world_size: the pair is validated in crtc_initialize)
    })
}

int crtc_get_option(crtc_renderer *r, const char *key, int64_t *value)
{
    CRTC_TRY({ require_renderer(r);
        const std::string k = key ? key : "";
        if (!value) {
            throw std::runtime_error("crtc_get_option: value is null");
        }
        if (k == "max_depth") {
            *value = r->max_depth;
        } else if (k == "rank") {
            *value = r->rank;
        } else if (k == "world_size") {
            *value = r->world_size;
        } else if (k == "bvh_threads") {
            *value = r->bvh_threads;
        } else if (k == "tri_pass_defer") {
            *value = r->tri_pass_defer;
        } else if (k == "hw_textures") {
            *value = r->hw_textures ? 1 : 0;
        } else if (k == "shade_sort") {
            *value = r->shade_sort;
        } else if (k == "bvh_builder") {
            *value = r->bvh_builder;
        } else if (k == "bvh_builder_fallbacks") {
            *value = r->builder_fallbacks;
        } else if (k == "bvh_build_rounds") {
            *value = r->build_rounds;  // PLOC rounds of the last device build
        } else if (k == "count_traversal") {
            *value = r->count_traversal ? 1 : 0;
        } else if (k == "refill_idle") {
            *value = r->refill_idle;
        } else if (k == "any_far_first") {
            *value = r->any_far_first;
        } else if (k == "any_far_first_decision") {
            // the order frames are rendered with from now on: 0 near-first, 1 far-first, -1 auto mode still undecided
            *value = r->any_far_first != 2 ? (r->any_far_first == 1 ? 1 : 0) : (r->auto_decided ? (r->auto_choice ? 1 : 0) : -1);
        } else {
            throw std::runtime_error("unknown option '" + k + "'");
        }
    })
}

int crtc_set_stream(crtc_renderer *r, void *cuda_stream)
{
    CRTC_TRY({ require_renderer(r); r->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : r->own_stream; })
}

int crtc_initialize(crtc_renderer *r, int fb_width, int fb_height)
{
    CRTC_TRY({ require_renderer(r);
        if (r->rank < 0 || r->rank >= r->world_size) {
            throw std::runtime_error("rank must be in [0, world_size)");
        }
        r->initialize(fb_width, fb_height);
    })
}

int crtc_set_scene(crtc_renderer *r, const crt_scene_t *scene)
{
    CRTC_TRY({ require_renderer(r);
        if (!scene) {
            throw std::runtime_error("crtc_set_scene: scene is NULL");
        }
        r->set_scene(scene);
    })
}

int crtc_render(crtc_renderer *r, const float *pos, const float *dir, const float *up, float fovy, int camera_changed,
                int readback_framebuffer, uint32_t *img, crt_render_stats_t *stats)
{
    CRTC_TRY({ require_renderer(r); r->render(pos, dir, up, fovy, camera_changed != 0, readback_framebuffer != 0, img, stats); })
}

int crtc_render_async(crtc_renderer *r, const float *pos, const float *dir, const float *up, float fovy,
                      int camera_changed, uint32_t num_frames)
{
    CRTC_TRY({ require_renderer(r); r->enqueue_frame(pos, dir, up, fovy, camera_changed != 0, num_frames); })
}

int crtc_sync(crtc_renderer *r, crt_render_stats_t *total, float *stage_ms_sum, uint64_t *counters_sum,
              uint32_t *num_frames)
{
    CRTC_TRY({ require_renderer(r);
        const uint32_t n = r->sync_frames(total, stage_ms_sum, counters_sum);
        if (num_frames) {
            *num_frames = n;
        }
    })
}

int crtc_read_accum(crtc_renderer *r, float *rgb_out)
{
    CRTC_TRY({ require_renderer(r);
        r->make_current();
        r->frame_wait();
        CUDA_CHECK(cudaMemcpyAsync(rgb_out, r->d_accum_full.ptr, (size_t)r->fb_w * r->fb_h * 3 * sizeof(float),
                                   cudaMemcpyDeviceToHost, r->stream));
        CUDA_CHECK(cudaStreamSynchronize(r->stream));
        r->check_sync_error();
    })
}

int crtc_read_img(crtc_renderer *r, uint32_t *img)
{
    CRTC_TRY({ require_renderer(r);
        r->make_current();
        r->frame_wait();  // (a shared frame: every rank's tiles first; no-op otherwise)
        if (img == r->pinned_img || r->pin_repeated_reads) {  // (a caller that reads into the same buffer every frame)
            r->pin_img(img, (size_t)r->fb_w * r->fb_h * 4);
        }
        CUDA_CHECK(cudaMemcpyAsync(img, r->d_img_full.ptr, (size_t)r->fb_w * r->fb_h * 4, cudaMemcpyDeviceToHost,
                                   r->stream));
        CUDA_CHECK(cudaStreamSynchronize(r->stream));
        r->check_sync_error();
    })
}

int crtc_get_stage_times(crtc_renderer *r, float *ms_out, int n)
{
    if (!r || !ms_out || n <= 0) {
        return 0;
    }
    const int m = std::min(n, (int)kNumStages);
    for (int i = 0; i < m; ++i) {
        ms_out[i] = r->stage_ms[i];
    }
    return m;
}

int crtc_get_counters(crtc_renderer *r, uint64_t *out, int n)
{
    if (!r || !out || n <= 0) {
        return 0;
    }
    const int m = std::min(n, 8);
    for (int i = 0; i < m; ++i) {
        out[i] = r->counters_out[i];
    }
    return m;
}

int crtc_get_scene_info(crtc_renderer *r, double *out, int n)
{
    if (!r || !out || n <= 0) {
        return 0;
    }
    const int m = std::min(n, 12);
    for (int i = 0; i < m; ++i) {
        out[i] = r->scene_info[i];
    }
    return m;
}

int crtc_trace_closest(crtc_renderer *r, const float *rays, uint64_t n, float *hits)
{
    CRTC_TRY({ require_renderer(r); r->trace_closest(rays, n, hits); })
}

int crtc_trace_any(crtc_renderer *r, const float *rays, uint64_t n, uint8_t *occluded)
{
    CRTC_TRY({ require_renderer(r); r->trace_any(rays, n, occluded); })
}

int crtc_bench_trace(crtc_renderer *r, const float *rays_host, uint64_t n, int any_hit, int iters, float *ms_out)
{
    CRTC_TRY({ require_renderer(r); *ms_out = r->bench_trace(rays_host, n, any_hit != 0, iters); })
}

int crtc_local_buffers(crtc_renderer *r, void **accum_dev, void **img_dev, uint32_t *num_local_tiles)
{
    CRTC_TRY({ require_renderer(r);
        *accum_dev = r->d_accum_local.ptr;
        *img_dev = r->d_img_local.ptr;
        *num_local_tiles = (uint32_t)r->local_tiles.size();
    })
}

int crtc_assemble_rank(crtc_renderer *r, int src_rank, int world_size, const void *accum_dev, const void *img_dev)
{
    CRTC_TRY({ require_renderer(r); r->assemble_rank(src_rank, world_size, accum_dev, img_dev); })
}

int crtc_export_frame(crtc_renderer *r, void *handles_out)
{
    CRTC_TRY({ require_renderer(r);
        if (!handles_out) {
            throw std::runtime_error("crtc_export_frame: null output");
        }
        r->export_frame(handles_out);
    })
}

int crtc_import_frame(crtc_renderer *r, const void *handles)
{
    CRTC_TRY({ require_renderer(r); r->import_frame(handles); })
}

int crtc_copy_img_to_array(crtc_renderer *r, void *cuda_array)
{
    CRTC_TRY({ require_renderer(r);
        if (!cuda_array) {
            throw std::runtime_error("crtc_copy_img_to_array: array is NULL");
        }
        r->make_current();
        r->frame_wait();
        CUDA_CHECK(cudaMemcpy2DToArrayAsync(static_cast<cudaArray_t>(cuda_array), 0, 0, r->d_img_full.ptr, (size_t)r->fb_w * 4,
                                            (size_t)r->fb_w * 4, (size_t)r->fb_h, cudaMemcpyDeviceToDevice, r->stream));
        CUDA_CHECK(cudaStreamSynchronize(r->stream));
        r->check_sync_error();
    })
}

int crtc_frame_wait(crtc_renderer *r)
{
    CRTC_TRY({ require_renderer(r); r->frame_wait(); })
}

int crtc_share_frame(crtc_renderer *dst, crtc_renderer *src)
{
    CRTC_TRY({
        if (!dst || !src) {
            throw std::runtime_error("crtc_share_frame: null renderer");
        }
        src->share_frame_of(dst);
    })
}
}
