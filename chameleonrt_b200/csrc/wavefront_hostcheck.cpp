// wavefront_hostcheck.cpp — TEST-ONLY shared library (libcrt_wavefront_hostcheck.so).
//
// Runs the product's KERNEL SOURCE on the host: kernels.cuh (k_raygen, k_shade, k_nee_resolve, k_resolve —
// the whole per-bounce logic: hit decode, material unpack, next-event estimation with MIS, continuation,
// Russian roulette, queue compaction, running-mean resolve, frames in flight) is compiled with g++ and every
// CUDA thread is executed as a plain function call, one "thread" = lane 0 of its own one-lane warp (the warp
// intrinsics degenerate to the identity, atomics to plain arithmetic). The closest-hit / any-hit stages use the
// host instantiation of bvh8_traverse.h (the warp-cooperative scheduling of k_traverse is GPU-only; its per-ray
// arithmetic is the same header). The launch sequence below mirrors crtc_renderer::enqueue_frame
// (crt_cuda_core.cu). What it buys: on a machine without a GPU, tests/test_wavefront_host.py renders frames
// through the kernels' own source and compares them with the oracle / the reference build at rounding level.
// It is not linked into libcrt_cuda_core.so and render() cannot reach it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>  // vector types only

// ---- the CUDA execution environment, for one lane ----
namespace {
struct Idx3 {
    unsigned x = 0, y = 0, z = 0;
};
}  // namespace
static thread_local Idx3 threadIdx, blockIdx, blockDim, gridDim;

template <typename T>
static inline T __ldg(const T *p)
{
    return *p;
}
static inline uint32_t __float_as_uint(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
static inline float __uint_as_float(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
static inline float __uint2float_rn(uint32_t u) { return (float)u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __ballot_sync(unsigned, bool pred) { return pred ? 1u : 0u; }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int)
{
    return v;
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int)
{
    return v;
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T, int)
{
    return T(0);  // lanes 1..31 do not exist: they contribute nothing to a reduction
}
static inline void __syncwarp() {}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) { return v; }
static inline unsigned __reduce_or_sync(unsigned, unsigned v) { return v; }
// (declarations for the block-level kernels of kernels.cuh, which are templates and never instantiated in this build)
static inline void __syncthreads() {}
static inline void __threadfence_system() {}
static inline unsigned long long crt_host_clock_ns() { return 0ull; }
template <typename T>
static inline unsigned __match_any_sync(unsigned, T)
{
    return 1u;
}
template <typename T>
static inline T atomicAdd(T *p, T v)
{
    const T old = *p;
    *p = old + v;
    return old;
}
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    const unsigned long long old = *p;
    *p = std::min(old, v);
    return old;
}
using std::max;
using std::min;
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif

#include "bvh8.h"
#include "host_scene.h"
#include "kernels.cuh"

namespace {

std::string g_err;

struct HostWavefront {
    // scene (device-format records, host memory)
    std::vector<float4> nodes, tris, shade, materials, lights;
    std::vector<uint32_t> texels;
    std::vector<crt::DevTex> tex;
    uint32_t num_lights = 0, spp = 1;
    // framebuffer
    int fb_w = 0, fb_h = 0;
    uint32_t ntx = 0, npx_local = 0;
    std::vector<uint32_t> tile_ids;
    std::vector<float> accum_local, accum_full;
    std::vector<uint32_t> img_local, img_full;
    uint32_t frame_id = 0;
    int max_depth = 5;
    int rank = 0, world = 1;
    // path state
    std::vector<float4> ray_o, ray_d, hit, thr_rng, radiance, nee_T, nee_l1, nee_l2, sray_o, sray_d;
    std::vector<uint8_t> vis;
    std::vector<uint32_t> queue0, queue1, counters;
    unsigned long long trav_counters[4] = {0, 0, 0, 0};
    uint64_t last_rays = 0;

    crt::DeviceScene device_scene() const
    {
        crt::DeviceScene sc;
        sc.nodes = nodes.data();
        sc.tris = tris.data();
        sc.shade = shade.data();
        sc.materials = materials.data();
        sc.lights = lights.data();
        sc.texels = texels.data();
        sc.tex = tex.data();
        sc.num_lights = num_lights;
        sc.float_one = 0x3F800000u;
        return sc;
    }

    void set_scene(const crt_scene_t *scene)
    {
        // crtc_renderer::set_scene
        crt::HostScene hs;
        crt::flatten_scene(scene, hs, 0);
        crt::Bvh8 bvh;
        crt::build_bvh8(hs.tri_verts.data(), hs.num_tris(), 0, bvh);
        std::vector<float> tri_records;
        std::vector<crt::TriShade> sh;
        crt::pack_triangles(hs, bvh, tri_records, sh, 0);
        nodes.resize(bvh.nodes.size() * 5);
        std::memcpy(nodes.data(), bvh.nodes.data(), bvh.nodes.size() * 80);
        if (tri_records.empty()) {
            tri_records.assign(12, 0.f);
            sh.resize(1);
            std::memset(&sh[0], 0, sizeof(crt::TriShade));
        }
        tris.resize(tri_records.size() / 4);
        std::memcpy(tris.data(), tri_records.data(), tri_records.size() * 4);
        shade.resize(sh.size() * 3);
        std::memcpy(shade.data(), sh.data(), sh.size() * sizeof(crt::TriShade));
        if (hs.materials.empty()) {
            hs.materials.resize(1);
            std::memset(&hs.materials[0], 0, sizeof(crt_material_t));
        }
        materials.resize(hs.materials.size() * 4);
        std::memcpy(materials.data(), hs.materials.data(), hs.materials.size() * 64);
        lights.resize(hs.lights.size() * 5);
        std::memcpy(lights.data(), hs.lights.data(), hs.lights.size() * 80);
        num_lights = (uint32_t)hs.lights.size();
        texels.assign(hs.texels.begin(), hs.texels.end());
        if (texels.empty()) {
            texels.assign(1, 0u);
        }
        tex.assign(std::max<size_t>(1, hs.tex_desc.size()), crt::DevTex{0, 0, 0, 0});
        for (size_t i = 0; i < hs.tex_desc.size(); ++i) {
            tex[i] = crt::DevTex{hs.tex_desc[i].offset, hs.tex_desc[i].width, hs.tex_desc[i].height, 0};
        }
        spp = std::max<uint32_t>(1u, hs.samples_per_pixel);
        frame_id = 0;
        std::fill(accum_local.begin(), accum_local.end(), 0.f);
    }

    void initialize(int w, int h)
    {
        // crtc_renderer::initialize
        fb_w = w;
        fb_h = h;
        frame_id = 0;
        ntx = w / crt::kTile + (w % crt::kTile != 0 ? 1 : 0);
        const uint32_t nty = h / crt::kTile + (h % crt::kTile != 0 ? 1 : 0);
        tile_ids.clear();
        for (uint32_t t = 0; t < ntx * nty; ++t) {
            if ((int)(t % (uint32_t)world) == rank) {
                tile_ids.push_back(t);
            }
        }
        npx_local = (uint32_t)tile_ids.size() * crt::kTilePixels;
        accum_local.assign((size_t)npx_local * 3, 0.f);
        img_local.assign(npx_local, 0u);
        accum_full.assign((size_t)w * h * 3, 0.f);
        img_full.assign((size_t)w * h, 0u);
    }

    static float3 glm_normalize(float3 v)
    {
        const float inv = 1.f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
        return make_float3(v.x * inv, v.y * inv, v.z * inv);
    }
    static float3 cross3(float3 a, float3 b)
    {
        return make_float3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
    }
    // crtc_renderer::view_params
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
        v.dir_top_left = make_float3(dir.x - 0.5f * v.dir_du.x - 0.5f * v.dir_dv.x, dir.y - 0.5f * v.dir_du.y - 0.5f * v.dir_dv.y,
                                     dir.z - 0.5f * v.dir_du.z - 0.5f * v.dir_dv.z);
        v.frame_id = frame_id;
        return v;
    }

    template <typename F>
    static void launch(size_t nthreads, const F &kernel)
    {
        blockDim.x = 1;
        threadIdx.x = 0;
        gridDim.x = (unsigned)nthreads;  // (k_shade / k_nee_resolve stride over their queue by gridDim.x * blockDim.x)
        for (size_t i = 0; i < nthreads; ++i) {
            blockIdx.x = (unsigned)i;
            kernel();
        }
    }

    // the work of one k_traverse launch: shadow rays [0, n_any) then closest-hit rays through `queue`
    void traverse(const crt::PathState &ps, const uint32_t *queue, uint32_t n_closest, uint32_t n_any, bool far_first)
    {
        for (uint32_t j = 0; j < n_any; ++j) {
            const float4 o = ps.sray_o[j], d = ps.sray_d[j];
            const crt::Ray ray{o.x, o.y, o.z, crt::kEpsilon, d.x, d.y, d.z, o.w};
            crt::HitRecord h;
            crt::bvh8_trace<true, false>(nodes.data(), tris.data(), ray, h, nullptr, far_first);
            ps.vis[__float_as_uint(d.w)] = h.tri != crt::kMiss ? 0 : 1;
        }
        for (uint32_t j = 0; j < n_closest; ++j) {
            const uint32_t slot = queue ? queue[j] : j;
            const float4 o = ps.ray_o[slot], d = ps.ray_d[slot];
            const crt::Ray ray{o.x, o.y, o.z, o.w, d.x, d.y, d.z, d.w};
            crt::HitRecord h;
            crt::bvh8_trace<false, false>(nodes.data(), tris.data(), ray, h, nullptr);
            const bool hit = h.tri != crt::kMiss;
            ps.hit[slot] = make_float4(h.t, hit ? h.u : 0.f, hit ? h.v : 0.f, __uint_as_float(h.tri));
        }
    }

    // crtc_renderer::enqueue_frame
    void render(const float *pos, const float *dir, const float *up, float fovy, bool camera_changed, uint32_t num_frames,
                bool far_first)
    {
        if (camera_changed) {
            frame_id = 0;
        }
        const crt::ViewParams view = view_params(pos, dir, up, fovy);
        const size_t npaths = (size_t)npx_local * spp * num_frames;
        for (auto *v : {&ray_o, &ray_d, &hit, &thr_rng, &radiance, &nee_T, &nee_l1, &nee_l2}) {
            v->assign(npaths, make_float4(0.f, 0.f, 0.f, 0.f));
        }
        sray_o.assign(2 * npaths, make_float4(0.f, 0.f, 0.f, 0.f));
        sray_d.assign(2 * npaths, make_float4(0.f, 0.f, 0.f, 0.f));
        vis.assign(2 * npaths, 0);
        queue0.assign(npaths, 0u);
        queue1.assign(npaths, 0u);
        counters.assign(crt::kNumCounters, 0u);
        crt::PathState ps;
        ps.ray_o = ray_o.data();
        ps.ray_d = ray_d.data();
        ps.hit = hit.data();
        ps.thr_rng = thr_rng.data();
        ps.radiance = radiance.data();
        ps.nee_T = nee_T.data();
        ps.nee_l1 = nee_l1.data();
        ps.nee_l2 = nee_l2.data();
        ps.sray_o = sray_o.data();
        ps.sray_d = sray_d.data();
        ps.vis = vis.data();
        ps.queue[0] = queue0.data();
        ps.queue[1] = queue1.data();
        ps.counters = counters.data();
        ps.trav_counters = trav_counters;
        const crt::DeviceScene sc = device_scene();
        crt::FrameLayout fl;
        fl.fb_w = fb_w;
        fl.fb_h = fb_h;
        fl.ntx = ntx;
        fl.npx_local = npx_local;
        fl.spp = spp;
        fl.frames = num_frames;
        fl.tile_ids = tile_ids.data();

        if (npaths) {
            launch(npaths, [&] { crt::k_raygen(view, fl, ps); });
            traverse(ps, ps.queue[0], counters[crt::kCntQueue], 0, far_first);
            for (int b = 0; b < max_depth; ++b) {
                uint32_t *qin = ps.queue[b & 1], *qout = ps.queue[(b + 1) & 1];
                launch(npaths, [&] { crt::k_shade<false>(sc, ps, qin, qout, b, max_depth); });
                const bool last = b + 1 == max_depth;
                traverse(ps, qout, last ? 0u : counters[crt::kCntQueue + b + 1], counters[crt::kCntShadow + b], far_first);
                launch(npaths, [&] { crt::k_nee_resolve(ps, qin, b); });
            }
            const bool full = world == 1;
            launch(npx_local, [&] {
                crt::k_resolve(fl, ps, frame_id, accum_local.data(), img_local.data(), full ? accum_full.data() : nullptr,
                               full ? img_full.data() : nullptr, nullptr, nullptr, 0u);
            });
        }
        last_rays = 0;
        for (int b = 0; b < max_depth; ++b) {
            last_rays += counters[crt::kCntQueue + b] + counters[crt::kCntShadow + b];
        }
        frame_id += num_frames;
    }
};

}  // namespace

extern "C" {

const char *crt_wavecheck_last_error() { return g_err.c_str(); }

void *crt_wavecheck_create(const crt_scene_t *scene, int w, int h, int max_depth, int rank, int world)
{
    try {
        auto *r = new HostWavefront();
        r->max_depth = max_depth;
        r->rank = rank;
        r->world = world;
        r->initialize(w, h);
        r->set_scene(scene);
        return r;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

void crt_wavecheck_destroy(void *p) { delete static_cast<HostWavefront *>(p); }

// renders `num_frames` consecutive frames as one wavefront; returns the ray count (closest-hit + occlusion casts)
uint64_t crt_wavecheck_render(void *p, const float *pos, const float *dir, const float *up, float fovy, int camera_changed,
                              uint32_t num_frames, int far_first)
{
    HostWavefront *r = static_cast<HostWavefront *>(p);
    r->render(pos, dir, up, fovy, camera_changed != 0, num_frames, far_first != 0);
    return r->last_rays;
}

// crtc_renderer::assemble_rank: scatter rank `src`'s tile-local buffers into `dst`'s full frame (k_assemble)
void crt_wavecheck_assemble(void *dst_, void *src_)
{
    HostWavefront *dst = static_cast<HostWavefront *>(dst_), *src = static_cast<HostWavefront *>(src_);
    if (src->npx_local == 0) {
        return;
    }
    crt::FrameLayout f;
    f.fb_w = dst->fb_w;
    f.fb_h = dst->fb_h;
    f.ntx = dst->ntx;
    f.npx_local = src->npx_local;
    f.spp = dst->spp;
    f.frames = 1;
    f.tile_ids = src->tile_ids.data();
    HostWavefront::launch(f.npx_local, [&] {
        crt::k_assemble(f, src->accum_local.data(), src->img_local.data(), dst->accum_full.data(), dst->img_full.data());
    });
}

void crt_wavecheck_read(void *p, float *accum_full, uint32_t *img_full)
{
    HostWavefront *r = static_cast<HostWavefront *>(p);
    std::memcpy(accum_full, r->accum_full.data(), r->accum_full.size() * sizeof(float));
    std::memcpy(img_full, r->img_full.data(), r->img_full.size() * sizeof(uint32_t));
}
}
