// simt_env.h — TEST-ONLY: a SIMT execution environment for running the CUDA kernels of kernels.cuh on the host.
// Every CUDA thread of a block is an OS thread; the 32 threads of a warp rendezvous at each warp collective
// (__ballot_sync, __shfl_*_sync, __match_any_sync, __syncwarp are barrier-backed exchanges; __syncthreads is a
// barrier over the block's threads), __shared__ variables become
// function-local statics (blocks run one at a time) and atomics are real atomics. Included BEFORE kernels.cuh by
// simt_hostcheck.cpp (k_traverse alone) and by tests/simt_emu (the whole renderer). Never part of the product.
#pragma once

#include <pthread.h>
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include <vector_types.h>
#include <vector_functions.h>

// ---- SIMT execution environment ----
namespace simt {
struct Idx3 {
    unsigned x = 0, y = 0, z = 0;
};
// Sense-reversing barrier that yields instead of sleeping: a warp's 32 OS threads rendezvous hundreds of times
// per kernel, and futex sleeps (pthread_barrier) cost far more than a handful of sched_yield rounds.
struct YieldBarrier {
    std::atomic<unsigned> arrived{0}, generation{0};
    unsigned parties;
    explicit YieldBarrier(unsigned n = 32) : parties(n) {}
    void wait()
    {
        const unsigned gen = generation.load(std::memory_order_acquire);
        if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == parties) {
            arrived.store(0, std::memory_order_relaxed);
            generation.store(gen + 1, std::memory_order_release);
            return;
        }
        unsigned spins = 0, nap_us = 50;
        while (generation.load(std::memory_order_acquire) == gen) {
            if (++spins < 2000) {
                sched_yield();
            } else {  // nothing is happening (an idle pool between launches): stop burning the CPU
                std::this_thread::sleep_for(std::chrono::microseconds(nap_us));
                nap_us = std::min(nap_us * 2, 5000u);
            }
        }
    }
};
struct Warp {
    YieldBarrier barrier{32};
    unsigned long long slots[32];
    void sync() { barrier.wait(); }
};
static thread_local Warp *warp = nullptr;
static thread_local int lane = 0;
static thread_local YieldBarrier *block_barrier = nullptr;  // all threads of the running block (__syncthreads)

template <typename T>
inline unsigned long long to_bits(T v)
{
    unsigned long long b = 0;
    static_assert(sizeof(T) <= 8, "exchange of at most 8 bytes");
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(unsigned long long b)
{
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
// every lane deposits a value, then reads the lane `src`'s (two rendezvous: publish, consume)
template <typename T>
inline T exchange(T v, int src)
{
    warp->slots[lane] = to_bits(v);
    warp->sync();
    const T r = (src >= 0 && src < 32) ? from_bits<T>(warp->slots[src]) : v;
    warp->sync();
    return r;
}
}  // namespace simt

static thread_local simt::Idx3 threadIdx, blockIdx, blockDim, gridDim;

template <typename T>
static inline T __ldg(const T *p)
{
    return *p;
}
template <typename T>
static inline T __ldcg(const T *p)
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
static inline unsigned __ballot_sync(unsigned, bool pred)
{
    simt::warp->slots[simt::lane] = pred ? 1ull : 0ull;
    simt::warp->sync();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) {
        m |= simt::warp->slots[l] ? (1u << l) : 0u;
    }
    simt::warp->sync();
    return m;
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src)
{
    return simt::exchange(v, src & 31);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int delta)
{
    return simt::exchange(v, simt::lane - delta);  // lanes below `delta` keep their own value
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int delta)
{
    return simt::exchange(v, simt::lane + delta > 31 ? -1 : simt::lane + delta);
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
    simt::warp->slots[simt::lane] = v;
    simt::warp->sync();
    unsigned sum = 0;
    for (int l = 0; l < 32; ++l) {
        sum += (unsigned)simt::warp->slots[l];
    }
    simt::warp->sync();
    return sum;
}
static inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    simt::warp->slots[simt::lane] = v;
    simt::warp->sync();
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) {
        r |= (unsigned)simt::warp->slots[l];
    }
    simt::warp->sync();
    return r;
}
static inline void __syncwarp() { simt::warp->sync(); }
static inline void __syncthreads() { simt::block_barrier->wait(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned long long crt_host_clock_ns()
{
    return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// lanes holding the same value
template <typename T>
static inline unsigned __match_any_sync(unsigned, T v)
{
    simt::warp->slots[simt::lane] = simt::to_bits(v);
    simt::warp->sync();
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) {
        m |= simt::warp->slots[l] == simt::to_bits(v) ? (1u << l) : 0u;
    }
    simt::warp->sync();
    return m;
}
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline uint32_t atomicMin(uint32_t *p, uint32_t v)
{
    uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
static inline uint32_t atomicMax(uint32_t *p, uint32_t v)
{
    uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
using std::max;
using std::min;
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#undef __shared__
#define __shared__ static  // one block at a time: a function-local static is the block's shared memory

// ---- warp-level work profile: what the kernel's instruction stream costs, counted per warp ----
// A warp executes the node phase once per loop iteration if ANY of its lanes has a node to intersect, and one
// triangle pass per 32 pooled (ray, triangle) pairs; lanes without work in a phase are idle issue slots. The
// counters give node-phase executions and triangle passes (~ the warp instructions the hardware issues, the limiter
// ncu shows for this kernel) and the lanes that did useful work in them (~ warp execution efficiency).
namespace simt {
struct Profile {
    std::atomic<unsigned long long> node_phases{0}, node_lanes{0}, tri_passes{0}, tri_lanes{0};
};
static Profile profile;
inline void prof_phase(bool lane_active, std::atomic<unsigned long long> &phases, std::atomic<unsigned long long> &lanes)
{
    const unsigned m = __ballot_sync(0xffffffffu, lane_active);
    if (lane == 0 && m) {
        phases.fetch_add(1, std::memory_order_relaxed);
        lanes.fetch_add((unsigned long long)__builtin_popcount(m), std::memory_order_relaxed);
    }
}
}  // namespace simt
#define CRT_PROF_NODE_PHASE(x) simt::prof_phase((x), simt::profile.node_phases, simt::profile.node_lanes)
#define CRT_PROF_TRI_PASS(x) simt::prof_phase((x), simt::profile.tri_passes, simt::profile.tri_lanes)

// ---- kernel launch: `grid` blocks of `block` threads, one block at a time, on a persistent pool of OS threads ----
namespace simt {
constexpr unsigned kMaxBlock = 256;

struct Pool {
    YieldBarrier start{kMaxBlock + 1}, finish{kMaxBlock + 1};  // kMaxBlock workers + the launching thread
    std::vector<std::thread> workers;
    Warp warps[kMaxBlock / 32];
    YieldBarrier block_bar{kMaxBlock};  // parties = the launch's block size
    const std::function<void()> *kernel = nullptr;
    unsigned block = 0, grid = 0, block_index = 0;
    bool quit = false;

    Pool()
    {
        for (unsigned t = 0; t < kMaxBlock; ++t) {
            workers.emplace_back([this, t] {
                for (;;) {
                    start.wait();
                    if (quit) {
                        return;
                    }
                    if (t < block) {
                        threadIdx.x = t;
                        blockIdx.x = block_index;
                        blockDim.x = block;
                        gridDim.x = grid;
                        warp = &warps[t / 32];
                        lane = (int)(t % 32);
                        block_barrier = &block_bar;
                        (*kernel)();
                    }
                    finish.wait();
                }
            });
        }
    }
    ~Pool()
    {
        quit = true;
        start.wait();
        for (auto &w : workers) {
            w.join();
        }
    }
};

// (internal linkage: two test libraries in one process — libcrt_simt_hostcheck.so and the emulated renderer — must not
// share one pool through the unified function-local static of an inline function: its workers set the thread-local
// threadIdx / warp pointers of the library that created them)
static inline Pool &pool()
{
    static Pool p;
    return p;
}
static inline std::mutex &launch_mutex()
{
    static std::mutex m;
    return m;
}

template <typename Kernel>
inline void launch(unsigned grid, unsigned block, const Kernel &kernel)
{
    if (block == 0 || block > kMaxBlock || block % 32 != 0) {
        throw std::runtime_error("simt::launch: block size must be a multiple of 32, at most 256");
    }
    std::lock_guard<std::mutex> lk(launch_mutex());  // one kernel at a time (streams are synchronous here)
    Pool &p = pool();
    const std::function<void()> fn = kernel;
    p.kernel = &fn;
    p.block = block;
    p.grid = grid;
    p.block_bar.parties = block;
    for (unsigned b = 0; b < grid; ++b) {
        p.block_index = b;
        p.start.wait();
        p.finish.wait();
    }
}
}  // namespace simt
