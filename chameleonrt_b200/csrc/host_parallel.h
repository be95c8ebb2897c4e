// host_parallel.h — the one threading primitive of the host-side set_scene work (scene flattening, BVH8
// build, triangle packing): run f(block) for every block of a fixed decomposition on a few std::threads.
// The decomposition never depends on the thread count, so results do not either.
#pragma once

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

namespace crt {

// Runs f(block) for block in [0, nblocks) on `nthreads` threads (the caller is one of them). f must not throw.
template <typename F>
void parallel_blocks(uint32_t nblocks, int nthreads, const F &f)
{
    nthreads = (int)std::min<uint32_t>((uint32_t)std::max(1, nthreads), nblocks);
    if (nthreads <= 1) {
        for (uint32_t b = 0; b < nblocks; ++b) {
            f(b);
        }
        return;
    }
    std::atomic<uint32_t> next(0);
    auto run = [&] {
        for (;;) {
            const uint32_t b = next.fetch_add(1);
            if (b >= nblocks) {
                return;
            }
            f(b);
        }
    };
    std::vector<std::thread> helpers;
    for (int t = 1; t < nthreads; ++t) {
        helpers.emplace_back(run);
    }
    run();
    for (auto &h : helpers) {
        h.join();
    }
}

inline int host_threads(int requested)
{
    return requested > 0 ? requested : (int)std::max(1u, std::thread::hardware_concurrency());
}

}  // namespace crt
