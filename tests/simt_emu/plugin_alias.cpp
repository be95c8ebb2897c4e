// plugin_alias.cpp — TEST-ONLY: the headless driver looks up crt_<backend>_read_accum; the emulated plugin is loaded
// under the name "cuda_simt" (libcrt_cuda_simt.so), so give its export that name as well.
#include <cstdint>
struct RenderBackend;
extern "C" int crt_cuda_read_accum(RenderBackend *backend, float *rgb_out);
extern "C" int crt_cuda_simt_read_accum(RenderBackend *backend, float *rgb_out) { return crt_cuda_read_accum(backend, rgb_out); }

extern "C" int crt_cuda_get_stats(RenderBackend *backend, float *stage_ms, int num_stages, uint64_t *counters, int num_counters);
extern "C" int crt_cuda_simt_get_stats(RenderBackend *backend, float *stage_ms, int num_stages, uint64_t *counters, int num_counters)
{
    return crt_cuda_get_stats(backend, stage_ms, num_stages, counters, num_counters);
}
