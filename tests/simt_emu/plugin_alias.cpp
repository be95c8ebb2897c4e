// plugin_alias.cpp — TEST-ONLY: the headless driver looks up crt_<backend>_read_accum; the emulated plugin is loaded
// under the name "cuda_simt" (libcrt_cuda_simt.so), so give its export that name as well.
struct RenderBackend;
extern "C" int crt_cuda_read_accum(RenderBackend *backend, float *rgb_out);
extern "C" int crt_cuda_simt_read_accum(RenderBackend *backend, float *rgb_out) { return crt_cuda_read_accum(backend, rgb_out); }
