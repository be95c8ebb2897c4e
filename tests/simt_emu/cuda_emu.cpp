// cuda_emu.cpp — TEST-ONLY: host definitions of the two dozen CUDA runtime entry points crt_cuda_core.cu calls, so
// that the renderer object and its C ABI can run on a machine without a GPU on top of the SIMT environment of
// chameleonrt_b200/csrc/simt_env.h (tests/simt_emu/build.py turns the <<<...>>> launches into simt::launch calls).
// "Device" memory is host memory, streams are synchronous, events are wall-clock timestamps. Eight "devices" (all the same host) with
// two "SMs" each (so that the persistent traversal kernel is launched with more than one block).
#include <chrono>
#include <cstdlib>
#include <cstring>

#include <cuda_runtime.h>

namespace {
struct EmuEvent {
    std::chrono::steady_clock::time_point t;
};
}  // namespace

extern "C" {

cudaError_t cudaGetDeviceCount(int *count)
{
    *count = 8;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int device) { return device >= 0 && device < 8 ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
cudaError_t cudaDeviceGetAttribute(int *value, enum cudaDeviceAttr attr, int)
{
    *value = attr == cudaDevAttrMultiProcessorCount ? 2 : 0;
    return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *numBlocks, const void *, int, size_t, unsigned int)
{
    *numBlocks = 1;
    return cudaSuccess;
}
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *numBlocks, const void *, int, size_t)
{
    *numBlocks = 1;
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int)
{
    *s = reinterpret_cast<cudaStream_t>(new int(0));
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    delete reinterpret_cast<int *>(s);
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaMalloc(void **p, size_t n)
{
    *p = std::malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p)
{
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void **p, size_t n)
{
    *p = std::calloc(1, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned int)
{
    *p = std::calloc(1, n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p)
{
    std::free(p);
    return cudaSuccess;
}
// texture objects (option "hw_textures") have no host emulation: the option fails cleanly here
cudaError_t cudaMallocArray(cudaArray_t *, const cudaChannelFormatDesc *, size_t, size_t, unsigned int) { return cudaErrorNotSupported; }
cudaError_t cudaFreeArray(cudaArray_t) { return cudaSuccess; }
cudaError_t cudaMemcpy2DToArrayAsync(cudaArray_t, size_t, size_t, const void *, size_t, size_t, size_t, enum cudaMemcpyKind, cudaStream_t)
{
    return cudaErrorNotSupported;
}
cudaError_t cudaCreateTextureObject(cudaTextureObject_t *, const cudaResourceDesc *, const cudaTextureDesc *, const cudaResourceViewDesc *)
{
    return cudaErrorNotSupported;
}
cudaError_t cudaDestroyTextureObject(cudaTextureObject_t) { return cudaSuccess; }
cudaChannelFormatDesc cudaCreateChannelDesc(int x, int y, int z, int w, enum cudaChannelFormatKind f)
{
    cudaChannelFormatDesc d;
    d.x = x, d.y = y, d.z = z, d.w = w, d.f = f;
    return d;
}
cudaError_t cudaHostGetDevicePointer(void **dev, void *host, unsigned int)
{
    *dev = host;
    return cudaSuccess;
}
cudaError_t cudaHostRegister(void *, size_t, unsigned int)
{
    return cudaSuccess;
}
cudaError_t cudaHostUnregister(void *)
{
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, enum cudaMemcpyKind, cudaStream_t)
{
    std::memcpy(dst, src, n);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t)
{
    std::memset(dst, v, n);
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e)
{
    *e = reinterpret_cast<cudaEvent_t>(new EmuEvent());
    return cudaSuccess;
}
cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    delete reinterpret_cast<EmuEvent *>(e);
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t)
{
    reinterpret_cast<EmuEvent *>(e)->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(reinterpret_cast<EmuEvent *>(b)->t - reinterpret_cast<EmuEvent *>(a)->t).count();
    return cudaSuccess;
}
// CUDA IPC inside one process: the handle carries the pointer (real CUDA refuses to open a handle in the process
// that exported it; with the emulation all "ranks" of a test live in one process)
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *handle, void *dev_ptr)
{
    std::memset(handle, 0, sizeof(*handle));
    std::memcpy(handle, &dev_ptr, sizeof(dev_ptr));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **dev_ptr, cudaIpcMemHandle_t handle, unsigned int)
{
    std::memcpy(dev_ptr, &handle, sizeof(*dev_ptr));
    return *dev_ptr ? cudaSuccess : cudaErrorInvalidValue;
}
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int)
{
    *can = 1;
    return cudaSuccess;
}
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned int) { return cudaSuccess; }
}
