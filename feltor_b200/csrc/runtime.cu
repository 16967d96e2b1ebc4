// Runtime part of the C ABI: device memory, copies, error reporting.
#include "common.cuh"
#include <cstring>

namespace dgb {
static thread_local char g_err[512] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    set_error("CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    cudaGetLastError();  // clear sticky-less errors
    return (int)e;
}
int sm_count() {
    static int dev_cached = -1, sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != dev_cached) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        dev_cached = dev;
    }
    return sms > 0 ? sms : 148;
}
}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_version(void) { return 100; }
const char* dgb_last_error(void) { return g_err; }
int dgb_device_count(int* count) { DGB_CUDA(cudaGetDeviceCount(count)); return 0; }
int dgb_set_device(int device) { DGB_CUDA(cudaSetDevice(device)); return 0; }
int dgb_sm_count(int* count) {
    int dev;
    DGB_CUDA(cudaGetDevice(&dev));
    DGB_CUDA(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev));
    return 0;
}
int dgb_malloc(void** ptr, size_t bytes) { DGB_CUDA(cudaMalloc(ptr, bytes ? bytes : 1)); return 0; }
int dgb_free(void* ptr) { DGB_CUDA(cudaFree(ptr)); return 0; }
int dgb_malloc_host(void** ptr, size_t bytes) { DGB_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1)); return 0; }
int dgb_free_host(void* ptr) { DGB_CUDA(cudaFreeHost(ptr)); return 0; }
int dgb_memcpy_h2d(void* dst, const void* src, size_t bytes, dgb_stream_t s) {
    DGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(s)));
    return 0;
}
int dgb_memcpy_d2h(void* dst, const void* src, size_t bytes, dgb_stream_t s) {
    DGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(s)));
    return 0;
}
int dgb_memcpy_d2d(void* dst, const void* src, size_t bytes, dgb_stream_t s) {
    DGB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(s)));
    return 0;
}
int dgb_memset(void* dst, int value, size_t bytes, dgb_stream_t s) {
    DGB_CUDA(cudaMemsetAsync(dst, value, bytes, as_stream(s)));
    return 0;
}
int dgb_stream_synchronize(dgb_stream_t s) { DGB_CUDA(cudaStreamSynchronize(as_stream(s))); return 0; }
int dgb_device_synchronize(void) { DGB_CUDA(cudaDeviceSynchronize()); return 0; }
long long dgb_launch_count(void) { return g_launches; }
}
