// Shared helpers for libdgb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/dgb200.h"

namespace dgb {

// thread-local last-error message (dgb_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
extern long long g_launches;

#define DGB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t _e = (call);                                              \
        if (_e != cudaSuccess) return dgb::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

// check the launch that just happened and count it
#define DGB_LAUNCHED()                                                        \
    do {                                                                      \
        ++dgb::g_launches;                                                    \
        cudaError_t _e = cudaPeekAtLastError();                               \
        if (_e != cudaSuccess) return dgb::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with the attribute below may be scheduled while the previous
// kernel of the stream still runs -- its blocks become resident as the predecessor's blocks retire and run their prologue --
// and must execute pdl_wait() before it touches anything the predecessor reads or writes (the wait returns when the predecessor
// has completed and its memory is visible).  pdl_trigger() in the predecessor allows that early scheduling; without the
// launch attribute both instructions are no-ops.  Used by the two kernels of the PCG iteration: the finishing block of one
// kernel (exact-dot tail) overlaps the launch and prologue of the next.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef __CUDACC__
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

inline cudaStream_t as_stream(dgb_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (148 on B200); cached
int sm_count();

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// streaming (read-once) 128-bit load / store helpers
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double2 v) { *reinterpret_cast<double2*>(p) = v; }

}  // namespace dgb
