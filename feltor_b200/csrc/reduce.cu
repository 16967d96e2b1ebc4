// dg::blas1::reduce (inc/dg/blas1.h:213-223 -> doReduce_dispatch, backend/blas1_cuda.cuh:96-102: thrust::transform_reduce)
// for the closed set of (binary op, unary op) pairs the library and the applications use: plain sum (fast_l2norm,
// adaptive.h:31-35), maximum / minimum (toefl.cpp:141, feltor/init.h:422), logical-or of ISNAN / ISNOTFINITE
// (functors.h:185-220).  Maximum, minimum and logical-or are order independent, so the result equals the reference's for
// any launch geometry; the floating-point SUM is -- as in the reference -- not reproducible (use dgb_dot2 with a scalar
// operand for the exact sum: blas1::vdot / dot(1., x) map to the superaccumulator kernels).
#include "common.cuh"
#include <cmath>

namespace dgb {

__device__ __forceinline__ double red_unary(int u, double x) {
    switch (u) {
        case DGB_UNARY_ABS: return fabs(x);
        case DGB_UNARY_SQUARE: return __dmul_rn(x, x);
        case DGB_UNARY_ISNAN: return isnan(x) ? 1. : 0.;
        case DGB_UNARY_ISNOTFINITE: return isfinite(x) ? 0. : 1.;
        default: return x;
    }
}
__device__ __forceinline__ double red_binary(int op, double a, double b) {
    switch (op) {
        case DGB_REDUCE_MAX: return (a < b) ? b : a;  // thrust::maximum
        case DGB_REDUCE_MIN: return (b < a) ? b : a;  // thrust::minimum
        case DGB_REDUCE_OR: return (a != 0. || b != 0.) ? 1. : 0.;
        default: return __dadd_rn(a, b);
    }
}

__global__ void __launch_bounds__(256) reduce_kernel(size_t n, const double* __restrict__ x, int op, int unary, double init,
                                                     double* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                     double* __restrict__ result) {
    __shared__ double sh[8];
    __shared__ int s_last;
    // every thread starts from the neutral element of its op: the user's `init` joins once at the end
    const double neutral = op == DGB_REDUCE_MAX ? -INFINITY : (op == DGB_REDUCE_MIN ? INFINITY : 0.);
    double acc = neutral;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc = red_binary(op, acc, red_unary(unary, x[i]));
    for (int off = 16; off > 0; off >>= 1) acc = red_binary(op, acc, __shfl_down_sync(0xffffffffu, acc, off));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) acc = red_binary(op, acc, sh[w]);
        partial[blockIdx.x] = acc;
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    acc = neutral;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) acc = red_binary(op, acc, __ldcg(partial + b));
    for (int off = 16; off > 0; off >>= 1) acc = red_binary(op, acc, __shfl_down_sync(0xffffffffu, acc, off));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) acc = red_binary(op, acc, sh[w]);
        *result = red_binary(op, init, acc);
        *ticket = 0;
    }
}

}  // namespace dgb

using namespace dgb;

extern "C" int dgb_reduce(size_t n, const double* x, int op, int unary, double init, double* result_host, dgb_stream_t s) {
    if (op < 0 || op > DGB_REDUCE_OR || unary < 0 || unary > DGB_UNARY_ISNOTFINITE) { set_error("dgb_reduce: unknown op %d / unary %d", op, unary); return DGB_ERR_INVALID; }
    if (!result_host) { set_error("dgb_reduce: result pointer is NULL"); return DGB_ERR_INVALID; }
    if (n == 0) { *result_host = init; return 0; }
    // scratch: not thread-safe / re-entrant, like the reference's static buffers (blas1_cuda.cuh:18,33,50)
    static double* scratch = nullptr;      // [1024 partials][1 result]
    static unsigned int* ticket = nullptr;
    static double* pinned = nullptr;
    cudaStream_t st = as_stream(s);
    if (!scratch) {
        DGB_CUDA(cudaMalloc(&scratch, 1025 * sizeof(double)));
        DGB_CUDA(cudaMalloc(&ticket, sizeof(unsigned int)));
        DGB_CUDA(cudaMemset(ticket, 0, sizeof(unsigned int)));
        DGB_CUDA(cudaMallocHost(&pinned, sizeof(double)));
    }
    size_t want = (n + 256 * 8 - 1) / (256 * 8), cap = std::min<size_t>((size_t)sm_count() * 4, 1024);
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    reduce_kernel<<<grid, 256, 0, st>>>(n, x, op, unary, init, scratch, ticket, scratch + 1024);
    DGB_LAUNCHED();
    DGB_CUDA(cudaMemcpyAsync(pinned, scratch + 1024, sizeof(double), cudaMemcpyDeviceToHost, st));
    DGB_CUDA(cudaStreamSynchronize(st));
    *result_host = *pinned;
    return 0;
}
