// Replacement of inc/dg/backend/exblas/fpedot_cuda.cuh: the generalised dot product sum_i f( x_0i, x_1i, ...) behind
// dg::blas1::vdot (blas1.h:90-121) for ANY value type and functor, accumulated in floating-point expansions (FPE) of N
// terms per thread (ExSUM.FPE.hpp:100-116 / accumulate.h:100-118 semantics: a cascade of error-free TwoSum steps; a
// residue that falls off the end raises status 2).  The functor is the user's, so this is a kernel template compiled into
// the caller's translation unit like the reference's fpeDOT<...>; the design differs: expansions live in REGISTERS, warps
// merge them with shuffles, the blocks' expansions meet in global memory and the last block to finish merges them -- ONE
// launch instead of two, no per-thread shared-memory traffic.  As long as no residue is lost the expansion holds the sum
// exactly, so exblas::cpu::Round( fpe) gives the same value as the reference for any summation order.
#pragma once
#include <array>
#include "../dgb_shim.h"

namespace dg
{
namespace exblas
{
///@cond
namespace gpu
{
template<class T>
__device__ __forceinline__ std::enable_if_t<!std::is_integral<T>::value, T> two_sum( T a, T b, T& err)
{
    T s = a + b;
    T z = s - a;
    err = (a - (s - z)) + (b - z);
    return s;
}
template<class T>
__device__ __forceinline__ std::enable_if_t<std::is_integral<T>::value, T> two_sum( T a, T b, T& err)
{
    err = T(0);
    return a + b;
}
// e += x (exactly, unless a residue drops off the end -> lost = true)
template<class T, unsigned N>
__device__ __forceinline__ void fpe_add( T (&e)[N], T x, bool& lost)
{
#pragma unroll
    for( unsigned i = 0; i < N; i++)
    {
        T err;
        e[i] = two_sum( e[i], x, err);
        x = err;
    }
    if( x != T(0)) lost = true;   // also true for NaN / Inf input, as in the reference
}
template<class T>
__device__ __forceinline__ T shfl_down_any( T v, int delta)
{
    // value types wider than a register (complex<double>, ...) travel in 4-byte pieces
    static_assert( sizeof(T) % 4 == 0, "value type must be a multiple of 4 bytes");
    union { T v; int w[sizeof(T) / 4]; } in, out;
    in.v = v;
#pragma unroll
    for( unsigned k = 0; k < sizeof(T) / 4; k++) out.w[k] = __shfl_down_sync( 0xffffffffu, in.w[k], delta);
    return out.v;
}
template<class T, unsigned N>
__device__ __forceinline__ void fpe_warp_merge( T (&e)[N], bool& lost)
{
#pragma unroll
    for( int delta = 16; delta > 0; delta >>= 1)
    {
        T other[N];
#pragma unroll
        for( unsigned k = 0; k < N; k++) other[k] = shfl_down_any( e[k], delta);
#pragma unroll
        for( unsigned k = 0; k < N; k++) fpe_add<T,N>( e, other[k], lost);
    }
}

constexpr unsigned FPE_THREADS = 256;
template<class T, unsigned N, class Functor, class ...PointerOrValues>
__global__ void __launch_bounds__(FPE_THREADS) fpe_dot_kernel( int* status, size_t size, T* partial, unsigned* ticket, T* result, Functor f, PointerOrValues ...xs)
{
    __shared__ T sh[FPE_THREADS / 32][N];
    __shared__ bool is_last;
    T e[N];
#pragma unroll
    for( unsigned k = 0; k < N; k++) e[k] = T(0);
    bool lost = false;
    const size_t stride = (size_t)gridDim.x * FPE_THREADS;
    for( size_t i = (size_t)blockIdx.x * FPE_THREADS + threadIdx.x; i < size; i += stride)
        fpe_add<T,N>( e, (T)f( dgb::shim::elem( xs, i)...), lost);
    fpe_warp_merge<T,N>( e, lost);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if( lane == 0)
    {
#pragma unroll
        for( unsigned k = 0; k < N; k++) sh[warp][k] = e[k];
    }
    if( lost) *status = 2;
    __syncthreads();
    if( warp == 0)
    {
#pragma unroll
        for( unsigned k = 0; k < N; k++) e[k] = lane < FPE_THREADS / 32 ? sh[lane][k] : T(0);
        fpe_warp_merge<T,N>( e, lost);
        if( lane == 0)
        {
#pragma unroll
            for( unsigned k = 0; k < N; k++) partial[(size_t)blockIdx.x * N + k] = e[k];
            if( lost) *status = 2;
            __threadfence();
            is_last = atomicAdd( ticket, 1u) == gridDim.x - 1;
        }
    }
    __syncthreads();
    if( !is_last || warp != 0) return;
    __threadfence();
    // the last block: one warp folds the block expansions (lane-strided), then merges across lanes
    lost = false;
#pragma unroll
    for( unsigned k = 0; k < N; k++) e[k] = T(0);
    for( unsigned b = lane; b < gridDim.x; b += 32)
#pragma unroll
        for( unsigned k = 0; k < N; k++) fpe_add<T,N>( e, partial[(size_t)b * N + k], lost);
    fpe_warp_merge<T,N>( e, lost);
    if( lane == 0)
    {
#pragma unroll
        for( unsigned k = 0; k < N; k++) result[k] = e[k];
        if( lost) *status = 2;
        *ticket = 0;
    }
}
}//namespace gpu
///@endcond

// result on the host (what doDot_fpe_dispatch needs)
template<class T, size_t N, class Functor, class ...PointerOrValues>
inline void fpedot_gpu_host( int* status, unsigned size, T* fpe_host, Functor f, PointerOrValues ...xs_ptr)
{
    static dgb::shim::DeviceScratch<unsigned char> scratch;   // [result N | partial grid*N | status | ticket]
    const unsigned grid = dgb::shim::generic_grid( size, gpu::FPE_THREADS);
    const size_t words = (size_t)(grid + 1) * N * sizeof(T) + 16;
    const bool fresh = scratch.count < words;
    unsigned char* base = scratch.get( words);
    T* result = reinterpret_cast<T*>( base);
    T* partial = result + N;
    int* d_status = reinterpret_cast<int*>( base + (size_t)(grid + 1) * N * sizeof(T));
    unsigned* ticket = reinterpret_cast<unsigned*>( d_status + 1);
    (void)fresh;
    dgb::shim::check( dgb_memset( d_status, 0, 8, nullptr), "dgb_memset");
    dgb::shim::note_generic<Functor>( "vdot");
    gpu::fpe_dot_kernel<T, (unsigned)N, Functor, PointerOrValues...><<<grid, gpu::FPE_THREADS>>>( d_status, (size_t)size, partial, ticket, result, f, xs_ptr...);
    dgb::shim::check_launch( "dg::blas1::vdot");
    dgb::shim::check( dgb_memcpy_d2h( fpe_host, result, N * sizeof(T), nullptr), "dgb_memcpy_d2h");
    dgb::shim::check( dgb_memcpy_d2h( status, d_status, sizeof(int), nullptr), "dgb_memcpy_d2h");
    dgb::shim::check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
}

///@brief GPU version of fpe generalized dot product, result in device memory (signature of fpedot_cuda.cuh:170)
template<class T, size_t N, class Functor, class ...PointerOrValues>
inline void fpedot_gpu( int* status, unsigned size, T* fpe, Functor f, PointerOrValues ...xs_ptr)
{
    T host[N];
    fpedot_gpu_host<T,N,Functor,PointerOrValues...>( status, size, host, f, xs_ptr...);
    dgb::shim::check( dgb_memcpy_h2d( fpe, host, N * sizeof(T), nullptr), "dgb_memcpy_h2d");
    dgb::shim::check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
}

}//namespace exblas
}//namespace dg
