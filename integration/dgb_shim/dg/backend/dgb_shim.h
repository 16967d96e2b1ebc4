// dgb_shim.h -- common part of the Feltor <-> libdgb200.so binding.
//
// The reference resolves every data-parallel operation by overload on an execution-policy tag; the CudaTag overloads live
// in four backend files (blas1_cuda.cuh, sparseblockmat_gpu_kernels.cuh, sparsematrix_gpu.cuh, blas2_stencil.h) plus
// exblas/{exdot,fpedot}_cuda.cuh.  The files next to this one REPLACE those (integration/make_tree.py overlays them on
// a copy of the reference's inc/ tree): the same function names and signatures, bodies that call the C ABI of
// include/dgb200.h.  Operations on library functors / double vectors land in libdgb200.so; user functors (device
// lambdas, which cannot cross a C ABI) run through the small kernel templates below, compiled into the user's
// translation unit exactly like the reference's subroutine_kernel<...> instantiations are.
#pragma once
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <cuda_runtime.h>
#include "dgb200.h"
#include "exceptions.h"

namespace dgb
{
namespace shim
{
// non-zero C-ABI codes become dg::Error, as blas1_cuda.cuh:39-41 does for CUDA errors
inline void check( int code, const char* where)
{
    if( code == 0) return;
    throw dg::Error( dg::Message(_ping_) << where << ": libdgb200 error " << code << ": " << dgb_last_error());
}
inline void check_launch( const char* where)
{
    cudaError_t code = cudaGetLastError();
    if( code != cudaSuccess)
        throw dg::Error( dg::Message(_ping_) << where << ": " << cudaGetErrorString(code));
}
// how many operations went to the library / to a generic kernel template (tests prove the dispatch with these)
struct Counters { long long library = 0, generic = 0; };
inline Counters& counters() { static Counters c; return c; }
inline bool trace_enabled() { static int t = -1; if( t < 0) { const char* e = std::getenv("DGB_SHIM_TRACE"); t = (e && std::atoi(e)) ? 1 : 0; } return t == 1; }
template<class F>
inline void note_generic( const char* what)
{
    counters().generic++;
    if( trace_enabled()) std::fprintf( stderr, "[dgb shim] generic %s <%s>\n", what, typeid(F).name());
}
inline void note_library() { counters().library++; }
// run-time switch of the second binding level (dgb_fused.h): 1 = Elliptic2d / PCG use the fused kernels (default),
// 0 = only the backend dispatch is bound (DGB_SHIM_NO_FUSION=1 or fusion_flag() = 0) -- for A/B tests
inline int& fusion_flag()
{
    static int f = [](){ const char* e = std::getenv( "DGB_SHIM_NO_FUSION"); return (e && std::atoi( e)) ? 0 : 1; }();
    return f;
}

// SparseMatrix::operator* of host matrices goes to the device from this many stored entries (both operands) on; smaller
// products stay on the host (DGB_SHIM_SPGEMM_MIN overrides)
inline size_t spgemm_threshold()
{
    static size_t t = [](){ const char* e = std::getenv( "DGB_SHIM_SPGEMM_MIN"); return e ? (size_t)std::atoll( e) : (size_t)200000; }();
    return t;
}

// persistent-style launch geometry for the generic templates: enough CTAs to fill the machine, never more than needed
inline unsigned generic_grid( size_t size, unsigned threads = 256)
{
    static int sms = 0;
    if( sms == 0) { if( dgb_sm_count( &sms) != 0 || sms <= 0) sms = 148; }
    size_t want = (size + threads - 1) / threads, cap = (size_t)sms * 16;
    if( want == 0) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

// operand classification of the variadic dispatch functions
template<class P> struct is_cptr : std::false_type {};       // readable double vector
template<> struct is_cptr<const double*> : std::true_type {};
template<> struct is_cptr<double*> : std::true_type {};
template<class P> struct is_mptr : std::false_type {};       // writable double vector
template<> struct is_mptr<double*> : std::true_type {};
template<class P> using is_num = std::is_arithmetic<P>;      // scalar operand / functor coefficient (promoted to double, as DG_FMA does)

// Coefficients of the library functors of subroutines.h are private members of trivially copyable structs
// ({T0 m_a; T1 m_b; ...}); a layout-identical mirror reads them.  The asserts break the build, not the result, should a
// functor ever change shape.
template<class... Ts> struct Mirror;
template<class A> struct Mirror<A> { A a; };
template<class A, class B> struct Mirror<A,B> { A a; B b; };
template<class A, class B, class C> struct Mirror<A,B,C> { A a; B b; C c; };
template<class F, class... Ts>
inline Mirror<Ts...> coefficients( const F& f)
{
    static_assert( std::is_trivially_copyable<F>::value, "library functor is not trivially copyable");
    static_assert( sizeof(Mirror<Ts...>) <= sizeof(F), "library functor smaller than its coefficients");
    Mirror<Ts...> m;
    std::memcpy( &m, &f, sizeof(m));
    return m;
}

// element access of the generic templates: scalars broadcast, pointers index
template<class T> __device__ __forceinline__ T elem( T x, size_t) { return x; }
template<class T> __device__ __forceinline__ T& elem( T* x, size_t i) { return x[i]; }

template<class F, class... Ps>
__global__ void __launch_bounds__(256) map_kernel( size_t size, F f, Ps... ps)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride)
        f( elem( ps, i)...);
}
template<class F, class... Ps>
__global__ void __launch_bounds__(256) indexed_kernel( unsigned size, F f, Ps... ps)
{
    const unsigned stride = gridDim.x * blockDim.x;
    for( unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride)
        f( i, ps...);
}

// small owning device buffer (no thrust in the binding)
template<class T>
struct DeviceScratch
{
    T* ptr = nullptr;
    size_t count = 0;
    T* get( size_t n)
    {
        if( n > count)
        {
            if( ptr) dgb_free( ptr);
            void* p = nullptr;
            check( dgb_malloc( &p, n * sizeof(T)), "dgb_malloc");
            ptr = static_cast<T*>(p);
            count = n;
        }
        return ptr;
    }
    ~DeviceScratch() { if( ptr) dgb_free( ptr); }
    DeviceScratch() = default;
    DeviceScratch( const DeviceScratch&) = delete;
    DeviceScratch& operator=( const DeviceScratch&) = delete;
};

// the per-process dot workspace (the reference keeps a static device_vector the same way, blas1_cuda.cuh:33)
inline dgb_dot_ws* dot_workspace()
{
    struct Holder { dgb_dot_ws* ws = nullptr; ~Holder() { if( ws) dgb_dot_ws_destroy( ws); } };
    static Holder h;
    if( !h.ws) check( dgb_dot_ws_create( &h.ws), "dgb_dot_ws_create");
    return h.ws;
}

// Launch plan of an EllSparseBlockMat, owned by the matrix (member added by make_tree.py, mirroring what the reference
// itself does for CSR matrices with detail::CSRCache_gpu in sparsematrix.h:620-628).  Copies start without a plan.
struct EllCache
{
    EllCache() = default;
    EllCache( const EllCache&) {}
    EllCache( EllCache&& src) noexcept { swap( src); }
    EllCache& operator=( const EllCache& src) { if( &src != this) forget(); return *this; }
    EllCache& operator=( EllCache&& src) noexcept { if( &src != this) { forget(); swap( src); } return *this; }
    ~EllCache() { forget(); }
    void swap( EllCache& o) noexcept
    {
        std::swap( plan, o.plan); std::swap( data, o.data); std::swap( cols, o.cols); std::swap( didx, o.didx);
        std::swap( left, o.left); std::swap( right, o.right); std::swap( r0, o.r0); std::swap( r1, o.r1);
    }
    void forget() { if( plan) { dgb_ell_destroy( plan); plan = nullptr; } }
    dgb_ell* plan = nullptr;
    const void *data = nullptr, *cols = nullptr, *didx = nullptr;  // identity of the arrays the plan was built from
    int left = 0, right = 0, r0 = 0, r1 = 0;
};
// cache of a CSR matrix: the sliced-ELL gather plan of the library
struct CsrCache
{
    dgb_gather_plan* plan = nullptr;
    const void *pos = nullptr, *idx = nullptr, *val = nullptr;
    void forget() { if( plan) { dgb_gather_plan_destroy( plan); plan = nullptr; } }
};

}//namespace shim
}//namespace dgb
