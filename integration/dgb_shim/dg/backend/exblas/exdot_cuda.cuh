// Replacement of inc/dg/backend/exblas/exdot_cuda.cuh: the exact (superaccumulator) dot product of the CUDA backend,
// bound to dgb_exdot2 / dgb_exdot3 of libdgb200.so.  Same contract as the reference kernels (exdot_cuda.cuh:319-357): the
// exact sum of the individually rounded products x*y resp. (x*y)*z in a 39 x int64 superaccumulator, status 1 if a product
// was not finite.  The library returns the accumulator NORMALISED, which exblas::cpu::Round and the vector / MPI
// combination code (blas1_dispatch_vector.h:163-172, mpi_accumulate.h:94-125) accept unchanged.
#pragma once
#include <cstdint>
#include "config.h"
#include "../dgb_shim.h"

namespace dg
{
namespace exblas
{
///@cond
namespace detail
{
// operand of a dgb_exdot call: a device pointer to double, or a scalar that multiplies every element
struct DotOperand { const double* ptr; double value; };
inline DotOperand dot_operand( const double* p) { return DotOperand{ p, 0.}; }
inline DotOperand dot_operand( double* p) { return DotOperand{ p, 0.}; }
template<class T>
inline std::enable_if_t<std::is_arithmetic<T>::value, DotOperand> dot_operand( T v) { return DotOperand{ nullptr, (double)v}; }
template<class T>
inline std::enable_if_t<!std::is_same<std::remove_cv_t<T>, double>::value, DotOperand> dot_operand( T*)
{
    throw dg::Error( dg::Message(_ping_) << "libdgb200 computes the exact dot product for double vectors only");
}
struct DotResultBuffer
{
    dgb_dot_result* dev = nullptr;
    dgb_dot_result* get()
    {
        if( !dev) { void* p = nullptr; dgb::shim::check( dgb_malloc( &p, sizeof(dgb_dot_result)), "dgb_malloc"); dev = static_cast<dgb_dot_result*>(p); }
        return dev;
    }
    ~DotResultBuffer() { if( dev) dgb_free( dev); }
};
inline dgb_dot_result* dot_result_buffer() { static DotResultBuffer b; return b.get(); }
inline void finish_host( int64_t* h_superacc, int* status)
{
    dgb_dot_result host;
    dgb::shim::check( dgb_memcpy_d2h( &host, dot_result_buffer(), sizeof(host), nullptr), "dgb_memcpy_d2h");
    dgb::shim::check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
    for( int k = 0; k < BIN_COUNT; k++) h_superacc[k] = host.acc[k];
    *status = host.status;
}
}//namespace detail
///@endcond

// host-result forms used by doDot_dispatch (one device -> host copy of the finished record, no separate status vector)
template<class PointerOrValue1, class PointerOrValue2>
inline void exdot_gpu_host( unsigned size, PointerOrValue1 x1_ptr, PointerOrValue2 x2_ptr, int64_t* h_superacc, int* status)
{
    const detail::DotOperand a = detail::dot_operand( x1_ptr), b = detail::dot_operand( x2_ptr);
    dgb::shim::check( dgb_exdot2( dgb::shim::dot_workspace(), size, a.ptr, a.value, b.ptr, b.value, detail::dot_result_buffer(), nullptr), "dg::blas1::dot");
    dgb::shim::note_library();
    detail::finish_host( h_superacc, status);
}
template<class PointerOrValue1, class PointerOrValue2, class PointerOrValue3>
inline void exdot_gpu_host( unsigned size, PointerOrValue1 x1_ptr, PointerOrValue2 x2_ptr, PointerOrValue3 x3_ptr, int64_t* h_superacc, int* status)
{
    const detail::DotOperand a = detail::dot_operand( x1_ptr), b = detail::dot_operand( x2_ptr), c = detail::dot_operand( x3_ptr);
    dgb::shim::check( dgb_exdot3( dgb::shim::dot_workspace(), size, a.ptr, a.value, b.ptr, b.value, c.ptr, c.value, detail::dot_result_buffer(), nullptr), "dg::blas2::dot");
    dgb::shim::note_library();
    detail::finish_host( h_superacc, status);
}

///@brief GPU version of exact dot product, accumulator left in device memory (signature of exdot_cuda.cuh:319)
template<class PointerOrValue1, class PointerOrValue2>
inline void exdot_gpu( unsigned size, PointerOrValue1 x1_ptr, PointerOrValue2 x2_ptr, int64_t* d_superacc, int* status)
{
    int64_t host[BIN_COUNT];
    exdot_gpu_host( size, x1_ptr, x2_ptr, host, status);
    dgb::shim::check( dgb_memcpy_h2d( d_superacc, host, sizeof(host), nullptr), "dgb_memcpy_h2d");
    dgb::shim::check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
}
///@brief GPU version of exact triple dot product (signature of exdot_cuda.cuh:341)
template<class PointerOrValue1, class PointerOrValue2, class PointerOrValue3>
inline void exdot_gpu( unsigned size, PointerOrValue1 x1_ptr, PointerOrValue2 x2_ptr, PointerOrValue3 x3_ptr, int64_t* d_superacc, int* status)
{
    int64_t host[BIN_COUNT];
    exdot_gpu_host( size, x1_ptr, x2_ptr, x3_ptr, host, status);
    dgb::shim::check( dgb_memcpy_h2d( d_superacc, host, sizeof(host), nullptr), "dgb_memcpy_h2d");
    dgb::shim::check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
}

}//namespace exblas
}//namespace dg
