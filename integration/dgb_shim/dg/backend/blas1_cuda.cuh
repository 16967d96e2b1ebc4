// Replacement of inc/dg/backend/blas1_cuda.cuh: the dg::CudaTag overloads of the blas1 dispatch, bound to libdgb200.so.
//
//   doSubroutine_dispatch( CudaTag, size, functor, operands...)        blas1_cuda.cuh:88-94
//   doDot_dispatch( CudaTag, status, size, x, y [, z])                 blas1_cuda.cuh:30-64
//   doDot_fpe_dispatch( CudaTag, status, size, fpe, functor, xs...)    blas1_cuda.cuh:13-27
//   doReduce_dispatch( CudaTag, size, x, init, op, unary_op)           blas1_cuda.cuh:96-102
//   doKronecker_dispatch( CudaTag, y, size, binary, f, sizes, xs...)   blas1_cuda.cuh:133-146
//
// A call whose functor is one of the library's (inc/dg/subroutines.h, dg::blas1::axpby / pointwiseDot / ... all funnel
// through them) and whose vector operands are double goes to the matching dgb_* entry point (128-bit vectorised kernels,
// the functor's own rounding sequence); everything else -- user functors, device lambdas, other value types -- runs
// through the generic templates of dgb_shim.h.
#ifndef _DG_BLAS_CUDA_
#define _DG_BLAS_CUDA_
#include <array>
#include <vector>
#include <utility>
#include "dgb_shim.h"
#include "execution_policy.h"
#include "dg/subroutines.h"
#include "exblas/exdot_cuda.cuh"
#include "exblas/fpedot_cuda.cuh"
namespace dg
{
// functors.h / topology/multiply.h come after this file in the include order of dg/algorithm.h (and include blas1.h
// themselves): the routing table only needs their names
template<class T> struct EXP;
template<class T> struct LN;
template<class T> struct SQRT;
template<class T> struct INVERT;
template<class T> struct ABS;
template<class T> struct InvSqrt;
struct Square;
struct UpwindProduct;
struct TensorMultiply2d;
}//namespace dg
namespace dgb
{
namespace shim
{
// ---------------------------------------------------------------------------------------------------------------------
// Route<Functor, Operands...>: value = true if libdgb200 has an entry point for this call, call() makes it.
// ---------------------------------------------------------------------------------------------------------------------
template<class F, class Enable, class... Ps> struct RouteImpl : std::false_type {};
template<class F, class... Ps> using Route = RouteImpl<F, void, Ps...>;
#define DGB_ROUTE_IF(...) std::enable_if_t<(__VA_ARGS__)>

// dg::blas1::copy: subroutine( equals(), source, target)                                               blas1.h:246
template<class X, class Y>
struct RouteImpl<dg::equals, DGB_ROUTE_IF( is_cptr<X>::value && is_mptr<Y>::value), X, Y> : std::true_type
{
    static int call( size_t n, const dg::equals&, const double* x, double* y) { return x == y ? 0 : dgb_copy( n, x, y, nullptr); }
};
template<class X, class Y>
struct RouteImpl<dg::equals, DGB_ROUTE_IF( is_num<X>::value && is_mptr<Y>::value), X, Y> : std::true_type
{
    static int call( size_t n, const dg::equals&, X x, double* y) { return dgb_fill( n, (double)x, y, nullptr); }
};
// scal, plus                                                                                         blas1.h:267,288
template<class T, class Y>
struct RouteImpl<dg::Scal<T>, DGB_ROUTE_IF( is_num<T>::value && is_mptr<Y>::value), Y> : std::true_type
{
    static int call( size_t n, const dg::Scal<T>& f, double* y) { return dgb_scal( n, y, (double)coefficients<dg::Scal<T>, T>( f).a, nullptr); }
};
template<class T, class Y>
struct RouteImpl<dg::Plus<T>, DGB_ROUTE_IF( is_num<T>::value && is_mptr<Y>::value), Y> : std::true_type
{
    static int call( size_t n, const dg::Plus<T>& f, double* y) { return dgb_plus( n, y, (double)coefficients<dg::Plus<T>, T>( f).a, nullptr); }
};
// axpby                                                                                                 blas1.h:316
template<class A, class B, class X, class Y>
struct RouteImpl<dg::Axpby<A,B>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_cptr<X>::value && is_mptr<Y>::value), X, Y> : std::true_type
{
    static int call( size_t n, const dg::Axpby<A,B>& f, const double* x, double* y)
    {
        auto c = coefficients<dg::Axpby<A,B>, A, B>( f);
        return dgb_axpby( n, (double)c.a, x, (double)c.b, y, nullptr);
    }
};
// y = a x y + b y (pointwiseDot with one factor aliasing the result)                                 blas1.h:414,419
template<class A, class B, class X, class Y>
struct RouteImpl<dg::AxyPby<A,B>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_cptr<X>::value && is_mptr<Y>::value), X, Y> : std::true_type
{
    static int call( size_t n, const dg::AxyPby<A,B>& f, const double* x, double* y)
    {
        auto c = coefficients<dg::AxyPby<A,B>, A, B>( f);
        return dgb_pointwise_dot( n, (double)c.a, y, x, (double)c.b, y, nullptr);  // x1 == y selects the AxyPby sequence
    }
};
// axpbypgz                                                                                              blas1.h:361
template<class A, class B, class G, class X, class Y, class Z>
struct RouteImpl<dg::Axpbypgz<A,B,G>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_num<G>::value && is_cptr<X>::value && is_cptr<Y>::value && is_mptr<Z>::value), X, Y, Z> : std::true_type
{
    static int call( size_t n, const dg::Axpbypgz<A,B,G>& f, const double* x, const double* y, double* z)
    {
        auto c = coefficients<dg::Axpbypgz<A,B,G>, A, B, G>( f);
        return dgb_axpbypgz( n, (double)c.a, x, (double)c.b, y, (double)c.c, z, nullptr);
    }
};
// pointwiseDot with two and three factors                                                          blas1.h:423,472
template<class A, class B, class X1, class X2, class Y>
struct RouteImpl<dg::PointwiseDot<A,B>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_cptr<X1>::value && is_cptr<X2>::value && is_mptr<Y>::value), X1, X2, Y> : std::true_type
{
    static int call( size_t n, const dg::PointwiseDot<A,B>& f, const double* x1, const double* x2, double* y)
    {
        if( x1 == y || x2 == y) return -1000;  // the library would switch to the AxyPby sequence: keep the functor's own
        auto c = coefficients<dg::PointwiseDot<A,B>, A, B>( f);
        return dgb_pointwise_dot( n, (double)c.a, x1, x2, (double)c.b, y, nullptr);
    }
};
template<class A, class B, class X1, class X2, class X3, class Y>
struct RouteImpl<dg::PointwiseDot<A,B>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_cptr<X1>::value && is_cptr<X2>::value && is_cptr<X3>::value && is_mptr<Y>::value), X1, X2, X3, Y> : std::true_type
{
    static int call( size_t n, const dg::PointwiseDot<A,B>& f, const double* x1, const double* x2, const double* x3, double* y)
    {
        auto c = coefficients<dg::PointwiseDot<A,B>, A, B>( f);
        return dgb_pointwise_dot3( n, (double)c.a, x1, x2, x3, (double)c.b, y, nullptr);
    }
};
// z = a x1 y1 + b x2 y2 + g z                                                                           blas1.h:565
template<class A, class B, class G, class X1, class Y1, class X2, class Y2, class Z>
struct RouteImpl<dg::PointwiseDot2<A,B,G>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_num<G>::value && is_cptr<X1>::value && is_cptr<Y1>::value && is_cptr<X2>::value && is_cptr<Y2>::value && is_mptr<Z>::value), X1, Y1, X2, Y2, Z> : std::true_type
{
    static int call( size_t n, const dg::PointwiseDot2<A,B,G>& f, const double* x1, const double* y1, const double* x2, const double* y2, double* z)
    {
        auto c = coefficients<dg::PointwiseDot2<A,B,G>, A, B, G>( f);
        return dgb_pointwise_dot2( n, (double)c.a, x1, y1, (double)c.b, x2, y2, (double)c.c, z, nullptr);
    }
};
// pointwiseDivide, three- and two-operand form                                                      blas1.h:502,506
template<class A, class B, class X1, class X2, class Y>
struct RouteImpl<dg::PointwiseDivide<A,B>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_cptr<X1>::value && is_cptr<X2>::value && is_mptr<Y>::value), X1, X2, Y> : std::true_type
{
    static int call( size_t n, const dg::PointwiseDivide<A,B>& f, const double* x1, const double* x2, double* y)
    {
        if( x1 == y) return -1000;
        auto c = coefficients<dg::PointwiseDivide<A,B>, A, B>( f);
        return dgb_pointwise_divide( n, (double)c.a, x1, x2, (double)c.b, y, nullptr);
    }
};
template<class A, class B, class X, class Y>
struct RouteImpl<dg::PointwiseDivide<A,B>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_cptr<X>::value && is_mptr<Y>::value), X, Y> : std::true_type
{
    static int call( size_t n, const dg::PointwiseDivide<A,B>& f, const double* x, double* y)
    {
        auto c = coefficients<dg::PointwiseDivide<A,B>, A, B>( f);
        return dgb_pointwise_divide( n, (double)c.a, y, x, (double)c.b, y, nullptr);  // x1 == y selects the two-operand sequence
    }
};
// evaluate( z, equals(), PairSum(), alpha, x, beta, y) = three-vector axpby                            blas1.h:384
template<class Z, class A, class X, class B, class Y>
struct RouteImpl<dg::Evaluate<dg::equals, dg::PairSum>, DGB_ROUTE_IF( is_mptr<Z>::value && is_num<A>::value && is_cptr<X>::value && is_num<B>::value && is_cptr<Y>::value), Z, A, X, B, Y> : std::true_type
{
    static int call( size_t n, const dg::Evaluate<dg::equals, dg::PairSum>&, double* z, A a, const double* x, B b, const double* y)
    {
        return dgb_axpbyz( n, (double)a, x, (double)b, y, z, nullptr);
    }
};
// evaluate( y, equals(), PairSum(), x1, x2) = pointwiseDot( x1, x2, y)                                 blas1.h:443
template<class Y, class X1, class X2>
struct RouteImpl<dg::Evaluate<dg::equals, dg::PairSum>, DGB_ROUTE_IF( is_mptr<Y>::value && is_cptr<X1>::value && is_cptr<X2>::value), Y, X1, X2> : std::true_type
{
    static int call( size_t n, const dg::Evaluate<dg::equals, dg::PairSum>&, double* y, const double* x1, const double* x2)
    {
        return dgb_pointwise_dot_xy( n, x1, x2, y, nullptr);
    }
};
// evaluate( y, equals(), divides(), x1, x2) = pointwiseDivide( x1, x2, y)                              blas1.h:527
template<class Y, class X1, class X2>
struct RouteImpl<dg::Evaluate<dg::equals, dg::divides>, DGB_ROUTE_IF( is_mptr<Y>::value && is_cptr<X1>::value && is_cptr<X2>::value), Y, X1, X2> : std::true_type
{
    static int call( size_t n, const dg::Evaluate<dg::equals, dg::divides>&, double* y, const double* x1, const double* x2)
    {
        return dgb_pointwise_divide_xy( n, x1, x2, y, nullptr);
    }
};
// evaluate( y, Axpby( alpha, beta), UpwindProduct(), v, backward, forward): dg::Advection::upwind  advection.h:112-120
template<class A, class B, class Y, class V, class Bk, class Fw>
struct RouteImpl<dg::Evaluate<dg::Axpby<A,B>, dg::UpwindProduct>, DGB_ROUTE_IF( is_num<A>::value && is_num<B>::value && is_mptr<Y>::value && is_cptr<V>::value && is_cptr<Bk>::value && is_cptr<Fw>::value), Y, V, Bk, Fw> : std::true_type
{
    static int call( size_t n, const dg::Evaluate<dg::Axpby<A,B>, dg::UpwindProduct>& f, double* y, const double* v, const double* back, const double* forw)
    {
        auto c = coefficients<dg::Evaluate<dg::Axpby<A,B>, dg::UpwindProduct>, A, B>( f);  // m_f = Axpby comes first
        return dgb_upwind_axpby( n, (double)c.a, v, back, forw, (double)c.b, y, nullptr);
    }
};
// transform( x, y, op) = subroutine( Evaluate<equals, op>, y, x) for the unary functors of functors.h   blas1.h:587
template<class Op> struct unary_code { static constexpr int value = -1; };
template<> struct unary_code<dg::EXP<double>> { static constexpr int value = DGB_OP_EXP; };
template<> struct unary_code<dg::LN<double>> { static constexpr int value = DGB_OP_LN; };
template<> struct unary_code<dg::SQRT<double>> { static constexpr int value = DGB_OP_SQRT; };
template<> struct unary_code<dg::INVERT<double>> { static constexpr int value = DGB_OP_INVERT; };
template<> struct unary_code<dg::ABS<double>> { static constexpr int value = DGB_OP_ABS; };
template<> struct unary_code<dg::Square> { static constexpr int value = DGB_OP_SQUARE; };
template<> struct unary_code<dg::InvSqrt<double>> { static constexpr int value = DGB_OP_INVSQRT; };
template<class Op, class Y, class X>
struct RouteImpl<dg::Evaluate<dg::equals, Op>, DGB_ROUTE_IF( unary_code<Op>::value >= 0 && is_mptr<Y>::value && is_cptr<X>::value), Y, X> : std::true_type
{
    static int call( size_t n, const dg::Evaluate<dg::equals, Op>&, double* y, const double* x)
    {
        return dgb_transform( n, unary_code<Op>::value, x, y, nullptr);
    }
};
// TensorMultiply2d with a scalar or a vector prefactor and a full tensor                        topology/multiply.h:217
template<class L, class T00, class T01, class T10, class T11, class I0, class I1, class M, class O0, class O1>
struct RouteImpl<dg::TensorMultiply2d, DGB_ROUTE_IF( (is_num<L>::value || is_cptr<L>::value) && is_cptr<T00>::value && is_cptr<T01>::value && is_cptr<T10>::value && is_cptr<T11>::value
        && is_cptr<I0>::value && is_cptr<I1>::value && is_num<M>::value && is_mptr<O0>::value && is_mptr<O1>::value), L, T00, T01, T10, T11, I0, I1, M, O0, O1> : std::true_type
{
    static const double* lam_ptr( L lambda) { if constexpr( is_cptr<L>::value) return lambda; else return nullptr; }
    static double lam_val( L lambda) { if constexpr( is_cptr<L>::value) return 0.; else return (double)lambda; }
    static int call( size_t n, const dg::TensorMultiply2d&, L lambda, const double* t00, const double* t01, const double* t10, const double* t11,
                     const double* in0, const double* in1, M mu, double* out0, double* out1)
    {
        return dgb_tensor_multiply2d( n, lam_ptr( lambda), lam_val( lambda), t00, t01, t10, t11, in0, in1, (double)mu, out0, out1, nullptr);
    }
};
#undef DGB_ROUTE_IF

// ---------------------------------------------------------------------------------------------------------------------
// generic reduction (any value type / functors): one launch, block partials + last-block finish
// ---------------------------------------------------------------------------------------------------------------------
template<class T, class Pointer, class BinaryOp, class UnaryOp>
__global__ void __launch_bounds__(256) reduce_kernel( size_t size, Pointer x, BinaryOp op, UnaryOp unary, T* partial, unsigned* ticket, T* result)
{
    __shared__ T sh[256];
    __shared__ bool last;
    const unsigned tid = threadIdx.x;
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + tid;
    // the launch guarantees gridDim.x * 256 <= size rounded up to one block, so only the very last threads can be empty
    bool have = i < size;
    T acc = have ? (T)unary( x[i]) : T();
    for( i += stride; i < size; i += stride) acc = op( acc, (T)unary( x[i]));
    // tree over the block; empty threads are skipped through the element count
    const size_t first = (size_t)blockIdx.x * 256;
    const unsigned valid = (unsigned)( size - first < 256 ? size - first : 256);  // threads of this block that own data
    sh[tid] = acc;
    __syncthreads();
    for( unsigned s = 128; s > 0; s >>= 1)
    {
        if( tid < s && tid + s < valid) sh[tid] = op( sh[tid], sh[tid + s]);
        __syncthreads();
    }
    if( tid == 0)
    {
        partial[blockIdx.x] = sh[0];
        __threadfence();
        last = atomicAdd( ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if( !last) return;
    __threadfence();
    const unsigned nb = gridDim.x;
    have = tid < nb;
    if( have) { acc = partial[tid]; for( unsigned b = tid + 256; b < nb; b += 256) acc = op( acc, partial[b]); }
    sh[tid] = acc;
    __syncthreads();
    const unsigned validb = nb < 256 ? nb : 256;
    for( unsigned s = 128; s > 0; s >>= 1)
    {
        if( tid < s && tid + s < validb) sh[tid] = op( sh[tid], sh[tid + s]);
        __syncthreads();
    }
    if( tid == 0) { *result = sh[0]; *ticket = 0; }
}

// ---------------------------------------------------------------------------------------------------------------------
// generic Kronecker evaluation: y[i] = binary( f( x0[i0], x1[i1], ...), y[i]) with i = i0 + n0 (i1 + n1 (...))
// ---------------------------------------------------------------------------------------------------------------------
template<size_t N> struct KronSizes { size_t n[N]; };
template<class Binary, class F, class Pointer, size_t N, size_t... I, class... Ps>
__device__ __forceinline__ void kron_apply( Binary& binary, F& f, Pointer y, size_t i, const size_t (&idx)[N], std::index_sequence<I...>, Ps... xs)
{
    binary( f( elem( xs, idx[I])...), y[i]);
}
template<class Binary, class F, size_t N, class Pointer, class... Ps>
__global__ void __launch_bounds__(256) kron_kernel( size_t size, KronSizes<N> sizes, Pointer y, Binary binary, F f, Ps... xs)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += stride)
    {
        size_t idx[N], rest = i;
#pragma unroll
        for( size_t k = 0; k < N; k++) { idx[k] = rest % sizes.n[k]; rest /= sizes.n[k]; }
        kron_apply( binary, f, y, i, idx, std::make_index_sequence<N>(), xs...);
    }
}
}//namespace shim
}//namespace dgb

namespace dg
{
namespace blas1
{
namespace detail
{
template< class Subroutine, class PointerOrValue, class ...PointerOrValues>
inline void doSubroutine_dispatch( CudaTag, int size, Subroutine f, PointerOrValue x, PointerOrValues... xs)
{
    if( size <= 0) return;
    using R = dgb::shim::Route<Subroutine, PointerOrValue, PointerOrValues...>;
    if constexpr( R::value)
    {
        const int code = R::call( (size_t)size, f, x, xs...);
        if( code != -1000)  // -1000: aliasing pattern the entry point treats differently from this functor
        {
            dgb::shim::check( code, "dg::blas1::subroutine");
            dgb::shim::note_library();
            return;
        }
    }
    dgb::shim::note_generic<Subroutine>( "subroutine");
    dgb::shim::map_kernel<Subroutine, PointerOrValue, PointerOrValues...><<<dgb::shim::generic_grid( size), 256>>>( (size_t)size, f, x, xs...);
    dgb::shim::check_launch( "dg::blas1::subroutine");
}

// exact dot: the normalised superaccumulator on the host, as the callers (blas1.h:159-166, blas2 dispatch) expect
template<class PointerOrValue1, class PointerOrValue2>
inline std::vector<int64_t> doDot_dispatch( CudaTag, int* status, unsigned size, PointerOrValue1 x_ptr, PointerOrValue2 y_ptr)
{
    std::vector<int64_t> h_superacc( exblas::BIN_COUNT);
    exblas::exdot_gpu_host( size, x_ptr, y_ptr, h_superacc.data(), status);
    return h_superacc;
}
template<class PointerOrValue1, class PointerOrValue2, class PointerOrValue3>
inline std::vector<int64_t> doDot_dispatch( CudaTag, int* status, unsigned size, PointerOrValue1 x_ptr, PointerOrValue2 y_ptr, PointerOrValue3 z_ptr)
{
    std::vector<int64_t> h_superacc( exblas::BIN_COUNT);
    exblas::exdot_gpu_host( size, x_ptr, y_ptr, z_ptr, h_superacc.data(), status);
    return h_superacc;
}

template<class T, size_t N, class Functor, class ...PointerOrValues>
inline void doDot_fpe_dispatch( CudaTag, int * status, unsigned size, std::array<T,N>& fpe, Functor f, PointerOrValues ...xs_ptr)
{
    exblas::fpedot_gpu_host<T,N,Functor,PointerOrValues...>( status, size, fpe.data(), f, xs_ptr...);
}

template<class T, class Pointer, class BinaryOp, class UnaryOp>
inline T doReduce_dispatch( CudaTag, int size, Pointer x, T init, BinaryOp op, UnaryOp unary_op)
{
    if( size <= 0) return init;
    static dgb::shim::DeviceScratch<unsigned char> scratch;
    static dgb::shim::DeviceScratch<unsigned> ticket;
    unsigned* t = ticket.count ? ticket.ptr : nullptr;
    if( !t) { t = ticket.get( 1); dgb::shim::check( dgb_memset( t, 0, sizeof(unsigned), nullptr), "dgb_memset"); }
    unsigned grid = dgb::shim::generic_grid( (size_t)size);
    if( (size_t)grid * 256 > (size_t)size + 255) grid = (unsigned)(( (size_t)size + 255) / 256);
    const size_t bytes = ((size_t)grid + 1) * sizeof(T) + 64;
    T* partial = reinterpret_cast<T*>( scratch.get( bytes));
    dgb::shim::note_generic<BinaryOp>( "reduce");
    dgb::shim::reduce_kernel<T, Pointer, BinaryOp, UnaryOp><<<grid, 256>>>( (size_t)size, x, op, unary_op, partial + 1, t, partial);
    dgb::shim::check_launch( "dg::blas1::reduce");
    T result;
    dgb::shim::check( dgb_memcpy_d2h( &result, partial, sizeof(T), nullptr), "dgb_memcpy_d2h");
    dgb::shim::check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
    return op( init, result);
}

template<class Binary, class F, size_t N, class Pointer, class ...PointerOrValues>
inline void doKronecker_dispatch( dg::CudaTag, Pointer y, size_t size, Binary && binary, F && f, const std::array<size_t, N>& sizes, PointerOrValues ...xs)
{
    if( size == 0) return;
    dgb::shim::KronSizes<N> s;
    for( size_t k = 0; k < N; k++) s.n[k] = sizes[k];
    dgb::shim::note_generic<std::decay_t<F>>( "kronecker");
    dgb::shim::kron_kernel<std::decay_t<Binary>, std::decay_t<F>, N, Pointer, PointerOrValues...><<<dgb::shim::generic_grid( size), 256>>>(
        size, s, y, binary, f, xs...);
    dgb::shim::check_launch( "dg::blas1::kronecker");
}

}//namespace detail
}//namespace blas1
}//namespace dg
#endif //_DG_BLAS_CUDA_
