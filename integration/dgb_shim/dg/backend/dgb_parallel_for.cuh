// dg::CudaTag overload of doParallelFor_dispatch (inc/dg/backend/blas2_stencil.h:20-36), spliced into blas2_stencil.h by
// integration/make_tree.py in place of the reference's CUDA section.  dg::blas2::parallel_for / dg::blas2::stencil call
//     f( i, x, xs...)   for i in [0, size)
// with a user functor.  The CSR stencil functors of the library (inc/dg/topology/filter.h:174-336: CSRMedianFilter,
// CSRSWMFilter, CSRAverageFilter, CSRSymvFilter, CSRSlopeLimiter applied through blas2::stencil( f, M, x, y)) go to dgb_csr_stencil;
// every other functor runs through the indexed kernel template of dgb_shim.h.
#pragma once
#include "dgb_shim.h"
#include "execution_policy.h"

namespace dg
{
struct CSRMedianFilter;
struct CSRAverageFilter;
struct CSRSymvFilter;
template<class T> struct CSRSWMFilter;
template<class T> struct CSRSlopeLimiter;
}//namespace dg
namespace dgb
{
namespace shim
{
template<class F> struct stencil_code { static constexpr int value = -1; };
template<> struct stencil_code<dg::CSRMedianFilter> { static constexpr int value = DGB_STENCIL_MEDIAN; };
template<> struct stencil_code<dg::CSRAverageFilter> { static constexpr int value = DGB_STENCIL_AVERAGE; };
template<> struct stencil_code<dg::CSRSymvFilter> { static constexpr int value = DGB_STENCIL_SYMV; };
template<> struct stencil_code<dg::CSRSWMFilter<double>> { static constexpr int value = DGB_STENCIL_SWM; };
template<> struct stencil_code<dg::CSRSlopeLimiter<double>> { static constexpr int value = DGB_STENCIL_SLOPE; };
template<class F> inline double stencil_alpha( const F&) { return 0.; }
inline double stencil_alpha( const dg::CSRSWMFilter<double>& f) { return coefficients<dg::CSRSWMFilter<double>, double>( f).a; }
inline double stencil_alpha( const dg::CSRSlopeLimiter<double>& f) { return coefficients<dg::CSRSlopeLimiter<double>, double>( f).a; }
// the argument pack blas2::stencil( f, SparseMatrix, x, y) produces: row offsets, columns, values, x, y
template<class... Ps> struct is_csr_pack : std::false_type {};
template<> struct is_csr_pack<const int*, const int*, const double*, const double*, double*> : std::true_type {};
}//namespace shim
}//namespace dgb
namespace dg
{
namespace blas2
{
namespace detail
{
template< class Stencil, class PointerOrValue, class ...PointerOrValues>
inline void doParallelFor_dispatch( CudaTag, unsigned size, Stencil f, PointerOrValue x, PointerOrValues... xs)
{
    if( size == 0) return;
    if constexpr( dgb::shim::stencil_code<Stencil>::value >= 0 && dgb::shim::is_csr_pack<PointerOrValue, PointerOrValues...>::value)
    {
        auto call = [&]( const int* pos, const int* idx, const double* val, const double* in, double* out)
        {
            return dgb_csr_stencil( dgb::shim::stencil_code<Stencil>::value, (int)size, pos, idx, val, dgb::shim::stencil_alpha( f), in, out, nullptr);
        };
        dgb::shim::check( call( x, xs...), "dg::blas2::stencil");
        dgb::shim::note_library();
    }
    else
    {
        dgb::shim::note_generic<Stencil>( "parallel_for");
        dgb::shim::indexed_kernel<Stencil, PointerOrValue, PointerOrValues...><<<dgb::shim::generic_grid( size), 256>>>( size, f, x, xs...);
        dgb::shim::check_launch( "dg::blas2::parallel_for");
    }
}
}//namespace detail
}//namespace blas2
}//namespace dg
