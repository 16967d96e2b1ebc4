// dgb_fused.h -- second level of the Feltor <-> libdgb200.so binding: the FUSED kernels behind the reference's own classes.
//
// With the backend files of this directory alone, dg::Elliptic2d::symv is still the reference's composition (six block-matrix
// products and two element-wise passes, now running in libdgb200) and dg::PCG::solve its loop of blas1 / blas2 calls with a
// host round trip per dot.  integration/make_tree.py therefore also splices three small hooks into elliptic.h, helmholtz.h
// and pcg.h (the classes keep their interface and every other code path):
//   dg::Elliptic2d<Geometry, DMatrix, DVec>::symv( alpha, x, beta, y)  ->  dgb_elliptic2d_symv   (ONE kernel, 24 B/dof)
//   dg::PCG<DVec>::solve( A, x, b, P, W, eps, ...) with A an Elliptic2d or a Helmholtz of one, P and W vectors
//                                                                    ->  dgb_pcg_solve_elliptic2d (3 kernels / iteration)
//   dg::Advection<Geometry, DMatrix, DVec>::upwind( alpha, vx, vy, f, beta, result)   ->  dgb_advection_upwind  (ONE kernel)
//   dg::ArakawaX<Geometry, DMatrix, DVec>::operator()( alpha, lhs, rhs, beta, result) ->  dgb_arakawa           (TWO kernels)
//   dg::MultiMatrix<DMatrix, DVec>::symv( alpha, x, beta, y) of a factor-2 projection / interpolation -> dgb_multimatrix2_symv (ONE kernel)
// dg::MultigridCG2d::solve reaches the second hook through the dg::PCG objects it owns.  Results are bitwise those of the
// un-hooked classes (tests/test_gpu_shim.py compares both builds with the OpenMP backend).
// The plan of an operator lives in the operator (EllipticPlanCache member); it is dropped when the operator is copied or its
// tensor is replaced, and it is only used when the operator has the shape the fused kernels cover: a 2-d grid, the
// derivative matrices of dx.h, a unit tensor chi (Cartesian metric); everything else falls through to the reference's code.
#pragma once
#include <vector>
#include <type_traits>
#include <thrust/device_vector.h>
#include "dgb_shim.h"
#include "tensor_traits.h"
#include "execution_policy.h"

namespace dgb
{
namespace shim
{
// a contiguous double vector living on the device (thrust::device_vector<double>, dg::View of one, ...)
template<class V, class = void>
struct is_device_dvec : std::false_type {};
template<class V>
struct is_device_dvec<V, std::enable_if_t<
    std::is_base_of<dg::SharedVectorTag, dg::get_tensor_category<V>>::value &&
    std::is_same<dg::get_execution_policy<V>, dg::CudaTag>::value &&
    std::is_same<std::remove_cv_t<dg::get_value_type<V>>, double>::value>> : std::true_type {};
template<class... Vs>
struct all_device_dvec : std::conjunction<is_device_dvec<std::decay_t<Vs>>...> {};

// a double EllSparseBlockMat on the device that carries the launch-plan cache make_tree.py adds (dg::DMatrix)
template<class M, class = void>
struct is_device_ell : std::false_type {};
template<class M>
struct is_device_ell<M, std::void_t<decltype( std::declval<const M&>().m_dgb_cache)>> : std::bool_constant<
    std::is_same<dg::get_execution_policy<M>, dg::CudaTag>::value && std::is_same<dg::get_value_type<M>, double>::value> {};
template<class Mat> inline dgb_ell* ell_plan( const Mat& m);  // sparseblockmat_gpu_kernels.cuh

template<class V> inline const double* cptr( const V& v) { return thrust::raw_pointer_cast( v.data()); }
template<class V> inline double* mptr( V& v) { return thrust::raw_pointer_cast( v.data()); }

// host copy of a device EllSparseBlockMat in the layout dgb_elliptic2d_create reads
template<class Matrix>
struct HostEll
{
    std::vector<double> data;
    std::vector<int> cols, didx, range;
    dgb_ell_host h;
    explicit HostEll( const Matrix& m)
    {
        data.resize( m.data.size()); cols.resize( m.cols_idx.size()); didx.resize( m.data_idx.size()); range.resize( 2);
        check( dgb_memcpy_d2h( data.data(), thrust::raw_pointer_cast( m.data.data()), data.size() * sizeof(double), nullptr), "dgb_memcpy_d2h");
        check( dgb_memcpy_d2h( cols.data(), thrust::raw_pointer_cast( m.cols_idx.data()), cols.size() * sizeof(int), nullptr), "dgb_memcpy_d2h");
        check( dgb_memcpy_d2h( didx.data(), thrust::raw_pointer_cast( m.data_idx.data()), didx.size() * sizeof(int), nullptr), "dgb_memcpy_d2h");
        check( dgb_memcpy_d2h( range.data(), thrust::raw_pointer_cast( m.right_range.data()), 2 * sizeof(int), nullptr), "dgb_memcpy_d2h");
        check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
        h.num_rows = m.num_rows; h.num_cols = m.num_cols; h.blocks_per_line = m.blocks_per_line; h.n = m.n;
        h.left_size = m.left_size; h.right_size = m.right_size; h.num_blocks = (int)(data.size() / ((size_t)m.n * m.n));
        h.right_range[0] = range[0]; h.right_range[1] = range[1];
        h.data = data.data(); h.cols_idx = cols.data(); h.data_idx = didx.data();
    }
};

// true if every element of the device vector equals `value` (two order-independent reductions in the library)
inline bool all_equal( const double* x, size_t n, double value)
{
    double mx = 0., mn = 0.;
    check( dgb_reduce( n, x, DGB_REDUCE_MAX, DGB_UNARY_IDENTITY, value, &mx, nullptr), "dgb_reduce");
    check( dgb_reduce( n, x, DGB_REDUCE_MIN, DGB_UNARY_IDENTITY, value, &mn, nullptr), "dgb_reduce");
    return mx == value && mn == value;
}

struct EllipticPlanCache
{
    EllipticPlanCache() = default;
    EllipticPlanCache( const EllipticPlanCache&) {}
    EllipticPlanCache& operator=( const EllipticPlanCache& o) { if( &o != this) forget(); return *this; }
    EllipticPlanCache( EllipticPlanCache&& o) noexcept { take( o); }
    EllipticPlanCache& operator=( EllipticPlanCache&& o) noexcept { if( &o != this) { forget(); take( o); } return *this; }
    ~EllipticPlanCache() { forget(); }
    void forget()
    {
        if( plan) dgb_elliptic2d_destroy( plan);
        plan = nullptr; state = 0;
    }
    dgb_elliptic2d* plan = nullptr;
    int state = 0;              // 0 not analysed, 1 fused plan ready, -1 this operator is outside the fused kernels' scope
    bool vol_is_one = false;
    const void* key[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // arrays the plan was built from
    private:
    void take( EllipticPlanCache& o) noexcept
    {
        plan = o.plan; state = o.state; vol_is_one = o.vol_is_one;
        for( int k = 0; k < 8; k++) key[k] = o.key[k];
        o.plan = nullptr; o.state = 0;
    }
};

// The plan of an Elliptic2d with the given members, or nullptr if the fused kernels do not cover it.  Cheap after the first
// call (pointer comparisons); the borrowed vectors (sigma, vol) are re-registered every time.
template<class Matrix, class Container, class Tensor>
inline dgb_elliptic2d* elliptic2d_plan( EllipticPlanCache& c, const Matrix& lx, const Matrix& ly, const Matrix& rx, const Matrix& ry,
    const Matrix& jx, const Matrix& jy, const Container& sigma, const Container& vol, const Tensor& chi, double jfactor, bool chi_weight_jump)
{
    if( !fusion_flag()) return nullptr;
    const void* now[8] = { thrust::raw_pointer_cast( lx.data.data()), thrust::raw_pointer_cast( ly.data.data()),
        thrust::raw_pointer_cast( rx.data.data()), thrust::raw_pointer_cast( ry.data.data()), thrust::raw_pointer_cast( jx.data.data()),
        thrust::raw_pointer_cast( jy.data.data()), chi.values().size() > 0 ? (const void*)thrust::raw_pointer_cast( chi.values()[0].data()) : nullptr,
        thrust::raw_pointer_cast( vol.data())};
    if( c.state != 0)
        for( int k = 0; k < 8; k++)
            if( c.key[k] != now[k]) { c.forget(); break; }
    if( c.state == 0)
    {
        for( int k = 0; k < 8; k++) c.key[k] = now[k];
        c.state = -1;
        if( chi_weight_jump) return nullptr;
        // unit tensor as SparseTensor( grid) builds it (topology/tensor.h:76-84): values {0, 1}, ones on the diagonal -- verified
        // on the data, not assumed
        if( chi.values().size() != 2 || chi.idx(0,0) != 1 || chi.idx(1,1) != 1 || chi.idx(0,1) != 0 || chi.idx(1,0) != 0) return nullptr;
        const size_t n = sigma.size();
        if( chi.values()[0].size() != n || chi.values()[1].size() != n || vol.size() != n) return nullptr;
        if( !all_equal( cptr( chi.values()[0]), n, 0.) || !all_equal( cptr( chi.values()[1]), n, 1.)) return nullptr;
        HostEll<Matrix> hlx( lx), hly( ly), hrx( rx), hry( ry), hjx( jx), hjy( jy);
        dgb_elliptic2d* plan = nullptr;
        if( dgb_elliptic2d_create( &plan, &hlx.h, &hly.h, &hrx.h, &hry.h, &hjx.h, &hjy.h, jfactor, 0) != 0) return nullptr;
        size_t size = 0; int fused = 0;
        dgb_elliptic2d_size( plan, &size, &fused);
        if( !fused || size != n) { dgb_elliptic2d_destroy( plan); return nullptr; }
        c.vol_is_one = all_equal( cptr( vol), n, 1.);
        c.plan = plan;
        c.state = 1;
    }
    if( c.state != 1) return nullptr;
    check( dgb_elliptic2d_set_sigma( c.plan, cptr( sigma)), "dgb_elliptic2d_set_sigma");
    check( dgb_elliptic2d_set_vol( c.plan, c.vol_is_one ? nullptr : cptr( vol)), "dgb_elliptic2d_set_vol");
    check( dgb_elliptic2d_set_jfactor( c.plan, jfactor), "dgb_elliptic2d_set_jfactor");
    check( dgb_elliptic2d_set_helmholtz( c.plan, 0, 0., nullptr), "dgb_elliptic2d_set_helmholtz");
    return c.plan;
}

// detection of the accessor the hooks add to the operator classes
template<class A, class = void> struct has_dgb_plan : std::false_type {};
template<class A> struct has_dgb_plan<A, std::void_t<decltype( std::declval<A&>().dgb_plan())>> : std::true_type {};

struct PcgCache
{
    PcgCache() = default;
    PcgCache( const PcgCache&) {}
    PcgCache& operator=( const PcgCache& o) { if( &o != this) forget(); return *this; }
    ~PcgCache() { forget(); }
    void forget() { if( pcg) dgb_pcg_destroy( pcg); pcg = nullptr; size = 0; }
    dgb_pcg* get( size_t n)
    {
        if( pcg && size != n) forget();
        if( !pcg) { check( dgb_pcg_create( &pcg, n), "dgb_pcg_create"); size = n; }
        return pcg;
    }
    dgb_pcg* pcg = nullptr;
    size_t size = 0;
};

// returns true if the solve was done by the library (iterations in `its`), false if the caller should run its own loop
template<class Operator, class X, class B, class P, class W>
inline bool pcg_solve( PcgCache& cache, Operator& A, X& x, const B& b, const P& precond, const W& weights, double eps, double nrmb_correction,
    int test_frequency, unsigned max_iter, bool throw_on_fail, unsigned& its)
{
    dgb_elliptic2d* plan = A.dgb_plan();
    if( !plan) return false;
    const size_t n = x.size();
    if( b.size() != n || precond.size() != n || weights.size() != n || test_frequency < 1) return false;
    if( ((uintptr_t)mptr( x) | (uintptr_t)cptr( b) | (uintptr_t)cptr( precond) | (uintptr_t)cptr( weights)) & 15u) return false;
    int iterations = 0;
    const int code = dgb_pcg_solve_elliptic2d( cache.get( n), plan, mptr( x), cptr( b), cptr( precond), cptr( weights), eps, nrmb_correction,
        test_frequency, (int)max_iter, &iterations, nullptr);
    note_library();
    its = (unsigned)iterations;
    if( code == DGB_ERR_NOCONVERGE)
    {
        if( throw_on_fail)
            throw dg::Fail( eps, dg::Message(_ping_) << "After " << max_iter << " PCG iterations with rtol " << eps << " and atol " << eps * nrmb_correction);
        its = max_iter;
        return true;
    }
    check( code, "dg::PCG::solve");
    return true;
}

// dg::MultiMatrix::symv of dimension 2: one kernel for the factor-2 projections / interpolations of dg::create::fast_projection /
// fast_interpolation (what dg::NestedGrids builds); false for every other pair (the caller runs its two products)
struct MultiCache
{
    const void *px = nullptr, *py = nullptr;  // launch plans the classification belongs to
    int kind = 0;
};
template<class Matrix>
inline bool multimatrix2_symv( MultiCache& c, const Matrix& mx, const Matrix& my, double alpha, const double* x, double beta, double* y)
{
    if( !fusion_flag() || x == y) return false;
    dgb_ell* px = ell_plan( mx);
    dgb_ell* py = ell_plan( my);
    if( c.px != px || c.py != py)
    {
        check( dgb_multimatrix2_fused( px, py, &c.kind), "dgb_multimatrix2_fused");
        c.px = px; c.py = py;
    }
    if( c.kind == 0) return false;
    check( dgb_multimatrix2_symv( px, py, c.kind, alpha, x, beta, y, nullptr, nullptr), "dg::MultiMatrix::symv");
    note_library();
    return true;
}
// scratch owned by an operator object; copies of the operator start with their own (empty) scratch
struct ScratchHolder
{
    ScratchHolder() = default;
    ScratchHolder( const ScratchHolder&) {}
    ScratchHolder& operator=( const ScratchHolder&) { return *this; }
    DeviceScratch<double> work;
};
// dg::Advection::upwind in one kernel; false if the four matrices are not the derivatives the kernel covers (the caller then
// runs the reference's sequence)
template<class Matrix>
inline bool advection_upwind( const Matrix& dxb, const Matrix& dxf, const Matrix& dyb, const Matrix& dyf, double alpha,
    const double* vx, const double* vy, const double* f, double beta, double* result)
{
    if( !fusion_flag() || f == result) return false;
    const int code = dgb_advection_upwind( ell_plan( dxb), ell_plan( dxf), ell_plan( dyb), ell_plan( dyf), alpha, vx, vy, f, beta, result, nullptr);
    if( code == DGB_ERR_UNSUPPORTED) return false;
    check( code, "dg::Advection::upwind");
    note_library();
    return true;
}
// dg::ArakawaX::operator() in two kernels; `work` holds the three mixed fields (the class' own temporaries are separate
// vectors, the kernels want one allocation); false if the matrices are not the centered derivatives the kernels cover
template<class Matrix>
inline bool arakawa( DeviceScratch<double>& work, const Matrix& bdx, const Matrix& bdy, double alpha, const double* lhs, const double* rhs,
    const double* chi, size_t size, double beta, double* result)
{
    if( !fusion_flag()) return false;
    const int code = dgb_arakawa( ell_plan( bdx), ell_plan( bdy), alpha, lhs, rhs, chi, beta, result, work.get( 3 * size), nullptr);
    if( code == DGB_ERR_UNSUPPORTED) return false;
    check( code, "dg::ArakawaX::operator()");
    note_library();
    return true;
}

}//namespace shim
}//namespace dgb
