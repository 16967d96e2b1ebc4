// Replacement of inc/dg/backend/sparsematrix_gpu.cuh (cuSPARSE SpMV, not reproducible -- CUSPARSE_SPMV_CSR_ALG1): the CSR
// matrix-vector product of the CUDA backend bound to libdgb200.so, which sums every row in CSR order exactly like the
// reference's OpenMP kernel (sparsematrix_omp.h:17-52) and therefore reproduces its results bit for bit.
//   detail::CSRCache_gpu                                              sparsematrix_gpu.cuh:93-187
//   detail::spmv_gpu_kernel( cache, rows, cols, nnz, pos, idx, val, alpha, beta, x, y)    sparsematrix_gpu.cuh:190-214
// No cuSPARSE handle, no link dependency on libcusparse.
#pragma once
#include <type_traits>
#include "dgb_shim.h"
#include "fma.h"

namespace dgb
{
namespace shim
{
// any index / value type: one thread per row, CSR order
template<class I, class V, class value_type, class C1, class C2>
__global__ void __launch_bounds__(256) csr_any_kernel( size_t num_rows, const I* pos, const I* idx, const V* val, value_type alpha, value_type beta, const C1* x, C2* y)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for( size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < num_rows; i += stride)
    {
        if( beta == value_type(1))   // sparsematrix_omp.h:23-33: accumulate onto y element by element
        {
            C2 acc = y[i];
            for( I k = pos[i]; k < pos[i + 1]; k++) acc = dg::detail::dg_fma( alpha * val[k], x[idx[k]], acc);
            y[i] = acc;
            continue;
        }
        C2 temp = C2(0);
        for( I k = pos[i]; k < pos[i + 1]; k++) temp = dg::detail::dg_fma( alpha * val[k], x[idx[k]], temp);
        y[i] = beta == value_type(0) ? temp : dg::detail::dg_fma( beta, y[i], temp);
    }
}
}//namespace shim
}//namespace dgb

namespace dg
{
namespace detail
{
// The cache slot of dg::SparseMatrix (sparsematrix.h:620-628).  The library's CSR kernel needs no analysis phase, so the
// cache only remembers that the matrix has been seen; copies start fresh like the reference's.
struct CSRCache_gpu
{
    CSRCache_gpu() = default;
    template<class I, class V>
    CSRCache_gpu( size_t num_rows, size_t num_cols, size_t nnz, const I* pos, const I* idx, const V* val) { update( num_rows, num_cols, nnz, pos, idx, val); }
    CSRCache_gpu( const CSRCache_gpu&) {}
    CSRCache_gpu( CSRCache_gpu&& src) noexcept { std::swap( m_active, src.m_active); }
    CSRCache_gpu& operator=( const CSRCache_gpu& src) { if( &src != this) m_active = false; return *this; }
    CSRCache_gpu& operator=( CSRCache_gpu&& src) noexcept { if( &src != this) { m_active = src.m_active; src.m_active = false; } return *this; }
    void forget() { m_active = false; }
    bool isUpToDate() const { return m_active; }
    template<class I, class V>
    void update( size_t, size_t, size_t, const I*, const I*, const V*) { m_active = true; }
    private:
    bool m_active = false;
};

//y = alpha A*x + beta y
template<class I, class V, class value_type, class C1, class C2>
void spmv_gpu_kernel( CSRCache_gpu& cache, size_t A_num_rows, size_t A_num_cols, size_t A_nnz,
    const I* A_pos, const I* A_idx, const V* A_val, value_type alpha, value_type beta, const C1* x_ptr, C2* y_ptr)
{
    if( A_num_rows == 0) return;
    if( not cache.isUpToDate())
        cache.update<I,V>( A_num_rows, A_num_cols, A_nnz, A_pos, A_idx, A_val);
    if constexpr( std::is_same_v<I, int> && std::is_same_v<V, double> && std::is_same_v<C1, double> && std::is_same_v<C2, double> && std::is_arithmetic_v<value_type>)
    {
        dgb::shim::check( dgb_csr_spmv( (int)A_num_rows, (int)A_num_cols, A_pos, A_idx, A_val, (double)alpha, x_ptr, (double)beta, y_ptr, nullptr), "dg::blas2::symv (SparseMatrix)");
        dgb::shim::note_library();
    }
    else
    {
        dgb::shim::note_generic<V>( "csr symv");
        dgb::shim::csr_any_kernel<I, V, value_type, C1, C2><<<dgb::shim::generic_grid( A_num_rows), 256>>>( A_num_rows, A_pos, A_idx, A_val, alpha, beta, x_ptr, y_ptr);
        dgb::shim::check_launch( "dg::blas2::symv (SparseMatrix)");
    }
}

} // namespace detail
} // namespace dg
