// Replacement of inc/dg/backend/sparseblockmat_gpu_kernels.cuh: the dg::CudaTag members
//   EllSparseBlockMat::launch_multiply_kernel( CudaTag, alpha, x, beta, y)      sparseblockmat.h:180-186
//   CooSparseBlockMat::launch_multiply_kernel( CudaTag, alpha, x[], beta, y)    sparseblockmat.h:349-357
// bound to dgb_ell_symv / dgb_coo_symv of libdgb200.so for double matrices and vectors.  The Ell launch plan of the
// library (row classification, interior blocks as constant-bank operands) is built on first use and kept in the matrix
// (m_dgb_cache, the member integration/make_tree.py adds to the struct -- the same arrangement the reference has for
// CSR matrices with CSRCache_gpu).  Other value types (float, complex) run through the small general kernels below.
#pragma once
#include <vector>
#include <thrust/device_vector.h>
#include "dgb_shim.h"
#include "fma.h"

namespace dgb
{
namespace shim
{
// y[I] = beta y[I] + alpha sum_d ( B_d x )[I]: one thread per output element, any n / value type
// (summation order of sparseblockmat.h:363-388)
template<class real_type, class value_type>
__global__ void __launch_bounds__(256) ell_any_kernel( value_type alpha, value_type beta, const real_type* data, const int* cols_idx, const int* data_idx,
    int num_rows, int num_cols, int bpl, int n, int left, int right, int r0, int r1, const value_type* x, value_type* y)
{
    const int width = r1 - r0;
    const size_t total = (size_t)left * num_rows * n * width, stride = (size_t)gridDim.x * blockDim.x;
    for( size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
    {
        const int j = r0 + (int)(t % width);
        size_t rest = t / width;
        const int k = (int)(rest % n); rest /= n;
        const int i = (int)(rest % num_rows);
        const int s = (int)(rest / num_rows);
        const size_t I = (((size_t)s * num_rows + i) * n + k) * right + j;
        value_type acc = beta == value_type(0) ? value_type(0) : y[I] * beta;
        for( int d = 0; d < bpl; d++)
        {
            const int J = cols_idx[i * bpl + d];
            if( J < 0) continue;
            const real_type* blk = data + ((size_t)data_idx[i * bpl + d] * n + k) * n;
            value_type temp = value_type(0);
            for( int q = 0; q < n; q++)
                temp = dg::detail::dg_fma( blk[q], x[(((size_t)s * num_cols + J) * n + q) * right + j], temp);
            acc = dg::detail::dg_fma( alpha, temp, acc);
        }
        y[I] = acc;
    }
}
// y[I] += alpha ( B_e x_{J_e} ): entries in order per output element (sparseblockmat_omp_kernels.h:354-380)
template<class real_type, class value_type>
__global__ void __launch_bounds__(256) coo_any_kernel( const real_type* data, const int* rows_idx, const int* cols_idx, const int* data_idx,
    int num_rows, int num_entries, int n, int left, int right, value_type alpha, const value_type** x, value_type* y)
{
    const size_t total = (size_t)left * n * right, stride = (size_t)gridDim.x * blockDim.x;
    for( size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
    {
        const int j = (int)(t % right), k = (int)((t / right) % n), s = (int)(t / ((size_t)right * n));
        for( int e = 0; e < num_entries; e++)
        {
            const size_t I = (((size_t)s * num_rows + rows_idx[e]) * n + k) * right + j;
            const real_type* blk = data + ((size_t)data_idx[e] * n + k) * n;
            const value_type* xe = x[cols_idx[e]];
            value_type temp = value_type(0);
            for( int q = 0; q < n; q++)
                temp = dg::detail::dg_fma( blk[q], xe[((size_t)q * left + s) * right + j], temp);
            y[I] = dg::detail::dg_fma( alpha, temp, y[I]);
        }
    }
}
template<class T>
inline std::vector<T> to_host( const T* dev, size_t count)
{
    std::vector<T> h( count);
    if( count) check( dgb_memcpy_d2h( h.data(), dev, count * sizeof(T), nullptr), "dgb_memcpy_d2h");
    check( dgb_stream_synchronize( nullptr), "dgb_stream_synchronize");
    return h;
}
// The library's launch plan of a double EllSparseBlockMat on the device: built on first use from a host copy of the arrays,
// kept in the matrix (m_dgb_cache), rebuilt when the arrays or the Kronecker sizes changed.
template<class Mat>
inline dgb_ell* ell_plan( const Mat& m)
{
    const double* data_ptr = thrust::raw_pointer_cast( m.data.data());
    const int* cols_ptr = thrust::raw_pointer_cast( m.cols_idx.data());
    const int* block_ptr = thrust::raw_pointer_cast( m.data_idx.data());
    const int* range_ptr = thrust::raw_pointer_cast( m.right_range.data());
    EllCache& c = m.m_dgb_cache;
    if( !c.plan || c.data != data_ptr || c.cols != cols_ptr || c.didx != block_ptr || c.left != m.left_size || c.right != m.right_size)
    {
        c.forget();
        const std::vector<double> h_data = to_host( data_ptr, m.data.size());
        const std::vector<int> h_cols = to_host( cols_ptr, m.cols_idx.size());
        const std::vector<int> h_didx = to_host( block_ptr, m.data_idx.size());
        const std::vector<int> h_range = to_host( range_ptr, 2);
        dgb_ell_host h;
        h.num_rows = m.num_rows; h.num_cols = m.num_cols; h.blocks_per_line = m.blocks_per_line; h.n = m.n;
        h.left_size = m.left_size; h.right_size = m.right_size;
        h.num_blocks = (int)(m.data.size() / ((size_t)m.n * m.n));
        h.right_range[0] = h_range[0]; h.right_range[1] = h_range[1];
        h.data = h_data.data(); h.cols_idx = h_cols.data(); h.data_idx = h_didx.data();
        check( dgb_ell_create( &c.plan, &h), "dgb_ell_create");
        c.data = data_ptr; c.cols = cols_ptr; c.didx = block_ptr;
        c.left = m.left_size; c.right = m.right_size; c.r0 = h_range[0]; c.r1 = h_range[1];
    }
    return c.plan;
}
}//namespace shim
}//namespace dgb

namespace dg
{

template<class real_type, template<class> class Vector>
template<class value_type>
void EllSparseBlockMat<real_type, Vector>::launch_multiply_kernel( CudaTag, value_type alpha, const value_type* x_ptr, value_type beta, value_type* y_ptr) const
{
    const real_type* data_ptr = thrust::raw_pointer_cast( data.data());
    const int* cols_ptr = thrust::raw_pointer_cast( cols_idx.data());
    const int* block_ptr = thrust::raw_pointer_cast( data_idx.data());
    const int* range_ptr = thrust::raw_pointer_cast( right_range.data());
    if( num_rows == 0 || left_size == 0 || right_size == 0) return;
    if constexpr( std::is_same_v<real_type, double> && std::is_same_v<value_type, double>)
    {
        dgb_ell* plan = dgb::shim::ell_plan( *this);
        dgb::shim::check( dgb_ell_symv( plan, alpha, x_ptr, beta, y_ptr, nullptr), "dg::blas2::symv (EllSparseBlockMat)");
        dgb::shim::note_library();
    }
    else
    {
        const std::vector<int> h_range = dgb::shim::to_host( range_ptr, 2);
        const size_t total = (size_t)left_size * num_rows * n * (h_range[1] - h_range[0]);
        if( total == 0) return;
        dgb::shim::note_generic<value_type>( "ell symv");
        dgb::shim::ell_any_kernel<real_type, value_type><<<dgb::shim::generic_grid( total), 256>>>( alpha, beta, data_ptr, cols_ptr, block_ptr,
            num_rows, num_cols, blocks_per_line, n, left_size, right_size, h_range[0], h_range[1], x_ptr, y_ptr);
        dgb::shim::check_launch( "dg::blas2::symv (EllSparseBlockMat)");
    }
}

template<class real_type, template<class> class Vector>
template<class value_type>
void CooSparseBlockMat<real_type, Vector>::launch_multiply_kernel( CudaTag, value_type alpha, const value_type** x_ptr, value_type beta, value_type* y_ptr) const
{
    if( num_entries == 0)
        return;
    if( beta != value_type(1))
        throw dg::Error( dg::Message(_ping_) << "CooSparseBlockMat::symv needs beta == 1");  // sparseblockmat_gpu_kernels.cuh:321
    const real_type* data_ptr = thrust::raw_pointer_cast( data.data());
    const int* rows_ptr = thrust::raw_pointer_cast( rows_idx.data());
    const int* cols_ptr = thrust::raw_pointer_cast( cols_idx.data());
    const int* block_ptr = thrust::raw_pointer_cast( data_idx.data());
    if constexpr( std::is_same_v<real_type, double> && std::is_same_v<value_type, double>)
    {
        dgb_coo m;
        m.num_rows = num_rows; m.num_cols = num_cols; m.num_entries = num_entries; m.n = n;
        m.left_size = left_size; m.right_size = right_size;
        m.data = data_ptr; m.rows_idx = rows_ptr; m.cols_idx = cols_ptr; m.data_idx = block_ptr;
        dgb::shim::check( dgb_coo_symv( &m, alpha, x_ptr, beta, y_ptr, nullptr), "dg::blas2::symv (CooSparseBlockMat)");
        dgb::shim::note_library();
    }
    else
    {
        const size_t total = (size_t)left_size * n * right_size;
        dgb::shim::note_generic<value_type>( "coo symv");
        dgb::shim::coo_any_kernel<real_type, value_type><<<dgb::shim::generic_grid( total), 256>>>( data_ptr, rows_ptr, cols_ptr, block_ptr,
            num_rows, num_entries, n, left_size, right_size, alpha, x_ptr, y_ptr);
        dgb::shim::check_launch( "dg::blas2::symv (CooSparseBlockMat)");
    }
}

}//namespace dg
