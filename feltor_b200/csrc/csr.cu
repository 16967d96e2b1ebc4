// CSR sparse matrix - vector product with the reference's OpenMP summation order (bit-reproducible), and the
// "same 2-D matrix applied to all planes" variant that dg::geo::Fieldaligned needs.
// Replaces detail::spmv_gpu_kernel = cusparseSpMV(CSR_ALG1) (inc/dg/backend/sparsematrix_gpu.cuh:190-214), whose
// result is not reproducible, by the order of inc/dg/backend/sparsematrix_omp.h:17-52:
//   beta == 1:  y = fma(alpha*v_k, x[j_k], y)  for k ascending;
//   else:       t = 0; t = fma(alpha*v_k, x[j_k], t);  y = fma(beta, y, t)    (beta == 0: y is not read)
// Layout: one thread owns one row and PL planes; the row's (col, val) pairs are read once and reused for the PL
// planes held in registers, neighbouring threads own neighbouring rows, so the x-gathers of a warp fall into the
// few cache lines that hold the (spatially local) interpolation stencil.
#include "common.cuh"

namespace dgb {

template <int PL>
__global__ void __launch_bounds__(128)
csr_planes_kernel(int num_rows, int num_cols, const int* __restrict__ pos, const int* __restrict__ idx,
                  const double* __restrict__ val, double alpha, const double* __restrict__ x, double beta,
                  double* __restrict__ y, int nplanes, int shift) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * PL;
    if (row >= num_rows) return;
    const double* xp[PL];
    double acc[PL];
    double* yp[PL];
#pragma unroll
    for (int p = 0; p < PL; p++) {
        int pl = p0 + p;
        bool ok = pl < nplanes;
        int src = ok ? (pl + shift) % nplanes : 0;
        if (src < 0) src += nplanes;
        xp[p] = x + (size_t)src * num_cols;
        yp[p] = ok ? y + (size_t)pl * num_rows + row : nullptr;
        acc[p] = (beta == 1. && ok) ? *yp[p] : 0.;
    }
    const int b = pos[row], e = pos[row + 1];
    for (int jj = b; jj < e; jj++) {
        const double av = __dmul_rn(alpha, __ldg(val + jj));
        const int j = __ldg(idx + jj);
#pragma unroll
        for (int p = 0; p < PL; p++) acc[p] = __fma_rn(av, __ldg(xp[p] + j), acc[p]);
    }
#pragma unroll
    for (int p = 0; p < PL; p++) {
        if (!yp[p]) continue;
        if (beta == 1. || beta == 0.) *yp[p] = acc[p];
        else *yp[p] = __fma_rn(beta, *yp[p], acc[p]);
    }
}

// pack of a gather: out[i] = x[idx[i]] (the send buffer of MPIGather, mpi_gather.h:454-705)
__global__ void __launch_bounds__(256) gather_indexed_kernel(size_t n, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = __ldg(x + __ldg(idx + i));
}
// the outer part of MPIDistMat::symv (mpi_matrix.h:505-521) in one kernel: comp[i] = (outer matrix row i) . buffer with the
// CSR order of sparsematrix_omp.h:39-48 (alpha = 1, beta = 0), then y[scatter[i]] += comp[i].  Rows of the outer matrix map to
// distinct rows of y, so no atomics.
__global__ void __launch_bounds__(128)
csr_scatter_add_kernel(int num_rows, const int* __restrict__ pos, const int* __restrict__ idx, const double* __restrict__ val,
                       const double* __restrict__ buffer, const int* __restrict__ scatter, double* __restrict__ y) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    double t = 0.;
    for (int jj = __ldg(pos + row); jj < __ldg(pos + row + 1); jj++) t = __fma_rn(__dmul_rn(1., __ldg(val + jj)), __ldg(buffer + __ldg(idx + jj)), t);
    const int dst = __ldg(scatter + row);
    y[dst] = __dadd_rn(y[dst], t);
}

// the allreduce mode of MPIDistMat (mpi_matrix.h:438-441,487-490: every rank applies its column block of the matrix, the
// partial results are summed over the ranks): y[i] = ((part_0[i] + part_1[i]) + part_2[i]) + ... in RANK ORDER on every rank,
// so all ranks hold the same bits whatever the arrival order (MPI_Allreduce of doubles gives no such guarantee).
__global__ void __launch_bounds__(256) sum_ranks_kernel(int nranks, size_t m, const double* __restrict__ parts, double* __restrict__ y) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (size_t)gridDim.x * blockDim.x) {
        double t = __ldg(parts + i);
        for (int r = 1; r < nranks; r++) t = __dadd_rn(t, __ldg(parts + (size_t)r * m + i));
        y[i] = t;
    }
}

static int csr_launch(int num_rows, int num_cols, const int* pos, const int* idx, const double* val, double alpha,
                      const double* x, double beta, double* y, int nplanes, int shift, cudaStream_t st) {
    if (num_rows < 0 || num_cols < 0 || nplanes < 0) { set_error("dgb_csr_spmv: negative size"); return DGB_ERR_INVALID; }
    if (num_rows == 0 || nplanes == 0) return 0;
    if (x == y) { set_error("dgb_csr_spmv: x must not alias y"); return DGB_ERR_INVALID; }
    dim3 block(128);
    if (nplanes >= 4) {
        dim3 grid((num_rows + 127) / 128, (nplanes + 3) / 4);
        csr_planes_kernel<4><<<grid, block, 0, st>>>(num_rows, num_cols, pos, idx, val, alpha, x, beta, y, nplanes, shift);
    } else {
        dim3 grid((num_rows + 127) / 128, nplanes);
        csr_planes_kernel<1><<<grid, block, 0, st>>>(num_rows, num_cols, pos, idx, val, alpha, x, beta, y, nplanes, shift);
    }
    DGB_LAUNCHED();
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_csr_spmv(int num_rows, int num_cols, const int* pos, const int* idx, const double* val, double alpha,
                 const double* x, double beta, double* y, dgb_stream_t s) {
    return csr_launch(num_rows, num_cols, pos, idx, val, alpha, x, beta, y, 1, 0, as_stream(s));
}
int dgb_csr_spmv_planes(int num_rows, int num_cols, const int* pos, const int* idx, const double* val, double alpha,
                        const double* x, double beta, double* y, int nplanes, int shift, dgb_stream_t s) {
    return csr_launch(num_rows, num_cols, pos, idx, val, alpha, x, beta, y, nplanes, shift, as_stream(s));
}
int dgb_gather_indexed(size_t n, const int* idx, const double* x, double* out, dgb_stream_t s) {
    if (n == 0) return 0;
    if (!idx || !x || !out) { set_error("dgb_gather_indexed: NULL argument"); return DGB_ERR_INVALID; }
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 16;
    gather_indexed_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(s)>>>(n, idx, x, out);
    DGB_LAUNCHED();
    return 0;
}
int dgb_csr_spmv_scatter_add(int num_rows, const int* pos, const int* idx, const double* val, const double* buffer, const int* scatter,
                             double* y, dgb_stream_t s) {
    if (num_rows == 0) return 0;
    if (num_rows < 0 || !pos || !idx || !val || !buffer || !scatter || !y) { set_error("dgb_csr_spmv_scatter_add: invalid argument"); return DGB_ERR_INVALID; }
    csr_scatter_add_kernel<<<(num_rows + 127) / 128, 128, 0, as_stream(s)>>>(num_rows, pos, idx, val, buffer, scatter, y);
    DGB_LAUNCHED();
    return 0;
}
int dgb_sum_ranks(int nranks, size_t m, const double* parts, double* y, dgb_stream_t s) {
    if (m == 0) return 0;
    if (nranks < 1 || !parts || !y) { set_error("dgb_sum_ranks: invalid argument"); return DGB_ERR_INVALID; }
    size_t want = (m + 255) / 256, cap = (size_t)sm_count() * 16;
    sum_ranks_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(s)>>>(nranks, m, parts, y);
    DGB_LAUNCHED();
    return 0;
}
}
