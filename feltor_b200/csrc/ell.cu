// EllSparseBlockMat / CooSparseBlockMat symv:  y = alpha (1_left (x) M (x) 1_right) x + beta y.
// Replaces launch_multiply_kernel(CudaTag,...) (inc/dg/backend/sparseblockmat_gpu_kernels.cuh:8-354).
// Per output element the rounding sequence is the reference OpenMP kernel's
// (sparseblockmat_omp_kernels.h:36-50): y = beta==0 ? 0 : y*beta; for d: temp = fma-chain over q; y = fma(alpha,temp,y).
//
// Kernels
//  ell_x_kernel   right_size == 1 (derivative along the contiguous dimension): one thread per block-row computes
//                 the n outputs of its cell from the bpl neighbouring cells held in registers.
//  ell_y_kernel   right_size  > 1: one thread per (block-row, column j[, j+1]); all accesses coalesced along j.
//  In both, rows that follow the matrix' interior pattern take their n x n blocks from the kernel parameter
//  space (constant-bank DFMA operands); boundary rows read data_idx/cols_idx (warp/block-uniform branch).
//  ell_generic_kernel  any n / blocks_per_line / pattern: one thread per output element.
#include "ell.cuh"

namespace dgb {

template <int N, int BPL>
__global__ void __launch_bounds__(256)
ell_x_kernel(EllArgs a, EllCoef<N, BPL> cf, double alpha, double beta, const double* __restrict__ x,
             double* __restrict__ y) {
    const size_t total = (size_t)a.left * a.num_rows;
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(r / a.num_rows);
        const int i = (int)(r - (size_t)s * a.num_rows);
        double out[N];
        double* yp = y + r * N;
        if (beta == 0.) {
#pragma unroll
            for (int k = 0; k < N; k++) out[k] = 0.;
        } else {
#pragma unroll
            for (int k = 0; k < N; k++) out[k] = __dmul_rn(yp[k], beta);
        }
        if (i >= a.i_lo && i < a.i_hi) {
#pragma unroll
            for (int d = 0; d < BPL; d++) {
                const double* xp = x + ((size_t)s * a.num_cols + (i + a.off[d])) * N;
                double xv[N];
#pragma unroll
                for (int q = 0; q < N; q++) xv[q] = xp[q];
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double temp = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) temp = __fma_rn(cf.c[d][k][q], xv[q], temp);
                    out[k] = __fma_rn(alpha, temp, out[k]);
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < BPL; d++) {
                const int C = a.cols[i * BPL + d];
                if (C == -1) continue;
                const double* xp = x + ((size_t)s * a.num_cols + C) * N;
                const double* bp = a.data + (size_t)a.didx[i * BPL + d] * N * N;
                double xv[N];
#pragma unroll
                for (int q = 0; q < N; q++) xv[q] = xp[q];
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double temp = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) temp = __fma_rn(__ldg(bp + k * N + q), xv[q], temp);
                    out[k] = __fma_rn(alpha, temp, out[k]);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < N; k++) yp[k] = out[k];
    }
}

template <int V> struct Vec;
template <> struct Vec<1> {
    double v[1];
    __device__ __forceinline__ void load(const double* p) { v[0] = *p; }
    __device__ __forceinline__ void store(double* p) const { *p = v[0]; }
};
template <> struct Vec<2> {
    double v[2];
    __device__ __forceinline__ void load(const double* p) { double2 t = ld2(p); v[0] = t.x; v[1] = t.y; }
    __device__ __forceinline__ void store(double* p) const { st2(p, make_double2(v[0], v[1])); }
};

template <int N, int BPL, int V>
__global__ void __launch_bounds__(256)
ell_y_kernel(EllArgs a, EllCoef<N, BPL> cf, double alpha, double beta, const double* __restrict__ x,
             double* __restrict__ y, int chunks) {
    const size_t b = blockIdx.x;
    const size_t r = b / chunks;
    const int c = (int)(b - r * chunks);
    const int s = (int)(r / a.num_rows);
    const int i = (int)(r - (size_t)s * a.num_rows);
    const int j = a.rr0 + (c * (int)blockDim.x + (int)threadIdx.x) * V;
    if (j >= a.rr1) return;
    const size_t right = a.right;
    Vec<V> out[N];
    double* yp = y + ((size_t)r * N) * right + j;
    if (beta == 0.) {
#pragma unroll
        for (int k = 0; k < N; k++)
#pragma unroll
            for (int v = 0; v < V; v++) out[k].v[v] = 0.;
    } else {
#pragma unroll
        for (int k = 0; k < N; k++) {
            out[k].load(yp + k * right);
#pragma unroll
            for (int v = 0; v < V; v++) out[k].v[v] = __dmul_rn(out[k].v[v], beta);
        }
    }
    if (i >= a.i_lo && i < a.i_hi) {
#pragma unroll
        for (int d = 0; d < BPL; d++) {
            const double* xp = x + (((size_t)s * a.num_cols + (i + a.off[d])) * N) * right + j;
            Vec<V> xv[N];
#pragma unroll
            for (int q = 0; q < N; q++) xv[q].load(xp + q * right);
#pragma unroll
            for (int k = 0; k < N; k++)
#pragma unroll
                for (int v = 0; v < V; v++) {
                    double temp = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) temp = __fma_rn(cf.c[d][k][q], xv[q].v[v], temp);
                    out[k].v[v] = __fma_rn(alpha, temp, out[k].v[v]);
                }
        }
    } else {
#pragma unroll
        for (int d = 0; d < BPL; d++) {
            const int C = a.cols[i * BPL + d];
            if (C == -1) continue;
            const double* xp = x + (((size_t)s * a.num_cols + C) * N) * right + j;
            const double* bp = a.data + (size_t)a.didx[i * BPL + d] * N * N;
            Vec<V> xv[N];
#pragma unroll
            for (int q = 0; q < N; q++) xv[q].load(xp + q * right);
#pragma unroll
            for (int k = 0; k < N; k++)
#pragma unroll
                for (int v = 0; v < V; v++) {
                    double temp = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) temp = __fma_rn(__ldg(bp + k * N + q), xv[q].v[v], temp);
                    out[k].v[v] = __fma_rn(alpha, temp, out[k].v[v]);
                }
        }
    }
#pragma unroll
    for (int k = 0; k < N; k++) out[k].store(yp + k * right);
}

__global__ void __launch_bounds__(256)
ell_generic_kernel(EllArgs a, double alpha, double beta, const double* __restrict__ x, double* __restrict__ y) {
    const int w = a.rr1 - a.rr0;
    const size_t size = (size_t)a.left * a.num_rows * a.n * w;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < size; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = a.rr0 + (int)(idx % w);
        size_t rest = idx / w;
        const int k = (int)(rest % a.n);
        rest /= a.n;
        const int i = (int)(rest % a.num_rows);
        const int s = (int)(rest / a.num_rows);
        const size_t I = (((size_t)s * a.num_rows + i) * a.n + k) * a.right + j;
        double yy = beta == 0. ? 0. : __dmul_rn(y[I], beta);
        for (int d = 0; d < a.bpl; d++) {
            const int C = a.cols[i * a.bpl + d];
            if (C == -1) continue;
            const size_t J = ((size_t)s * a.num_cols + C) * a.n;
            const double* bp = a.data + ((size_t)a.didx[i * a.bpl + d] * a.n + k) * a.n;
            double temp = 0.;
            for (int q = 0; q < a.n; q++) temp = __fma_rn(__ldg(bp + q), x[(J + q) * a.right + j], temp);
            yy = __fma_rn(alpha, temp, yy);
        }
        y[I] = yy;
    }
}

// CooSparseBlockMat (sparseblockmat_omp_kernels.h:356-377): y[I] = fma(alpha, temp, y[I]) entry by entry
__global__ void __launch_bounds__(256)
coo_kernel(dgb_coo m, double alpha, const double* const* __restrict__ x, double* __restrict__ y) {
    const size_t size = (size_t)m.left_size * m.n * m.right_size;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < size; idx += (size_t)gridDim.x * blockDim.x) {
        const int s = (int)(idx / ((size_t)m.n * m.right_size));
        const int k = (int)((idx % ((size_t)m.n * m.right_size)) / m.right_size);
        const int j = (int)(idx % m.right_size);
        for (int e = 0; e < m.num_entries; e++) {
            const size_t I = (((size_t)s * m.num_rows + m.rows_idx[e]) * m.n + k) * m.right_size + j;
            const double* xc = x[m.cols_idx[e]];
            const double* bp = m.data + ((size_t)m.data_idx[e] * m.n + k) * m.n;
            double temp = 0.;
            for (int q = 0; q < m.n; q++)
                temp = __fma_rn(__ldg(bp + q), xc[((size_t)q * m.left_size + s) * m.right_size + j], temp);
            y[I] = __fma_rn(alpha, temp, y[I]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------- host
int ell_upload(EllDev& m, const dgb_ell_host* h) {
    if (!h || h->n < 1 || h->blocks_per_line < 1 || h->num_rows < 1 || h->num_cols < 1 || h->num_blocks < 1) {
        set_error("dgb_ell_create: invalid matrix description");
        return DGB_ERR_INVALID;
    }
    m.num_rows = h->num_rows; m.num_cols = h->num_cols; m.bpl = h->blocks_per_line; m.n = h->n;
    m.left = h->left_size; m.right = h->right_size; m.nblocks = h->num_blocks;
    m.rr0 = h->right_range[0]; m.rr1 = h->right_range[1];
    size_t nd = (size_t)m.nblocks * m.n * m.n, ni = (size_t)m.num_rows * m.bpl;
    m.h_data.assign(h->data, h->data + nd);
    m.h_cols.assign(h->cols_idx, h->cols_idx + ni);
    m.h_didx.assign(h->data_idx, h->data_idx + ni);
    for (size_t t = 0; t < ni; t++) {
        if (m.h_cols[t] < -1 || m.h_cols[t] >= m.num_cols || m.h_didx[t] < 0 || m.h_didx[t] >= m.nblocks) {
            set_error("dgb_ell_create: index out of range at slot %zu (col %d, block %d)", t, m.h_cols[t], m.h_didx[t]);
            return DGB_ERR_INVALID;
        }
    }
    DGB_CUDA(cudaMalloc(&m.data, nd * sizeof(double)));
    DGB_CUDA(cudaMalloc(&m.cols, ni * sizeof(int)));
    DGB_CUDA(cudaMalloc(&m.didx, ni * sizeof(int)));
    DGB_CUDA(cudaMemcpy(m.data, m.h_data.data(), nd * sizeof(double), cudaMemcpyHostToDevice));
    DGB_CUDA(cudaMemcpy(m.cols, m.h_cols.data(), ni * sizeof(int), cudaMemcpyHostToDevice));
    DGB_CUDA(cudaMemcpy(m.didx, m.h_didx.data(), ni * sizeof(int), cudaMemcpyHostToDevice));
    // interior pattern: take the middle row as the template and grow the matching range around it
    m.has_pattern = false;
    if (m.bpl <= ELL_MAX_BPL && m.n <= ELL_MAX_N) {
        int mid = m.num_rows / 2;
        bool ok = true;
        for (int d = 0; d < m.bpl; d++) {
            int C = m.h_cols[(size_t)mid * m.bpl + d];
            if (C == -1) ok = false;
            m.off[d] = C - mid;
            m.did[d] = m.h_didx[(size_t)mid * m.bpl + d];
        }
        auto match = [&](int i) {
            for (int d = 0; d < m.bpl; d++)
                if (m.h_cols[(size_t)i * m.bpl + d] != i + m.off[d] || m.h_didx[(size_t)i * m.bpl + d] != m.did[d]) return false;
            return true;
        };
        if (ok) {
            int lo = mid, hi = mid + 1;
            while (lo > 0 && match(lo - 1)) lo--;
            while (hi < m.num_rows && match(hi)) hi++;
            m.i_lo = lo; m.i_hi = hi;
            m.has_pattern = true;
        }
    }
    return 0;
}
void ell_release(EllDev& m) {
    cudaFree(m.data); cudaFree(m.cols); cudaFree(m.didx);
    m.data = nullptr; m.cols = nullptr; m.didx = nullptr;
}

template <int N, int BPL>
static int ell_launch_fast(const EllDev& m, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    EllArgs a = ell_args(m);
    EllCoef<N, BPL> cf = ell_coef<N, BPL>(m);
    if (m.right == 1) {
        size_t total = (size_t)m.left * m.num_rows;
        size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 16;
        ell_x_kernel<N, BPL><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a, cf, alpha, beta, x, y);
    } else {
        int w = m.rr1 - m.rr0;
        bool v2 = (m.right % 2 == 0) && (m.rr0 % 2 == 0) && (w % 2 == 0) && aligned16(x) && aligned16(y);
        int V = v2 ? 2 : 1;
        int per = (w + V - 1) / V;
        int threads = per >= 256 ? 256 : ((per + 31) / 32) * 32;
        int chunks = (per + threads - 1) / threads;
        size_t blocks = (size_t)m.left * m.num_rows * chunks;
        if (blocks > 0x7fffffffull) { set_error("dgb_ell_symv: grid too large"); return DGB_ERR_UNSUPPORTED; }
        if (v2) ell_y_kernel<N, BPL, 2><<<(unsigned)blocks, threads, 0, st>>>(a, cf, alpha, beta, x, y, chunks);
        else ell_y_kernel<N, BPL, 1><<<(unsigned)blocks, threads, 0, st>>>(a, cf, alpha, beta, x, y, chunks);
    }
    DGB_LAUNCHED();
    return 0;
}
template <int N>
static int ell_launch_n(const EllDev& m, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    switch (m.bpl) {
        case 1: return ell_launch_fast<N, 1>(m, alpha, x, beta, y, st);
        case 2: return ell_launch_fast<N, 2>(m, alpha, x, beta, y, st);
        case 3: return ell_launch_fast<N, 3>(m, alpha, x, beta, y, st);
        case 4: return ell_launch_fast<N, 4>(m, alpha, x, beta, y, st);
    }
    return DGB_ERR_UNSUPPORTED;
}

int ell_symv(const EllDev& m, double alpha, const double* x, double beta, double* y, cudaStream_t st,
             bool force_generic) {
    if (x == y) { set_error("dgb_ell_symv: x must not alias y"); return DGB_ERR_INVALID; }
    if (m.rr1 <= m.rr0) return 0;
    bool fast = !force_generic && m.n <= ELL_MAX_N && m.bpl <= ELL_MAX_BPL && (m.right > 1 || (m.rr0 == 0 && m.rr1 == 1));
    if (fast) {
        switch (m.n) {
            case 1: return ell_launch_n<1>(m, alpha, x, beta, y, st);
            case 2: return ell_launch_n<2>(m, alpha, x, beta, y, st);
            case 3: return ell_launch_n<3>(m, alpha, x, beta, y, st);
            case 4: return ell_launch_n<4>(m, alpha, x, beta, y, st);
            case 5: return ell_launch_n<5>(m, alpha, x, beta, y, st);
        }
    }
    EllArgs a = ell_args(m);
    size_t size = (size_t)m.left * m.num_rows * m.n * (m.rr1 - m.rr0);
    size_t want = (size + 255) / 256, cap = (size_t)sm_count() * 16;
    ell_generic_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(a, alpha, beta, x, y);
    DGB_LAUNCHED();
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_ell_create(dgb_ell** out, const dgb_ell_host* host) {
    EllDev* m = new EllDev();
    int e = ell_upload(*m, host);
    if (e) { ell_release(*m); delete m; return e; }
    *out = reinterpret_cast<dgb_ell*>(m);
    return 0;
}
int dgb_ell_destroy(dgb_ell* p) {
    EllDev* m = reinterpret_cast<EllDev*>(p);
    if (!m) return 0;
    ell_release(*m);
    delete m;
    return 0;
}
int dgb_ell_set_left_size(dgb_ell* p, int left) { reinterpret_cast<EllDev*>(p)->left = left; return 0; }
int dgb_ell_set_right_size(dgb_ell* p, int right) {  // sparseblockmat.h:146-150
    EllDev* m = reinterpret_cast<EllDev*>(p);
    m->right = right; m->rr0 = 0; m->rr1 = right;
    return 0;
}
int dgb_ell_set_right_range(dgb_ell* p, int begin, int end) {
    EllDev* m = reinterpret_cast<EllDev*>(p);
    if (begin < 0 || end > m->right || begin > end) { set_error("dgb_ell_set_right_range: invalid range"); return DGB_ERR_INVALID; }
    m->rr0 = begin; m->rr1 = end;
    return 0;
}
int dgb_ell_total_num_rows(const dgb_ell* p, size_t* rows) { *rows = reinterpret_cast<const EllDev*>(p)->total_rows(); return 0; }
int dgb_ell_total_num_cols(const dgb_ell* p, size_t* cols) { *cols = reinterpret_cast<const EllDev*>(p)->total_cols(); return 0; }
int dgb_ell_symv(const dgb_ell* p, double alpha, const double* x, double beta, double* y, dgb_stream_t s) {
    return ell_symv(*reinterpret_cast<const EllDev*>(p), alpha, x, beta, y, as_stream(s), false);
}
int dgb_ell_symv_generic(const dgb_ell* p, double alpha, const double* x, double beta, double* y, dgb_stream_t s) {
    return ell_symv(*reinterpret_cast<const EllDev*>(p), alpha, x, beta, y, as_stream(s), true);
}
int dgb_coo_symv(const dgb_coo* m, double alpha, const double* const* x, double beta, double* y, dgb_stream_t s) {
    if (m->num_entries == 0) return 0;
    if (beta != 1.) { set_error("dgb_coo_symv: beta must be 1 (sparseblockmat.h:324)"); return DGB_ERR_INVALID; }
    size_t size = (size_t)m->left_size * m->n * m->right_size;
    size_t want = (size + 255) / 256, cap = (size_t)sm_count() * 8;
    coo_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(s)>>>(*m, alpha, x, y);
    DGB_LAUNCHED();
    return 0;
}
}
