// Field-line interpolation matrices (dg::geo::Fieldaligned's IDMatrix m_plus / m_minus, inc/geometries/fieldaligned.h)
// in a layout made for the gather: "sliced ELL".  The CSR matrix of the reference (sparsematrix.h:305-345) is converted
// ONCE into slices of 32 consecutive rows; inside a slice the k-th entries of the 32 rows are adjacent in memory, so a
// warp (lane = row) reads indices and values with fully coalesced 128-/256-byte requests instead of 32 different cache
// lines per request (ncu on the thread-per-row CSR kernel: 18 sectors per request, 64 % of all L1 sectors were matrix
// entries).  One thread then keeps PL planes of accumulators in registers and re-uses every (index, value) pair PL
// times; the same 2-d matrix serves all Nz planes.  Summation order = CSR order (k ascending), so the results are
// bitwise those of the CSR kernels / the reference's OpenMP kernel (sparsematrix_omp.h:17-52).
#include "common.cuh"
#include <vector>
#include <algorithm>

namespace dgb {

struct GatherPlan {
    int num_rows = 0, num_cols = 0, nslices = 0;
    long long nnz = 0, padded = 0;
    long long* slice_off = nullptr;  // [nslices + 1] first entry of a slice (entries are stored [k][lane])
    int* len = nullptr;              // [num_rows] entries per row
    int* idx = nullptr;              // [padded]
    double* val = nullptr;           // [padded]
};

__global__ void __launch_bounds__(128) gather_convert_kernel(int num_rows, const int* __restrict__ pos, const int* __restrict__ idx,
                                                             const double* __restrict__ val, const long long* __restrict__ slice_off,
                                                             int* __restrict__ len, int* __restrict__ eidx, double* __restrict__ eval) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= num_rows) return;
    const int b = pos[row], e = pos[row + 1];
    len[row] = e - b;
    const long long base = slice_off[row >> 5] + (row & 31);
    for (int k = 0; k < e - b; k++) {
        eidx[base + 32ll * k] = idx[b + k];
        eval[base + 32ll * k] = val[b + k];
    }
}

// y[pl] = alpha M x[(pl + shift) mod nplanes] + beta y[pl]  with the reference's order: beta == 1 accumulates into y,
// otherwise t = sum, y = fma(beta, y, t) (beta == 0: y is not read)
template <int PL>
__global__ void __launch_bounds__(128)
gather_planes_kernel(int num_rows, int num_cols, const long long* __restrict__ slice_off, const int* __restrict__ len,
                     const int* __restrict__ eidx, const double* __restrict__ eval, double alpha, const double* __restrict__ x,
                     double beta, double* __restrict__ y, int nplanes, int shift) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * PL;
    if (row >= num_rows) return;
    const double* xp[PL];
    double acc[PL];
#pragma unroll
    for (int p = 0; p < PL; p++) {
        const int pl = min(p0 + p, nplanes - 1);
        int src = (pl + shift) % nplanes;
        if (src < 0) src += nplanes;
        xp[p] = x + (size_t)src * num_cols;
        acc[p] = beta == 1. ? y[(size_t)pl * num_rows + row] : 0.;
    }
    const long long base = slice_off[row >> 5] + (row & 31);
    const int n = len[row];
    for (int k = 0; k < n; k++) {
        const double av = __dmul_rn(alpha, __ldg(eval + base + 32ll * k));
        const int j = __ldg(eidx + base + 32ll * k);
#pragma unroll
        for (int p = 0; p < PL; p++) acc[p] = __fma_rn(av, __ldg(xp[p] + j), acc[p]);
    }
#pragma unroll
    for (int p = 0; p < PL; p++) {
        if (p0 + p >= nplanes) break;
        double* yp = y + (size_t)(p0 + p) * num_rows + row;
        if (beta == 1. || beta == 0.) *yp = acc[p];
        else *yp = __fma_rn(beta, *yp, acc[p]);
    }
}

// DS::centered for periodic z (ds.h:481-485): g = alpha bphi (I+ f[k+1] - I- f[k-1]) / 2 / dphi + beta g
template <int PL>
__global__ void __launch_bounds__(128)
gather_ds_centered_kernel(int num_rows, int nplanes, const long long* __restrict__ poff, const int* __restrict__ plen,
                          const int* __restrict__ pidx, const double* __restrict__ pval, const long long* __restrict__ moff,
                          const int* __restrict__ mlen, const int* __restrict__ midx, const double* __restrict__ mval, double alpha,
                          const double* __restrict__ f, const double* __restrict__ bphi, double delta, double beta,
                          double* __restrict__ g) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * PL;
    if (row >= num_rows) return;
    double fp[PL], fm[PL];
    const double *xp[PL], *xm[PL];
#pragma unroll
    for (int p = 0; p < PL; p++) {
        const int pl = min(p0 + p, nplanes - 1);
        const int up = pl + 1 == nplanes ? 0 : pl + 1, dn = pl == 0 ? nplanes - 1 : pl - 1;
        xp[p] = f + (size_t)up * num_rows;
        xm[p] = f + (size_t)dn * num_rows;
        fp[p] = 0.;
        fm[p] = 0.;
    }
    {
        const long long base = poff[row >> 5] + (row & 31);
        const int n = plen[row];
        for (int k = 0; k < n; k++) {
            const double v = __dmul_rn(1., __ldg(pval + base + 32ll * k));
            const int j = __ldg(pidx + base + 32ll * k);
#pragma unroll
            for (int p = 0; p < PL; p++) fp[p] = __fma_rn(v, __ldg(xp[p] + j), fp[p]);
        }
    }
    {
        const long long base = moff[row >> 5] + (row & 31);
        const int n = mlen[row];
        for (int k = 0; k < n; k++) {
            const double v = __dmul_rn(1., __ldg(mval + base + 32ll * k));
            const int j = __ldg(midx + base + 32ll * k);
#pragma unroll
            for (int p = 0; p < PL; p++) fm[p] = __fma_rn(v, __ldg(xm[p] + j), fm[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < PL; p++) {
        if (p0 + p >= nplanes) break;
        const size_t i = (size_t)(p0 + p) * num_rows + row;
        // ds_centered (ds.h:776-786), same operation order as ds.cu ds_formula<DS_CENTERED>
        const double v = __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(alpha, bphi[i]), __dsub_rn(fp[p], fm[p])), 2.), delta);
        g[i] = beta == 0. ? v : __dadd_rn(v, __dmul_rn(beta, g[i]));
    }
}

// measured (B200, n = 3, 96 x 96 x 64): 4 planes per thread are best for ~36 entries per row ("dg" interpolation), 8 for ~81
// ("cubic"): longer rows amortise their index loads over more planes
static int planes_per_thread(const GatherPlan* P) {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DGB_GATHER_PLANES"); v = e ? atoi(e) : 0; if (v != 2 && v != 4 && v != 8) v = 0; }
    if (v) return v;
    return P->nnz > 56ll * P->num_rows ? 8 : 4;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_gather_plan_create(dgb_gather_plan** out, int num_rows, int num_cols, const int* pos_dev, const int* idx_dev,
                           const double* val_dev, dgb_stream_t s) {
    if (num_rows < 0 || num_cols < 0) { set_error("dgb_gather_plan_create: negative size"); return DGB_ERR_INVALID; }
    cudaStream_t st = as_stream(s);
    GatherPlan* P = new GatherPlan();
    P->num_rows = num_rows; P->num_cols = num_cols; P->nslices = (num_rows + 31) / 32;
    std::vector<int> pos(num_rows + 1, 0);
    if (num_rows > 0) {
        DGB_CUDA(cudaMemcpyAsync(pos.data(), pos_dev, sizeof(int) * (num_rows + 1), cudaMemcpyDeviceToHost, st));
        DGB_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<long long> off(P->nslices + 1, 0);
    for (int sl = 0; sl < P->nslices; sl++) {
        int mx = 0;
        for (int r = sl * 32; r < std::min(num_rows, sl * 32 + 32); r++) mx = std::max(mx, pos[r + 1] - pos[r]);
        off[sl + 1] = off[sl] + 32ll * mx;
    }
    P->nnz = pos[num_rows];
    P->padded = off[P->nslices];
    DGB_CUDA(cudaMalloc(&P->slice_off, sizeof(long long) * (P->nslices + 1)));
    DGB_CUDA(cudaMalloc(&P->len, sizeof(int) * std::max(num_rows, 1)));
    DGB_CUDA(cudaMalloc(&P->idx, sizeof(int) * std::max<long long>(P->padded, 1)));
    DGB_CUDA(cudaMalloc(&P->val, sizeof(double) * std::max<long long>(P->padded, 1)));
    DGB_CUDA(cudaMemsetAsync(P->idx, 0, sizeof(int) * std::max<long long>(P->padded, 1), st));
    DGB_CUDA(cudaMemsetAsync(P->val, 0, sizeof(double) * std::max<long long>(P->padded, 1), st));
    DGB_CUDA(cudaMemcpyAsync(P->slice_off, off.data(), sizeof(long long) * (P->nslices + 1), cudaMemcpyHostToDevice, st));
    DGB_CUDA(cudaStreamSynchronize(st));  // `off` is pageable host memory
    if (num_rows > 0) {
        gather_convert_kernel<<<(num_rows + 127) / 128, 128, 0, st>>>(num_rows, pos_dev, idx_dev, val_dev, P->slice_off, P->len, P->idx, P->val);
        DGB_LAUNCHED();
    }
    *out = reinterpret_cast<dgb_gather_plan*>(P);
    return 0;
}
int dgb_gather_plan_destroy(dgb_gather_plan* h) {
    GatherPlan* P = reinterpret_cast<GatherPlan*>(h);
    if (!P) return 0;
    cudaFree(P->slice_off); cudaFree(P->len); cudaFree(P->idx); cudaFree(P->val);
    delete P;
    return 0;
}
int dgb_gather_spmv_planes(const dgb_gather_plan* h, double alpha, const double* x, double beta, double* y, int nplanes, int shift,
                           dgb_stream_t s) {
    const GatherPlan* P = reinterpret_cast<const GatherPlan*>(h);
    if (!P) { set_error("dgb_gather_spmv_planes: NULL plan"); return DGB_ERR_INVALID; }
    if (P->num_rows == 0 || nplanes <= 0) return 0;
    if (x == y) { set_error("dgb_gather_spmv_planes: x must not alias y"); return DGB_ERR_INVALID; }
    cudaStream_t st = as_stream(s);
    const int pl = nplanes >= 8 ? planes_per_thread(P) : (nplanes >= 4 ? 4 : (nplanes >= 2 ? 2 : 1));
    dim3 grid((P->num_rows + 127) / 128, (nplanes + pl - 1) / pl);
#define DGB_GP(PLV) gather_planes_kernel<PLV><<<grid, 128, 0, st>>>(P->num_rows, P->num_cols, P->slice_off, P->len, P->idx, P->val, alpha, x, beta, y, nplanes, shift)
    if (pl == 8) DGB_GP(8); else if (pl == 4) DGB_GP(4); else if (pl == 2) DGB_GP(2); else DGB_GP(1);
#undef DGB_GP
    DGB_LAUNCHED();
    return 0;
}
int dgb_gather_ds_centered(const dgb_gather_plan* plus, const dgb_gather_plan* minus, int nplanes, double alpha, const double* f,
                           const double* bphi, double delta_phi, double beta, double* g, dgb_stream_t s) {
    const GatherPlan* P = reinterpret_cast<const GatherPlan*>(plus);
    const GatherPlan* M = reinterpret_cast<const GatherPlan*>(minus);
    if (!P || !M || P->num_rows != M->num_rows || P->num_cols != P->num_rows || M->num_cols != M->num_rows) {
        set_error("dgb_gather_ds_centered: plans must be square and of equal size");
        return DGB_ERR_INVALID;
    }
    if (P->num_rows == 0 || nplanes <= 0) return 0;
    if (f == g) { set_error("dgb_gather_ds_centered: f must not alias g"); return DGB_ERR_INVALID; }
    cudaStream_t st = as_stream(s);
    const int pl = nplanes >= 8 ? planes_per_thread(P) : (nplanes >= 4 ? 4 : (nplanes >= 2 ? 2 : 1));
    dim3 grid((P->num_rows + 127) / 128, (nplanes + pl - 1) / pl);
#define DGB_GD(PLV) gather_ds_centered_kernel<PLV><<<grid, 128, 0, st>>>(P->num_rows, nplanes, P->slice_off, P->len, P->idx, P->val, M->slice_off, M->len, M->idx, M->val, alpha, f, bphi, delta_phi, beta, g)
    if (pl == 8) DGB_GD(8); else if (pl == 4) DGB_GD(4); else if (pl == 2) DGB_GD(2); else DGB_GD(1);
#undef DGB_GD
    DGB_LAUNCHED();
    return 0;
}
}
