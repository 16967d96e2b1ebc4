// Fused Elliptic2d apply (see elliptic.cu for the overview).  One CTA = TX x TY cells, 256 threads.
//
// Rounding sequence replayed per output element (inc/dg/elliptic.h:431-458 on top of
// inc/dg/backend/sparseblockmat_omp_kernels.h:36-50 and inc/dg/topology/multiply.h:18-32), with
// blk(M,d) = fma-chain over q of M's block in slot d:
//   gx = 0; for d: gx = fma(1, blk(Rx,d), gx)            tx = fma(sigma, gx, gx*0)      (identity chi tensor)
//   gy likewise with Ry                                   ty = fma(sigma, gy, gy*0)
//   t  = 0; for d: t = fma(1, blk(Ly,d)[ty], t);  t = t*(-1);  for d: t = fma(-1, blk(Lx,d)[tx], t)
//   for d: t = fma(jfactor, blk(Jx,d)[x], t);  for d: t = fma(jfactor, blk(Jy,d)[x], t)
//   y = fma(alpha, t/vol, beta*y)
#include "elliptic.cuh"
#include "superacc.cuh"
#include "pcg.cuh"
#include <algorithm>

namespace dgb {

constexpr int TX = 32, TY = 8, FUSED_THREADS = TX * TY;

struct MatView {
    const double* data;
    const int* cols;
    const int* didx;
    int i_lo, i_hi, num;
    int off[3];
};

template <int N, int B>
struct EllipticCoef {
    double rx[B][N][N], ry[B][N][N], lx[B][N][N], ly[B][N][N], jx[3][N][N], jy[3][N][N];
};

struct FusedArgs {
    MatView rx, ry, lx, ly, jx, jy;
    int Nx, Ny, wrapx, wrapy;
    int txlo, txhi, tylo, tyhi;  // range of flux cells (relative to the tile) the adjoint derivatives reach
    const double* sigma;
    const double* vol;
    const double* x;
    double* y;
    double alpha, beta, jfactor;
    // optional fused dot(x, w, y) of the PCG step (pcg.h:165-166): products round(round(x*w)*y)
    const double* dot_w;
    sa::DotSlot slot;
    PcgState* pcg;
};

// out[k] = fma(a, sum_q C[d][k][q] * s[(off_d*N + q)*stride], out[k]) for the slots d of block-row `cell` of M
template <int N, int BPL>
__device__ __forceinline__ void apply_row(const MatView& M, const double (&C)[BPL][N][N], int cell, const double* s,
                                          int stride, double a, double (&out)[N]) {
    if (cell >= M.i_lo && cell < M.i_hi) {
#pragma unroll
        for (int d = 0; d < BPL; d++) {
            const double* p = s + (M.off[d] * N) * stride;
            double xv[N];
#pragma unroll
            for (int q = 0; q < N; q++) xv[q] = p[q * stride];
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) t = __fma_rn(C[d][k][q], xv[q], t);
                out[k] = __fma_rn(a, t, out[k]);
            }
        }
    } else {
#pragma unroll
        for (int d = 0; d < BPL; d++) {
            if (M.cols[cell * BPL + d] < 0) continue;
            const double* blk = M.data + (size_t)M.didx[cell * BPL + d] * N * N;
            const double* p = s + (M.off[d] * N) * stride;
            double xv[N];
#pragma unroll
            for (int q = 0; q < N; q++) xv[q] = p[q * stride];
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) t = __fma_rn(__ldg(blk + k * N + q), xv[q], t);
                out[k] = __fma_rn(a, t, out[k]);
            }
        }
    }
}

// global cell index of tile-relative cell c (may lie in the halo): wrapped if periodic, -1 if outside
__device__ __forceinline__ int gcell(int c, int num, int wrap) {
    if (c >= 0 && c < num) return c;
    if (!wrap) return -1;
    return c < 0 ? c + num : c - num;
}

template <int N, int B, bool DOT>
__global__ void __launch_bounds__(FUSED_THREADS, 2)
elliptic2d_fused_kernel(const __grid_constant__ FusedArgs A, const __grid_constant__ EllipticCoef<N, B> C) {
    constexpr int NW = FUSED_THREADS / 32;
    __shared__ long long dsm[DOT ? NW * sa::BINS : 1];
    sa::Fpe fpe;
    int bad = 0;
    if (DOT) {
        if (A.pcg->done) return;  // solver already converged: the remaining launches of the batch are no-ops
        sa::block_init<NW>(dsm);
        fpe.clear();
    }
    constexpr int H = (B == 2) ? 1 : 2;          // halo of the x tile in cells
    constexpr int XR = (TY + 2 * H) * N, XC = (TX + 2 * H) * N;  // x tile
    constexpr int SR = (TY + 2) * N, SC = (TX + 2) * N;          // sigma tile (one-cell ring)
    constexpr int TXR = TY * N, TXC = (TX + 2) * N;              // tx: tile rows, one-cell ring in x
    constexpr int TYR = (TY + 2) * N, TYC = TX * N;              // ty: one-cell ring in y, tile columns
    extern __shared__ double smem[];
    double* xs = smem;
    double* ss = xs + XR * XC;
    double* txs = ss + SR * SC;
    double* tys = txs + TXR * TXC;
    const int cx0 = blockIdx.x * TX, cy0 = blockIdx.y * TY;
    const int LDG = A.Nx * N;  // global row length
    const int tid = threadIdx.x;

    // ---- phase 0: stage x (halo H) and sigma (halo 1)
    for (int e = tid; e < XR * XC; e += FUSED_THREADS) {
        int r = e / XC, c = e - r * XC;
        int gy = gcell(cy0 - H + r / N, A.Ny, A.wrapy), gx = gcell(cx0 - H + c / N, A.Nx, A.wrapx);
        double v = 0.;
        if (gy >= 0 && gx >= 0) v = __ldg(A.x + (size_t)(gy * N + r % N) * LDG + gx * N + c % N);
        xs[e] = v;
    }
    for (int e = tid; e < SR * SC; e += FUSED_THREADS) {
        int r = e / SC, c = e - r * SC;
        int gy = gcell(cy0 - 1 + r / N, A.Ny, A.wrapy), gx = gcell(cx0 - 1 + c / N, A.Nx, A.wrapx);
        double v = 0.;
        if (gy >= 0 && gx >= 0) v = __ldg(A.sigma + (size_t)(gy * N + r % N) * LDG + gx * N + c % N);
        ss[e] = v;
    }
    __syncthreads();

    // ---- phase 1a: tx = sigma * (Rx x) on tile rows x (tile + ring) cells; item = (row r, cell c)
    for (int it = tid; it < TXR * (TX + 2); it += FUSED_THREADS) {
        int r = it / (TX + 2), c = it - r * (TX + 2);  // c = 0 is the cell left of the tile
        int gx = gcell(cx0 - 1 + c, A.Nx, A.wrapx);
        int gyc = cy0 + r / N;
        double g[N];
#pragma unroll
        for (int k = 0; k < N; k++) g[k] = 0.;
        if (gx >= 0 && gyc < A.Ny && c - 1 >= A.txlo && c - 1 <= TX - 1 + A.txhi) {
            apply_row<N, B>(A.rx, C.rx, gx, xs + (H * N + r) * XC + (H - 1 + c) * N, 1, 1., g);
            const double* sg = ss + (N + r) * SC + c * N;
#pragma unroll
            for (int k = 0; k < N; k++) g[k] = __fma_rn(sg[k], g[k], __dmul_rn(g[k], 0.));
        }
#pragma unroll
        for (int k = 0; k < N; k++) txs[r * TXC + c * N + k] = g[k];
    }
    // ---- phase 1b: ty = sigma * (Ry x) on (tile + ring) cells x tile columns; item = (cell r, column c)
    for (int it = tid; it < (TY + 2) * TYC; it += FUSED_THREADS) {
        int r = it / TYC, c = it - r * TYC;  // r = 0 is the cell below the tile
        int gy = gcell(cy0 - 1 + r, A.Ny, A.wrapy);
        int gxc = cx0 + c / N;
        double g[N];
#pragma unroll
        for (int k = 0; k < N; k++) g[k] = 0.;
        if (gy >= 0 && gxc < A.Nx && r - 1 >= A.tylo && r - 1 <= TY - 1 + A.tyhi) {
            apply_row<N, B>(A.ry, C.ry, gy, xs + ((H - 1 + r) * N) * XC + H * N + c, XC, 1., g);
            const double* sg = ss + (r * N) * SC + N + c;
#pragma unroll
            for (int k = 0; k < N; k++) g[k] = __fma_rn(sg[k * SC], g[k], __dmul_rn(g[k], 0.));
        }
#pragma unroll
        for (int k = 0; k < N; k++) tys[(r * N + k) * TYC + c] = g[k];
    }
    __syncthreads();

    // ---- phase 2: one thread per cell
    const int cx = tid % TX, cy = tid / TX;
    const int ix = cx0 + cx, iy = cy0 + cy;
    double acc[N][N];  // [ky][kx]
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < N; b++) acc[a][b] = 0.;
    const bool active = ix < A.Nx && iy < A.Ny;
    if (active) {
        // Ly ty (alpha = 1, beta = 0)
#pragma unroll
        for (int kx = 0; kx < N; kx++) {
            double col[N];
#pragma unroll
            for (int k = 0; k < N; k++) col[k] = 0.;
            apply_row<N, B>(A.ly, C.ly, iy, tys + ((cy + 1) * N) * TYC + cx * N + kx, TYC, 1., col);
#pragma unroll
            for (int k = 0; k < N; k++) acc[k][kx] = col[k];
        }
        // - Lx tx - t   (alpha = -1, beta = -1)
#pragma unroll
        for (int ky = 0; ky < N; ky++) {
            double row[N];
#pragma unroll
            for (int k = 0; k < N; k++) row[k] = __dmul_rn(acc[ky][k], -1.);
            apply_row<N, B>(A.lx, C.lx, ix, txs + (cy * N + ky) * TXC + (cx + 1) * N, 1, -1., row);
#pragma unroll
            for (int k = 0; k < N; k++) acc[ky][k] = row[k];
        }
        if (A.jfactor != 0.) {
#pragma unroll
            for (int ky = 0; ky < N; ky++) {
                double row[N];
#pragma unroll
                for (int k = 0; k < N; k++) row[k] = acc[ky][k];
                apply_row<N, 3>(A.jx, C.jx, ix, xs + ((cy + H) * N + ky) * XC + (cx + H) * N, 1, A.jfactor, row);
#pragma unroll
                for (int k = 0; k < N; k++) acc[ky][k] = row[k];
            }
#pragma unroll
            for (int kx = 0; kx < N; kx++) {
                double col[N];
#pragma unroll
                for (int k = 0; k < N; k++) col[k] = acc[k][kx];
                apply_row<N, 3>(A.jy, C.jy, iy, xs + ((cy + H) * N) * XC + (cx + H) * N + kx, XC, A.jfactor, col);
#pragma unroll
                for (int k = 0; k < N; k++) acc[k][kx] = col[k];
            }
        }
    }
    __syncthreads();  // every read of txs is done: reuse it as the output staging tile (TY*N x TX*N)
    double* outs = txs;
    constexpr int OC = TX * N;
#pragma unroll
    for (int ky = 0; ky < N; ky++)
#pragma unroll
        for (int kx = 0; kx < N; kx++) outs[(cy * N + ky) * OC + cx * N + kx] = acc[ky][kx];
    __syncthreads();
    // ---- phase 3: coalesced epilogue  y = fma(alpha, t/vol, beta*y)
    for (int e = tid; e < TY * N * OC; e += FUSED_THREADS) {
        int r = e / OC, c = e - r * OC;
        int gyc = cy0 + r / N, gxc = cx0 + c / N;
        if (gyc >= A.Ny || gxc >= A.Nx) continue;
        size_t g = (size_t)(cy0 * N + r) * LDG + cx0 * N + c;
        double t = outs[e];
        if (A.vol) t = __ddiv_rn(t, __ldg(A.vol + g));
        double b = A.beta == 0. ? 0. : __dmul_rn(A.y[g], A.beta);
        double v = __fma_rn(A.alpha, t, b);
        A.y[g] = v;
        if (DOT) {
            constexpr int Hh = (B == 2) ? 1 : 2;
            double xv = xs[(Hh * N + r) * ((TX + 2 * Hh) * N) + Hh * N + c];
            double pr = __dmul_rn(__dmul_rn(xv, __ldg(A.dot_w + g)), v);
            if (!isfinite(pr)) { bad = 1; pr = 0.; }
            fpe.add(pr, dsm + (tid >> 5) * sa::BINS);
        }
    }
    if (DOT) {
        fpe.flush(dsm + (tid >> 5) * sa::BINS);
        if (sa::block_finish<NW>(dsm, bad, A.slot, 0) && tid == 0) pcg_after_pAp(A.pcg, A.slot.result);
    }
}

static MatView view(const EllDev& m) {
    MatView v;
    v.data = m.data; v.cols = m.cols; v.didx = m.didx;
    v.i_lo = m.i_lo; v.i_hi = m.i_hi; v.num = m.num_rows;
    for (int d = 0; d < 3; d++) v.off[d] = m.off[d];
    return v;
}
template <int N, int BPL>
static void fill(double (&dst)[BPL][N][N], const EllDev& m) {
    for (int d = 0; d < BPL; d++)
        for (int k = 0; k < N; k++)
            for (int q = 0; q < N; q++) dst[d][k][q] = m.h_data[((size_t)m.did[d] * N + k) * N + q];
}

template <int N, int B, bool DOT>
static int launch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                  const FusedDot* fd) {
    constexpr int H = (B == 2) ? 1 : 2;
    constexpr size_t smem = sizeof(double) * ((size_t)(TY + 2 * H) * N * (TX + 2 * H) * N + (size_t)(TY + 2) * N * (TX + 2) * N +
                                              (size_t)TY * N * (TX + 2) * N + (size_t)(TY + 2) * N * TX * N);
    static bool configured = false;
    if (!configured) {
        DGB_CUDA(cudaFuncSetAttribute(elliptic2d_fused_kernel<N, B, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    FusedArgs A;
    A.rx = view(p.rightx); A.ry = view(p.righty); A.lx = view(p.leftx); A.ly = view(p.lefty);
    A.jx = view(p.jumpx); A.jy = view(p.jumpy);
    A.Nx = p.Nx; A.Ny = p.Ny; A.wrapx = p.wrapx; A.wrapy = p.wrapy;
    A.txlo = A.txhi = A.tylo = A.tyhi = 0;
    for (int d = 0; d < B; d++) {
        A.txlo = std::min(A.txlo, p.leftx.off[d]); A.txhi = std::max(A.txhi, p.leftx.off[d]);
        A.tylo = std::min(A.tylo, p.lefty.off[d]); A.tyhi = std::max(A.tyhi, p.lefty.off[d]);
    }
    A.sigma = p.sigma; A.vol = p.vol; A.x = x; A.y = y;
    A.alpha = alpha; A.beta = beta; A.jfactor = p.jfactor;
    A.dot_w = nullptr; A.pcg = nullptr; A.slot = sa::DotSlot{nullptr, nullptr, nullptr, nullptr};
    if (DOT) { A.dot_w = fd->w; A.slot = fd->slot; A.pcg = fd->pcg; }
    EllipticCoef<N, B> C;
    fill<N, B>(C.rx, p.rightx); fill<N, B>(C.ry, p.righty); fill<N, B>(C.lx, p.leftx); fill<N, B>(C.ly, p.lefty);
    fill<N, 3>(C.jx, p.jumpx); fill<N, 3>(C.jy, p.jumpy);
    dim3 grid((p.Nx + TX - 1) / TX, (p.Ny + TY - 1) / TY);
    elliptic2d_fused_kernel<N, B, DOT><<<grid, FUSED_THREADS, smem, st>>>(A, C);
    DGB_LAUNCHED();
    return 0;
}

template <bool DOT>
static int dispatch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                    const FusedDot* fd) {
    switch (p.n * 10 + p.bder) {
        case 22: return launch<2, 2, DOT>(p, alpha, x, beta, y, st, fd);
        case 23: return launch<2, 3, DOT>(p, alpha, x, beta, y, st, fd);
        case 32: return launch<3, 2, DOT>(p, alpha, x, beta, y, st, fd);
        case 33: return launch<3, 3, DOT>(p, alpha, x, beta, y, st, fd);
        case 42: return launch<4, 2, DOT>(p, alpha, x, beta, y, st, fd);
        case 43: return launch<4, 3, DOT>(p, alpha, x, beta, y, st, fd);
    }
    set_error("elliptic2d fused kernel: unsupported n=%d bpl=%d", p.n, p.bder);
    return DGB_ERR_UNSUPPORTED;
}
int elliptic2d_fused_launch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    return dispatch<false>(p, alpha, x, beta, y, st, nullptr);
}
// y = A x fused with dot(x, w, y) and the PCG alpha update (pcg.h:165-166)
int elliptic2d_fused_launch_dot(Elliptic2dPlan& p, const double* x, double* y, cudaStream_t st, const FusedDot& fd) {
    return dispatch<true>(p, 1., x, 0., y, st, &fd);
}

}  // namespace dgb
