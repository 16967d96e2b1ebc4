// dg::Advection::upwind (inc/dg/advection.h:112-120) in ONE pass: the reference applies four block-ELL derivatives
// (backward / forward in x and y) into two temporaries and combines them with two `evaluate( Axpby, UpwindProduct)` sweeps --
// six launches and 128 B/dof of traffic.  Here a thread owns one cell: it reads the cell's n x n values of f and those of its
// four neighbours (the neighbours' lines come from L1 / L2: adjacent threads read adjacent cells), forms the four derivatives
// of every node with the reference's rounding sequence (per block one FMA chain over q, blocks added in slot order with
// fma(1, t, y): sparseblockmat_omp_kernels.h:36-50; interior blocks are constant-bank operands) and applies the two upwind
// updates (functors.h:312-337, subroutines.h Axpby): 32 B/dof (f, vx, vy, result) + 8 if beta != 0.
// Cells in boundary block rows of any matrix take a general path that walks the matrix row (column index -> neighbour cell).
// Results are bitwise those of the six-launch composition (tests/test_gpu_core.py::test_advection_upwind_fused).
#include "ell.cuh"

namespace dgb {

template <int N>
struct UpwindCoef {
    double xb[2][N][N], xf[2][N][N], yb[2][N][N], yf[2][N][N];
};
struct UpwindArgs {
    EllArgs xb, xf, yb, yf;
    int Nx, Ny;
    int fx_lo, fx_hi, fy_lo, fy_hi;  // cells that are interior rows of both x- resp. both y-matrices
    double alpha, beta;
    const double* vx;
    const double* vy;
    const double* f;
    double* result;
};

// one derivative of the cell (cx, cy) along x (ALONGX) or y through the matrix row of a boundary cell: out[ky][kx]
template <int N, bool ALONGX>
__device__ __forceinline__ void upwind_general(const EllArgs& M, const double* __restrict__ f, int cx, int cy, int LD, double (&out)[N][N]) {
    const int row = ALONGX ? cx : cy;
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < N; b++) out[a][b] = 0.;
    for (int d = 0; d < M.bpl; d++) {
        const int J = M.cols[row * M.bpl + d];
        if (J < 0) continue;
        const double* blk = M.data + (size_t)M.didx[row * M.bpl + d] * N * N;
#pragma unroll
        for (int line = 0; line < N; line++)  // the other index: ky for x-derivatives, kx for y-derivatives
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) {
                    const double xv = ALONGX ? f[(size_t)(cy * N + line) * LD + J * N + q] : f[(size_t)(J * N + q) * LD + cx * N + line];
                    t = __fma_rn(__ldg(blk + k * N + q), xv, t);
                }
                if (ALONGX) out[line][k] = __fma_rn(1., t, out[line][k]);
                else out[k][line] = __fma_rn(1., t, out[k][line]);
            }
    }
}

template <int N>
__global__ void __launch_bounds__(128)
advection_upwind_kernel(const __grid_constant__ UpwindArgs A, const __grid_constant__ UpwindCoef<N> C) {
    const int LD = A.Nx * N;
    const long long ncells = (long long)A.Nx * A.Ny;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const int cy = (int)(c / A.Nx), cx = (int)(c - (long long)cy * A.Nx);
        const double* fc = A.f + (size_t)(cy * N) * LD + cx * N;
        double F0[N][N];
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) F0[a][b] = fc[(size_t)a * LD + b];
        double db[N][N], df[N][N];  // backward / forward derivative of the direction at hand
        double res[N][N];
        // ---------------------------------------------------------------- x
        if (cx >= A.fx_lo && cx < A.fx_hi) {
            double FL[N][N], FR[N][N];
#pragma unroll
            for (int a = 0; a < N; a++)
#pragma unroll
                for (int b = 0; b < N; b++) { FL[a][b] = fc[(size_t)a * LD + b - N]; FR[a][b] = fc[(size_t)a * LD + b + N]; }
#pragma unroll
            for (int ky = 0; ky < N; ky++)
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double t0 = 0., t1 = 0., u0 = 0., u1 = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        t0 = __fma_rn(C.xb[0][k][q], FL[ky][q], t0);   // backward: blocks at cells i-1, i
                        t1 = __fma_rn(C.xb[1][k][q], F0[ky][q], t1);
                        u0 = __fma_rn(C.xf[0][k][q], F0[ky][q], u0);   // forward: blocks at cells i, i+1
                        u1 = __fma_rn(C.xf[1][k][q], FR[ky][q], u1);
                    }
                    db[ky][k] = __fma_rn(1., t1, __fma_rn(1., t0, 0.));
                    df[ky][k] = __fma_rn(1., u1, __fma_rn(1., u0, 0.));
                }
        } else {
            upwind_general<N, true>(A.xb, A.f, cx, cy, LD, db);
            upwind_general<N, true>(A.xf, A.f, cx, cy, LD, df);
        }
        const size_t g0 = (size_t)(cy * N) * LD + cx * N;
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) {
                const size_t g = g0 + (size_t)a * LD + b;
                const double v = __ldg(A.vx + g);
                const double up = __dmul_rn(v, v >= 0. ? db[a][b] : df[a][b]);
                res[a][b] = __fma_rn(A.alpha, up, __dmul_rn(A.result[g], A.beta));
            }
        // ---------------------------------------------------------------- y
        if (cy >= A.fy_lo && cy < A.fy_hi) {
            double FD[N][N], FU[N][N];
#pragma unroll
            for (int a = 0; a < N; a++)
#pragma unroll
                for (int b = 0; b < N; b++) { FD[a][b] = fc[((long long)a - N) * LD + b]; FU[a][b] = fc[(size_t)(a + N) * LD + b]; }
#pragma unroll
            for (int kx = 0; kx < N; kx++)
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double t0 = 0., t1 = 0., u0 = 0., u1 = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) {
                        t0 = __fma_rn(C.yb[0][k][q], FD[q][kx], t0);
                        t1 = __fma_rn(C.yb[1][k][q], F0[q][kx], t1);
                        u0 = __fma_rn(C.yf[0][k][q], F0[q][kx], u0);
                        u1 = __fma_rn(C.yf[1][k][q], FU[q][kx], u1);
                    }
                    db[k][kx] = __fma_rn(1., t1, __fma_rn(1., t0, 0.));
                    df[k][kx] = __fma_rn(1., u1, __fma_rn(1., u0, 0.));
                }
        } else {
            upwind_general<N, false>(A.yb, A.f, cx, cy, LD, db);
            upwind_general<N, false>(A.yf, A.f, cx, cy, LD, df);
        }
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) {
                const size_t g = g0 + (size_t)a * LD + b;
                const double v = __ldg(A.vy + g);
                const double up = __dmul_rn(v, v >= 0. ? db[a][b] : df[a][b]);
                A.result[g] = __fma_rn(A.alpha, up, __dmul_rn(res[a][b], 1.));
            }
    }
}

// the two-block structure the fast path assumes: backward = cells (i-1, i), forward = (i, i+1) on interior rows
static bool two_block(const EllDev& m, int o0, int o1) { return m.bpl == 2 && m.has_pattern && m.off[0] == o0 && m.off[1] == o1; }

template <int N>
static int upwind_launch(const EllDev& xb, const EllDev& xf, const EllDev& yb, const EllDev& yf, UpwindArgs& A, cudaStream_t st) {
    UpwindCoef<N> C;
    const EllDev* ms[4] = {&xb, &xf, &yb, &yf};
    double (*dst[4])[N][N] = {C.xb, C.xf, C.yb, C.yf};
    for (int m = 0; m < 4; m++)
        for (int d = 0; d < 2; d++)
            for (int k = 0; k < N; k++)
                for (int q = 0; q < N; q++) dst[m][d][k][q] = ms[m]->h_data[((size_t)ms[m]->did[d] * N + k) * N + q];
    const long long ncells = (long long)A.Nx * A.Ny;
    long long want = (ncells + 127) / 128, cap = (long long)sm_count() * 16;
    advection_upwind_kernel<N><<<(unsigned)std::max(1ll, std::min(want, cap)), 128, 0, st>>>(A, C);
    DGB_LAUNCHED();
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" int dgb_advection_upwind(const dgb_ell* dxb, const dgb_ell* dxf, const dgb_ell* dyb, const dgb_ell* dyf, double alpha,
                                    const double* vx, const double* vy, const double* f, double beta, double* result, dgb_stream_t s) {
    const EllDev* xb = reinterpret_cast<const EllDev*>(dxb);
    const EllDev* xf = reinterpret_cast<const EllDev*>(dxf);
    const EllDev* yb = reinterpret_cast<const EllDev*>(dyb);
    const EllDev* yf = reinterpret_cast<const EllDev*>(dyf);
    if (!xb || !xf || !yb || !yf || !vx || !vy || !f || !result) { set_error("dgb_advection_upwind: NULL argument"); return DGB_ERR_INVALID; }
    if (f == result) { set_error("dgb_advection_upwind: f must not alias result"); return DGB_ERR_INVALID; }
    const int n = xb->n, Nx = xb->num_rows, Ny = yb->num_rows;
    const bool shapes = xf->n == n && yb->n == n && yf->n == n && xf->num_rows == Nx && yf->num_rows == Ny &&
                        xb->num_cols == Nx && xf->num_cols == Nx && yb->num_cols == Ny && yf->num_cols == Ny &&
                        xb->right == 1 && xf->right == 1 && xb->left == Ny * n && xf->left == Ny * n &&
                        yb->left == 1 && yf->left == 1 && yb->right == Nx * n && yf->right == Nx * n;
    if (!shapes || n < 2 || n > 4 || !two_block(*xb, -1, 0) || !two_block(*xf, 0, 1) || !two_block(*yb, -1, 0) || !two_block(*yf, 0, 1)) {
        set_error("dgb_advection_upwind: the matrices are not the backward / forward derivatives of one 2-d grid (n = 2..4)");
        return DGB_ERR_UNSUPPORTED;
    }
    UpwindArgs A;
    A.xb = ell_args(*xb); A.xf = ell_args(*xf); A.yb = ell_args(*yb); A.yf = ell_args(*yf);
    A.Nx = Nx; A.Ny = Ny;
    A.fx_lo = std::max(std::max(xb->i_lo, xf->i_lo), 1); A.fx_hi = std::min(std::min(xb->i_hi, xf->i_hi), Nx - 1);
    A.fy_lo = std::max(std::max(yb->i_lo, yf->i_lo), 1); A.fy_hi = std::min(std::min(yb->i_hi, yf->i_hi), Ny - 1);
    A.alpha = alpha; A.beta = beta; A.vx = vx; A.vy = vy; A.f = f; A.result = result;
    cudaStream_t st = as_stream(s);
    switch (n) {
        case 2: return upwind_launch<2>(*xb, *xf, *yb, *yf, A, st);
        case 3: return upwind_launch<3>(*xb, *xf, *yb, *yf, A, st);
        default: return upwind_launch<4>(*xb, *xf, *yb, *yf, A, st);
    }
}
