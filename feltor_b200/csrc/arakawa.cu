// dg::ArakawaX::operator() (inc/dg/arakawa.h:147-162) in TWO passes instead of eight launches.  The reference applies the
// centered derivatives to lhs and rhs (four block-ELL symv), mixes them pointwise (ArakawaFunctor, arakawa.h:125-145),
// applies dx / dy once more to two of the mixed fields (two symv with beta = 1) and scales by chi (pointwiseDot): 30 vector
// passes (its own "memops: 30").  Here
//   pass 1  a thread owns one cell: reads the cell's n x n values of lhs and rhs and those of the four neighbours (adjacent
//           threads read adjacent cells: L1 / L2 serve the re-reads), forms the four derivatives of every node and the three
//           mixed fields in registers and writes those: 16 B/dof read, 24 written;
//   pass 2  a thread owns one cell: dx of the first mixed field, dy of the second, added to the third in the reference's
//           order (y *= 1; y = fma(1, block sum, y) per block slot), then result = fma(alpha chi, y, beta result):
//           40 B/dof (+8 when chi is given... it always is).
// Every block product is the reference's FMA chain over q, blocks are added in slot order
// (sparseblockmat_omp_kernels.h:36-50); interior blocks are constant-bank operands, cells in boundary block rows walk the
// matrix row.  Results are bitwise those of the eight-launch composition (tests/test_gpu_toefl.py::test_arakawa_fused).
#include "ell.cuh"

namespace dgb {

template <int N>
struct BracketCoef {
    double x[3][N][N], y[3][N][N];
};
struct BracketArgs {
    EllArgs dx, dy;
    int Nx, Ny;
    int fx_lo, fx_hi, fy_lo, fy_hi;  // cells that are interior rows of dx resp. dy with both neighbours present
    double alpha, beta;
    const double* lhs;
    const double* rhs;
    const double* chi;
    double* m1;  // dylhs after the functor (arakawa.h:139)
    double* m2;  // dxrhs after the functor (arakawa.h:143)
    double* m3;  // dyrhs after the functor (arakawa.h:137)
    double* result;
};

// out += M f restricted to the cell (cx, cy): derivative along x (ALONGX) or y through the matrix row of the cell
template <int N, bool ALONGX>
__device__ __forceinline__ void bracket_general(const EllArgs& M, const double* __restrict__ f, int cx, int cy, int LD, double (&out)[N][N]) {
    const int row = ALONGX ? cx : cy;
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

// out += (interior centered derivative of f at the cell whose first value is fc), three blocks at cells -1, 0, +1
template <int N, bool ALONGX>
__device__ __forceinline__ void bracket_interior(const double (&C)[3][N][N], const double* __restrict__ fc, int LD, double (&out)[N][N]) {
#pragma unroll
    for (int d = 0; d < 3; d++) {
        double F[N][N];
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++)
                F[a][b] = ALONGX ? fc[(long long)a * LD + b + (d - 1) * N] : fc[((long long)a + (d - 1) * N) * LD + b];
#pragma unroll
        for (int line = 0; line < N; line++)
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) t = __fma_rn(C[d][k][q], ALONGX ? F[line][q] : F[q][line], t);
                if (ALONGX) out[line][k] = __fma_rn(1., t, out[line][k]);
                else out[k][line] = __fma_rn(1., t, out[k][line]);
            }
    }
}

template <int N, bool ALONGX>
__device__ __forceinline__ void bracket_deriv(const BracketArgs& A, const BracketCoef<N>& C, const double* __restrict__ f, int cx, int cy, int LD,
                                               double (&out)[N][N]) {
    const bool fast = ALONGX ? (cx >= A.fx_lo && cx < A.fx_hi) : (cy >= A.fy_lo && cy < A.fy_hi);
    if (fast) bracket_interior<N, ALONGX>(ALONGX ? C.x : C.y, f + (size_t)(cy * N) * LD + cx * N, LD, out);
    else bracket_general<N, ALONGX>(ALONGX ? A.dx : A.dy, f, cx, cy, LD, out);
}

template <int N>
__device__ __forceinline__ void zero(double (&o)[N][N]) {
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < N; b++) o[a][b] = 0.;
}

template <int N>
__global__ void __launch_bounds__(128)
arakawa_mix_kernel(const __grid_constant__ BracketArgs A, const __grid_constant__ BracketCoef<N> C) {
    const int LD = A.Nx * N;
    const long long ncells = (long long)A.Nx * A.Ny;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const int cy = (int)(c / A.Nx), cx = (int)(c - (long long)cy * A.Nx);
        double dxl[N][N], dyl[N][N], dxr[N][N], dyr[N][N];
        zero<N>(dxl); zero<N>(dyl); zero<N>(dxr); zero<N>(dyr);
        bracket_deriv<N, true>(A, C, A.lhs, cx, cy, LD, dxl);
        bracket_deriv<N, false>(A, C, A.lhs, cx, cy, LD, dyl);
        bracket_deriv<N, true>(A, C, A.rhs, cx, cy, LD, dxr);
        bracket_deriv<N, false>(A, C, A.rhs, cx, cy, LD, dyr);
        const size_t g0 = (size_t)(cy * N) * LD + cx * N;
        const double third = 1. / 3., mthird = -(1. / 3.);
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) {
                const size_t g = g0 + (size_t)a * LD + b;
                const double lhs = A.lhs[g], rhs = A.rhs[g];
                double result = 0.;
                result = __fma_rn(__dmul_rn(third, dxl[a][b]), dyr[a][b], result);
                result = __fma_rn(__dmul_rn(mthird, dyl[a][b]), dxr[a][b], result);
                double temp = 0.;
                temp = __fma_rn(__dmul_rn(third, lhs), dyr[a][b], temp);
                temp = __fma_rn(__dmul_rn(mthird, dyl[a][b]), rhs, temp);
                A.m3[g] = result;
                A.m1[g] = temp;
                temp = 0.;
                temp = __fma_rn(__dmul_rn(third, dxl[a][b]), rhs, temp);
                temp = __fma_rn(__dmul_rn(mthird, lhs), dxr[a][b], temp);
                A.m2[g] = temp;
            }
    }
}

template <int N>
__global__ void __launch_bounds__(128)
arakawa_close_kernel(const __grid_constant__ BracketArgs A, const __grid_constant__ BracketCoef<N> C) {
    const int LD = A.Nx * N;
    const long long ncells = (long long)A.Nx * A.Ny;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const int cy = (int)(c / A.Nx), cx = (int)(c - (long long)cy * A.Nx);
        const size_t g0 = (size_t)(cy * N) * LD + cx * N;
        double acc[N][N];
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) acc[a][b] = __dmul_rn(A.m3[g0 + (size_t)a * LD + b], 1.);  // symv( 1., dx, m1, 1., m3): y *= beta
        bracket_deriv<N, true>(A, C, A.m1, cx, cy, LD, acc);
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) acc[a][b] = __dmul_rn(acc[a][b], 1.);                      // symv( 1., dy, m2, 1., m3)
        bracket_deriv<N, false>(A, C, A.m2, cx, cy, LD, acc);
#pragma unroll
        for (int a = 0; a < N; a++)
#pragma unroll
            for (int b = 0; b < N; b++) {
                const size_t g = g0 + (size_t)a * LD + b;
                A.result[g] = __fma_rn(__dmul_rn(A.alpha, __ldg(A.chi + g)), acc[a][b], __dmul_rn(A.result[g], A.beta));  // subroutines.h:313
            }
    }
}

template <int N>
static int bracket_launch(const EllDev& dx, const EllDev& dy, BracketArgs& A, cudaStream_t st) {
    BracketCoef<N> C;
    for (int d = 0; d < 3; d++)
        for (int k = 0; k < N; k++)
            for (int q = 0; q < N; q++) {
                C.x[d][k][q] = dx.h_data[((size_t)dx.did[d] * N + k) * N + q];
                C.y[d][k][q] = dy.h_data[((size_t)dy.did[d] * N + k) * N + q];
            }
    const long long ncells = (long long)A.Nx * A.Ny;
    long long want = (ncells + 127) / 128, cap = (long long)sm_count() * 16;
    const unsigned grid = (unsigned)std::max(1ll, std::min(want, cap));
    arakawa_mix_kernel<N><<<grid, 128, 0, st>>>(A, C);
    DGB_LAUNCHED();
    arakawa_close_kernel<N><<<grid, 128, 0, st>>>(A, C);
    DGB_LAUNCHED();
    return 0;
}

static bool three_block(const EllDev& m) { return m.bpl == 3 && m.has_pattern && m.off[0] == -1 && m.off[1] == 0 && m.off[2] == 1; }

}  // namespace dgb

using namespace dgb;

extern "C" int dgb_arakawa(const dgb_ell* bdx, const dgb_ell* bdy, double alpha, const double* lhs, const double* rhs, const double* chi,
                           double beta, double* result, double* work3, dgb_stream_t s) {
    const EllDev* dx = reinterpret_cast<const EllDev*>(bdx);
    const EllDev* dy = reinterpret_cast<const EllDev*>(bdy);
    if (!dx || !dy || !lhs || !rhs || !chi || !result || !work3) { set_error("dgb_arakawa: NULL argument"); return DGB_ERR_INVALID; }
    const int n = dx->n, Nx = dx->num_rows, Ny = dy->num_rows;
    const bool shapes = dy->n == n && dx->num_cols == Nx && dy->num_cols == Ny && dx->right == 1 && dx->left == Ny * n && dy->left == 1 &&
                        dy->right == Nx * n;
    if (!shapes || n < 2 || n > 4 || !three_block(*dx) || !three_block(*dy)) {
        set_error("dgb_arakawa: the matrices are not the centered derivatives of one 2-d grid (n = 2..4)");
        return DGB_ERR_UNSUPPORTED;
    }
    const size_t size = (size_t)Nx * Ny * n * n;
    BracketArgs A;
    A.dx = ell_args(*dx); A.dy = ell_args(*dy);
    A.Nx = Nx; A.Ny = Ny;
    A.fx_lo = std::max(dx->i_lo, 1); A.fx_hi = std::min(dx->i_hi, Nx - 1);
    A.fy_lo = std::max(dy->i_lo, 1); A.fy_hi = std::min(dy->i_hi, Ny - 1);
    A.alpha = alpha; A.beta = beta; A.lhs = lhs; A.rhs = rhs; A.chi = chi;
    A.m1 = work3; A.m2 = work3 + size; A.m3 = work3 + 2 * size; A.result = result;
    if (result == work3 || result == A.m2 || result == A.m3) { set_error("dgb_arakawa: result must not alias the work space"); return DGB_ERR_INVALID; }
    cudaStream_t st = as_stream(s);
    switch (n) {
        case 2: return bracket_launch<2>(*dx, *dy, A, st);
        case 3: return bracket_launch<3>(*dx, *dy, A, st);
        default: return bracket_launch<4>(*dx, *dy, A, st);
    }
}
