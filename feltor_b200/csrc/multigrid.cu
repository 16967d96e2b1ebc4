// dg::NestedGrids / dg::nested_iterations / dg::MultigridCG2d::solve (inc/dg/multigrid.h:28-171,197-245,500-668)
// on top of the Elliptic2d plans and the device-resident PCG.
//  * stage u+1 has half the cells of stage u in x and y (multigrid.h:55-59); projection(u) = fast_projection(grid(u),1,2,2),
//    interpolation(u) = fast_interpolation(grid(u+1),1,2,2) (multigrid.h:64-71), each a MultiMatrix X-then-Y product of two
//    block-ELL matrices through a temporary (fast_interpolation.h:71-84,380-398);
//  * nested iteration = FAS: residual, restrict r and x, coarse right hand sides A x + r, solve coarse -> fine with PCG
//    (test_frequency 10 on the coarse stages, 1 on the fine one, multigrid.h:640-648), prolong the corrections.
#include "elliptic.cuh"
#include "pcg_internal.cuh"
#include "topology.h"
#include <cmath>
#include <vector>

namespace dgb {

extern "C" int dgb_axpby(size_t, double, const double*, double, double*, dgb_stream_t);
extern "C" int dgb_axpbyz(size_t, double, const double*, double, const double*, double*, dgb_stream_t);
extern "C" int dgb_copy(size_t, const double*, double*, dgb_stream_t);

struct MultiMat {  // Y o X with a temporary (MultiMatrix of dimension 2)
    EllDev mx, my;
    double* temp = nullptr;
    size_t temp_size = 0;
    int fused = 0;  // 1: factor-2 projection, 2: factor-2 interpolation (one-pass kernels below), 0: two Ell symv through temp
    int Nxc = 0, Nyc = 0;
};

struct Multigrid2d {
    int stages = 0;
    std::vector<dgb_grid> grids;
    std::vector<size_t> sizes;
    std::vector<MultiMat> project, inter;      // stages-1 each
    std::vector<double*> x, r, b, w;           // per stage work vectors (multigrid.h:72-75)
    std::vector<Pcg*> pcg;
};

static int upload(EllDev& d, const EllHost& h) {
    dgb_ell_host v;
    ell_view(h, &v);
    return ell_upload(d, &v);
}
static void analyse(MultiMat& M, bool projection);
static int build_multimat(MultiMat& M, const dgb_grid& g, bool projection) {
    // X first on grid g, then Y on the grid whose x axis already has the new resolution (fast_interpolation.h:380-398)
    EllHost hx, hy;
    int e;
    dgb_grid gx = g;
    if (projection) {
        if ((e = topo_fast_projection1d(hx, g.n[0], g.N[0], 1, 2))) return e;
        gx.N[0] = g.N[0] / 2;
        if ((e = topo_fast_projection1d(hy, g.n[1], g.N[1], 1, 2))) return e;
    } else {
        if ((e = topo_fast_interpolation1d(hx, g.n[0], g.N[0], 1, 2))) return e;
        gx.N[0] = g.N[0] * 2;
        if ((e = topo_fast_interpolation1d(hy, g.n[1], g.N[1], 1, 2))) return e;
    }
    update_left_right(hx, &g, 0);
    update_left_right(hy, &gx, 1);
    if ((e = upload(M.mx, hx))) return e;
    if ((e = upload(M.my, hy))) return e;
    M.temp_size = M.mx.total_rows();
    DGB_CUDA(cudaMalloc(&M.temp, M.temp_size * sizeof(double)));
    analyse(M, projection);
    return 0;
}
// ---- MultiMatrix::symv (fast_interpolation.h:71-84) of the factor-2 projection / interpolation in ONE pass ----------------
// The reference applies the x-matrix into a temporary and the y-matrix from there: 4 vector passes over the fine grid.  A thread
// here owns one COARSE cell: it reads the 2 x 2 fine cells (projection) resp. its own n x n values (interpolation), forms the
// rows of the temporary it needs in registers and applies the y-blocks to them -- 10 B per fine element (8 + 2), the
// algorithmic minimum.  Arithmetic per output element is the reference's: t = FMA chain over q per block, out = fma(1, t, 0) for
// the x-matrix (alpha = 1, beta = 0), out = fma(alpha, t, beta == 0 ? 0 : y beta) per block in slot order for the y-matrix
// (sparseblockmat_omp_kernels.h:36-50): bitwise the two-pass result (tests/test_gpu_multigrid.py::test_multimatrix_fused).
template <int N>
struct HalfCoef {
    double x[2][N][N], y[2][N][N];
};
template <int N>
__global__ void __launch_bounds__(128)
project_half_kernel(const __grid_constant__ HalfCoef<N> C, int Nxc, int Nyc, double alpha, double beta, const double* __restrict__ x,
                    double* __restrict__ y) {
    const size_t LDf = (size_t)2 * Nxc * N, LDc = (size_t)Nxc * N;
    const long long ncells = (long long)Nxc * Nyc;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const int cy = (int)(c / Nxc), cx = (int)(c - (long long)cy * Nxc);
        double T[2 * N][N];
#pragma unroll
        for (int r = 0; r < 2 * N; r++) {
            const double* xr = x + ((size_t)2 * cy * N + r) * LDf + (size_t)2 * cx * N;
            double v[2 * N];
#pragma unroll
            for (int q = 0; q < 2 * N; q++) v[q] = xr[q];
#pragma unroll
            for (int k = 0; k < N; k++) {
                double out = 0.;
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    double t = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) t = __fma_rn(C.x[d][k][q], v[d * N + q], t);
                    out = __fma_rn(1., t, out);
                }
                T[r][k] = out;
            }
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
            double* yr = y + ((size_t)cy * N + k) * LDc + (size_t)cx * N;
#pragma unroll
            for (int col = 0; col < N; col++) {
                double out = beta == 0. ? 0. : __dmul_rn(yr[col], beta);
#pragma unroll
                for (int d = 0; d < 2; d++) {
                    double t = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) t = __fma_rn(C.y[d][k][q], T[d * N + q][col], t);
                    out = __fma_rn(alpha, t, out);
                }
                yr[col] = out;
            }
        }
    }
}
template <int N>
__global__ void __launch_bounds__(128)
interpolate_double_kernel(const __grid_constant__ HalfCoef<N> C, int Nxc, int Nyc, double alpha, double beta, const double* __restrict__ x,
                          double* __restrict__ y) {
    const size_t LDf = (size_t)2 * Nxc * N, LDc = (size_t)Nxc * N;
    const long long ncells = (long long)Nxc * Nyc;
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncells; c += (long long)gridDim.x * blockDim.x) {
        const int cy = (int)(c / Nxc), cx = (int)(c - (long long)cy * Nxc);
        double T[N][2 * N];  // the temporary: coarse rows of the cell, fine columns
#pragma unroll
        for (int r = 0; r < N; r++) {
            const double* xr = x + ((size_t)cy * N + r) * LDc + (size_t)cx * N;
            double v[N];
#pragma unroll
            for (int q = 0; q < N; q++) v[q] = xr[q];
#pragma unroll
            for (int p = 0; p < 2; p++)
#pragma unroll
                for (int k = 0; k < N; k++) {
                    double t = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) t = __fma_rn(C.x[p][k][q], v[q], t);
                    T[r][p * N + k] = __fma_rn(1., t, 0.);
                }
        }
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int k = 0; k < N; k++) {
                double* yr = y + ((size_t)(2 * cy + p) * N + k) * LDf + (size_t)2 * cx * N;
#pragma unroll
                for (int col = 0; col < 2 * N; col++) {
                    double out = beta == 0. ? 0. : __dmul_rn(yr[col], beta);
                    double t = 0.;
#pragma unroll
                    for (int q = 0; q < N; q++) t = __fma_rn(C.y[p][k][q], T[q][col], t);
                    yr[col] = __fma_rn(alpha, t, out);
                }
            }
    }
}

// does the pair have the structure the one-pass kernels assume?  projection: rows i, two blocks at columns 2i, 2i+1 with data
// indices (a, b) for every row; interpolation: rows i, one block at column i/2 with data index alternating with the parity of i
static bool half_pattern(const EllDev& m, bool projection) {
    if (projection) {
        if (m.bpl != 2 || m.num_cols != 2 * m.num_rows) return false;
        for (int i = 0; i < m.num_rows; i++)
            for (int d = 0; d < 2; d++)
                if (m.h_cols[(size_t)i * 2 + d] != 2 * i + d || m.h_didx[(size_t)i * 2 + d] != m.h_didx[d]) return false;
    } else {
        if (m.bpl != 1 || m.num_rows != 2 * m.num_cols) return false;
        for (int i = 0; i < m.num_rows; i++)
            if (m.h_cols[i] != i / 2 || m.h_didx[i] != m.h_didx[i % 2]) return false;
    }
    return m.rr0 == 0 && m.rr1 == m.right;
}
static void analyse(MultiMat& M, bool projection) {
    const EllDev &X = M.mx, &Y = M.my;
    M.fused = 0;
    if (X.n != Y.n || X.n < 2 || X.n > 4 || !half_pattern(X, projection) || !half_pattern(Y, projection)) return;
    const int n = X.n;
    const int Nxc = projection ? X.num_rows : X.num_cols, Nyc = projection ? Y.num_rows : Y.num_cols;
    const int rows_in = (projection ? 2 : 1) * Nyc * n, cols_out = (projection ? 1 : 2) * Nxc * n;
    if (X.right != 1 || X.left != rows_in || Y.left != 1 || Y.right != cols_out) return;
    M.fused = projection ? 1 : 2;
    M.Nxc = Nxc; M.Nyc = Nyc;
}
// kind 1: factor-2 projection, 2: factor-2 interpolation; X / Y: the x- and y-matrix (their first two block slots name the blocks)
template <int N>
static int half_launch(const EllDev& X, const EllDev& Y, int kind, int Nxc, int Nyc, double alpha, double beta, const double* x, double* y,
                       cudaStream_t st) {
    HalfCoef<N> C;
    for (int d = 0; d < 2; d++)
        for (int k = 0; k < N; k++)
            for (int q = 0; q < N; q++) {
                C.x[d][k][q] = X.h_data[((size_t)X.h_didx[d] * N + k) * N + q];
                C.y[d][k][q] = Y.h_data[((size_t)Y.h_didx[d] * N + k) * N + q];
            }
    const long long ncells = (long long)Nxc * Nyc;
    long long want = (ncells + 127) / 128, cap = (long long)sm_count() * 16;
    const unsigned grid = (unsigned)std::max(1ll, std::min(want, cap));
    if (kind == 1) project_half_kernel<N><<<grid, 128, 0, st>>>(C, Nxc, Nyc, alpha, beta, x, y);
    else interpolate_double_kernel<N><<<grid, 128, 0, st>>>(C, Nxc, Nyc, alpha, beta, x, y);
    DGB_LAUNCHED();
    return 0;
}
static int half_dispatch(const EllDev& X, const EllDev& Y, int kind, int Nxc, int Nyc, double alpha, double beta, const double* x, double* y,
                         cudaStream_t st) {
    switch (X.n) {
        case 2: return half_launch<2>(X, Y, kind, Nxc, Nyc, alpha, beta, x, y, st);
        case 3: return half_launch<3>(X, Y, kind, Nxc, Nyc, alpha, beta, x, y, st);
        default: return half_launch<4>(X, Y, kind, Nxc, Nyc, alpha, beta, x, y, st);
    }
}
static int g_multimat_two_pass = 0;  // A/B switch of the tests (dgb_multigrid2d_set_two_pass)
static int multimat_symv(const MultiMat& M, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    if (M.fused && !g_multimat_two_pass && x != y) return half_dispatch(M.mx, M.my, M.fused, M.Nxc, M.Nyc, alpha, beta, x, y, st);
    int e;
    if ((e = ell_symv(M.mx, 1., x, 0., M.temp, st, false))) return e;
    return ell_symv(M.my, alpha, M.temp, beta, y, st, false);
}

static void destroy(Multigrid2d* m) {
    if (!m) return;
    for (auto& M : m->project) { ell_release(M.mx); ell_release(M.my); cudaFree(M.temp); }
    for (auto& M : m->inter) { ell_release(M.mx); ell_release(M.my); cudaFree(M.temp); }
    for (auto v : {&m->x, &m->r, &m->b, &m->w})
        for (double* p : *v) cudaFree(p);
    for (Pcg* p : m->pcg) pcg_delete(p);
    delete m;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_multigrid2d_create(dgb_multigrid2d** out, const dgb_grid* grid, int stages) {
    if (!grid || grid->ndim != 2) { set_error("dgb_multigrid2d_create: a 2-d grid is required"); return DGB_ERR_INVALID; }
    if (stages < 1) { set_error("There must be minimum 1 stage in nested Grids construction! You gave %d", stages); return DGB_ERR_INVALID; }
    Multigrid2d* m = new Multigrid2d();
    m->stages = stages;
    m->grids.resize(stages);
    m->grids[0] = *grid;
    for (int u = 1; u < stages; u++) {  // multiplyCellNumbers(0.5, 0.5), grid.h:404-410
        m->grids[u] = m->grids[u - 1];
        m->grids[u].N[0] = (int)std::round(0.5 * (double)m->grids[u - 1].N[0]);
        m->grids[u].N[1] = (int)std::round(0.5 * (double)m->grids[u - 1].N[1]);
    }
    m->project.resize(stages - 1);
    m->inter.resize(stages - 1);
    int e = 0;
    for (int u = 0; u < stages - 1 && !e; u++) {
        e = build_multimat(m->project[u], m->grids[u], true);
        if (!e) e = build_multimat(m->inter[u], m->grids[u + 1], false);
    }
    for (int u = 0; u < stages && !e; u++) {
        size_t n = grid_size(&m->grids[u]);
        m->sizes.push_back(n);
        for (auto v : {&m->x, &m->r, &m->b, &m->w}) {
            double* p = nullptr;
            if (!e && cudaMalloc(&p, n * sizeof(double)) != cudaSuccess) { set_error("dgb_multigrid2d_create: out of device memory"); e = DGB_ERR_INVALID; }
            if (p) cudaMemset(p, 0, n * sizeof(double));
            v->push_back(p);
        }
        if (!e) {
            Pcg* p = pcg_new(n, &e);
            m->pcg.push_back(p);
        }
    }
    if (e) { destroy(m); return e; }
    *out = reinterpret_cast<dgb_multigrid2d*>(m);
    return 0;
}
int dgb_multigrid2d_destroy(dgb_multigrid2d* h) { destroy(reinterpret_cast<Multigrid2d*>(h)); return 0; }
// dg::MultiMatrix::symv of dimension 2 for any two block matrices (fast_interpolation.h:71-84): the one-pass kernels when the
// pair is a factor-2 projection / interpolation, else the reference's two products through `temp` (may be NULL in the former case
// only when the caller has checked with dgb_multimatrix2_fused)
int dgb_multimatrix2_fused(const dgb_ell* hx, const dgb_ell* hy, int* kind) {
    if (!hx || !hy || !kind) { set_error("dgb_multimatrix2_fused: NULL argument"); return DGB_ERR_INVALID; }
    MultiMat M;
    M.mx = *reinterpret_cast<const EllDev*>(hx);
    M.my = *reinterpret_cast<const EllDev*>(hy);
    analyse(M, true);
    if (!M.fused) analyse(M, false);
    *kind = M.fused;
    return 0;
}
int dgb_multimatrix2_symv(const dgb_ell* hx, const dgb_ell* hy, int kind, double alpha, const double* x, double beta, double* y, double* temp,
                          dgb_stream_t s) {
    if (!hx || !hy || !x || !y) { set_error("dgb_multimatrix2_symv: NULL argument"); return DGB_ERR_INVALID; }
    const EllDev& X = *reinterpret_cast<const EllDev*>(hx);
    const EllDev& Y = *reinterpret_cast<const EllDev*>(hy);
    cudaStream_t st = as_stream(s);
    if ((kind == 1 || kind == 2) && !g_multimat_two_pass && x != y) {
        const bool projection = kind == 1;
        const int Nxc = projection ? X.num_rows : X.num_cols, Nyc = projection ? Y.num_rows : Y.num_cols;
        if (X.n != Y.n || X.n < 2 || X.n > 4 || X.bpl != (projection ? 2 : 1) || Y.bpl != X.bpl) { set_error("dgb_multimatrix2_symv: kind does not match the matrices"); return DGB_ERR_INVALID; }
        return half_dispatch(X, Y, kind, Nxc, Nyc, alpha, beta, x, y, st);
    }
    if (!temp) { set_error("dgb_multimatrix2_symv: the two-pass path needs the temporary"); return DGB_ERR_INVALID; }
    int e;
    if ((e = ell_symv(X, 1., x, 0., temp, st, false))) return e;
    return ell_symv(Y, alpha, temp, beta, y, st, false);
}
int dgb_multigrid2d_set_two_pass(int on) { g_multimat_two_pass = on ? 1 : 0; return 0; }
int dgb_multigrid2d_stages(const dgb_multigrid2d* h) { return reinterpret_cast<const Multigrid2d*>(h)->stages; }
int dgb_multigrid2d_grid(const dgb_multigrid2d* h, int stage, dgb_grid* grid, size_t* size) {
    const Multigrid2d* m = reinterpret_cast<const Multigrid2d*>(h);
    if (stage < 0 || stage >= m->stages) { set_error("dgb_multigrid2d_grid: stage out of range"); return DGB_ERR_INVALID; }
    if (grid) *grid = m->grids[stage];
    if (size) *size = m->sizes[stage];
    return 0;
}
// NestedGrids::project (multigrid.h:94-99): out[0] = src, out[u+1] = projection(u) out[u]
int dgb_multigrid2d_project(dgb_multigrid2d* h, const double* src, double* const* out, dgb_stream_t s) {
    Multigrid2d* m = reinterpret_cast<Multigrid2d*>(h);
    int e;
    if ((e = dgb_copy(m->sizes[0], src, out[0], s))) return e;
    for (int u = 0; u < m->stages - 1; u++)
        if ((e = multimat_symv(m->project[u], 1., out[u], 0., out[u + 1], as_stream(s)))) return e;
    return 0;
}
int dgb_multigrid2d_interpolate(dgb_multigrid2d* h, int coarse_stage, double alpha, const double* xc, double beta, double* xf,
                                dgb_stream_t s) {
    Multigrid2d* m = reinterpret_cast<Multigrid2d*>(h);
    if (coarse_stage < 1 || coarse_stage >= m->stages) { set_error("dgb_multigrid2d_interpolate: stage out of range"); return DGB_ERR_INVALID; }
    return multimat_symv(m->inter[coarse_stage - 1], alpha, xc, beta, xf, as_stream(s));
}
// MultigridCG2d::solve (multigrid.h:627-658) = nested_iterations (multigrid.h:197-245) with PCG on every stage
int dgb_multigrid2d_solve(dgb_multigrid2d* h, dgb_elliptic2d* const* ops, const double* const* precond,
                          const double* const* weights, double* x, const double* b, const double* eps, int* numbers,
                          dgb_stream_t s) {
    Multigrid2d* m = reinterpret_cast<Multigrid2d*>(h);
    cudaStream_t st = as_stream(s);
    const int S = m->stages;
    std::vector<Elliptic2dPlan*> A(S);
    for (int u = 0; u < S; u++) {
        A[u] = reinterpret_cast<Elliptic2dPlan*>(ops[u]);
        if (!A[u] || A[u]->size != m->sizes[u]) { set_error("dgb_multigrid2d_solve: operator %d does not match the stage size", u); return DGB_ERR_INVALID; }
        if (!precond[u] || !weights[u]) { set_error("dgb_multigrid2d_solve: preconditioner / weights of stage %d missing", u); return DGB_ERR_INVALID; }
    }
    int e;
    // residual r = b - A x                                                     multigrid.h:205-206
    if ((e = elliptic2d_symv(*A[0], 1., x, 0., m->r[0], st, false))) return e;
    if ((e = dgb_axpby(m->sizes[0], 1., b, -1., m->r[0], s))) return e;
    if ((e = dgb_copy(m->sizes[0], x, m->x[0], s))) return e;                   // :208
    for (int u = 0; u < S - 1; u++) {
        if ((e = multimat_symv(m->project[u], 1., m->r[u], 0., m->r[u + 1], st))) return e;   // :211
        if ((e = multimat_symv(m->project[u], 1., m->x[u], 0., m->x[u + 1], st))) return e;   // :212
        if ((e = elliptic2d_symv(*A[u + 1], 1., m->x[u + 1], 0., m->b[u + 1], st, false))) return e;  // :214
        if ((e = dgb_axpbyz(m->sizes[u + 1], 1., m->b[u + 1], 1., m->r[u + 1], m->b[u + 1], s))) return e;  // :215
        if ((e = dgb_copy(m->sizes[u + 1], m->x[u + 1], m->w[u + 1], s))) return e;          // :216
    }
    for (int u = S - 1; u > 0; u--) {
        int it = 0;
        e = pcg_solve(*m->pcg[u], *A[u], m->x[u], m->b[u], precond[u], weights[u], eps[u], 1., 10, (int)m->sizes[u], &it, st);  // :646
        if (numbers) numbers[u] = it;
        if (e) return e;
        if ((e = dgb_axpbyz(m->sizes[u], 1., m->x[u], -1., m->w[u], m->x[u], s))) return e;  // :230 delta
        if ((e = multimat_symv(m->inter[u - 1], 1., m->x[u], 1., m->x[u - 1], st))) return e;  // :232
    }
    if ((e = dgb_copy(m->sizes[0], m->x[0], x, s))) return e;                   // :236
    int it = 0;
    e = pcg_solve(*m->pcg[0], *A[0], x, b, precond[0], weights[0], eps[0], 1., 1, (int)m->sizes[0], &it, st);  // :643
    if (numbers) numbers[0] = it;
    return e;
}
}
