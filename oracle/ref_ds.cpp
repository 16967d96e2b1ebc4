// TEST INFRASTRUCTURE (never linked into or called by the product).
// The UNMODIFIED free functions dg::geo::ds_* / dss_centered / dssd_centered / ds_div* / ds_average of the reference
// (inc/geometries/ds.h:743-1016) and dg::TensorMultiply3d (inc/dg/topology/multiply.h:34-58), instantiated on a mock
// FieldAligned that only hands out the metric fields those templates ask for (deltaPhi, bphi*, sqrtG*).
// Built by oracle/Makefile into oracle/_ref/libdgref_ds.so from the sources where they lie under /root/reference.
#include <cstdint>
#include <vector>
#include <omp.h>
#include "dg/algorithm.h"
#include "dg/geometries/geometries.h"

namespace {
using Vec = thrust::host_vector<double>;
struct MockFA {
    double delta;
    Vec gm, g0, gp, bm, b0, bp;
    Vec v_hbm, v_hbp, v_bbm, v_bbo, v_bbp;
    double deltaPhi() const { return delta; }
    const Vec& hbm() const { return v_hbm; }
    const Vec& hbp() const { return v_hbp; }
    const Vec& bbm() const { return v_bbm; }
    const Vec& bbo() const { return v_bbo; }
    const Vec& bbp() const { return v_bbp; }
    const Vec& sqrtGm() const { return gm; }
    const Vec& sqrtG() const { return g0; }
    const Vec& sqrtGp() const { return gp; }
    const Vec& bphiM() const { return bm; }
    const Vec& bphi() const { return b0; }
    const Vec& bphiP() const { return bp; }
};
Vec mk(const double* p, int n) { return p ? Vec(p, p + n) : Vec(n, 0.); }
}  // namespace

extern "C" {
// kind 0..5 as dgb_ds_apply, 6 dssd_centered, 7 ds_divBackward (a=fm,b=f), 8 ds_divForward (a=f,b=fp),
// 9 ds_divCentered (a=fm,b=fp), 10 ds_average (a=fm,b=fp)
void ref_ds_apply(int kind, int n, double alpha, const double* a, const double* b, const double* c, const double* gm,
                  const double* g0, const double* gp, const double* bm, const double* b0, const double* bp, double delta,
                  double beta, double* g) {
    MockFA fa{delta, mk(gm, n), mk(g0, n), mk(gp, n), mk(bm, n), mk(b0, n), mk(bp, n)};
    Vec va = mk(a, n), vb = mk(b, n), vc = mk(c, n), vg(g, g + n);
    switch (kind) {
        case 0: dg::geo::ds_forward(fa, alpha, va, vb, beta, vg); break;
        case 1: dg::geo::ds_backward(fa, alpha, vb, va, beta, vg); break;       // (fm, f)
        case 2: dg::geo::ds_centered(fa, alpha, va, vb, beta, vg); break;       // (fm, fp)
        case 3: dg::geo::ds_forward2(fa, alpha, va, vb, vc, beta, vg); break;   // (f, fp, fpp)
        case 4: dg::geo::ds_backward2(fa, alpha, vc, vb, va, beta, vg); break;  // (fmm, fm, f)
        case 5: dg::geo::dss_centered(fa, alpha, va, vb, vc, beta, vg); break;  // (fm, f, fp)
        case 6: dg::geo::dssd_centered(fa, alpha, va, vb, vc, beta, vg); break;
        case 7: dg::geo::ds_divBackward(fa, alpha, va, vb, beta, vg); break;
        case 8: dg::geo::ds_divForward(fa, alpha, va, vb, beta, vg); break;
        case 9: dg::geo::ds_divCentered(fa, alpha, va, vb, beta, vg); break;
        case 10: dg::geo::ds_average(fa, alpha, va, vb, beta, vg); break;
    }
    for (int i = 0; i < n; i++) g[i] = vg[i];
}
// assign_bc_along_field_2nd (order 2: fm, f, fp) / _1st (order 1: fm, fp) of inc/geometries/ds.h:169-296; bound: 4 = NEU else DIR
void ref_assign_bc_along_field(int order, int bound, int n, double delta, const double* fm, const double* f, const double* fp,
                               const double* hbm, const double* hbp, const double* bbm, const double* bbo, const double* bbp,
                               double bv0, double bv1, double* fmg, double* fpg) {
    MockFA fa{delta, Vec(), Vec(), Vec(), Vec(), Vec(), Vec(), mk(hbm, n), mk(hbp, n), mk(bbm, n), mk(bbo, n), mk(bbp, n)};
    Vec vfm = mk(fm, n), vf = mk(f, n), vfp = mk(fp, n), g0(n, 0.), g1(n, 0.);
    dg::bc b = bound == 4 ? dg::NEU : dg::DIR;
    if (order == 2) dg::geo::assign_bc_along_field_2nd(fa, vfm, vf, vfp, g0, g1, b, {bv0, bv1});
    else dg::geo::assign_bc_along_field_1st(fa, vfm, vfp, g0, g1, b, {bv0, bv1});
    for (int i = 0; i < n; i++) { fmg[i] = g0[i]; fpg[i] = g1[i]; }
}
// t: 9 arrays (row major), in/out: 3 arrays each
void ref_tensor_multiply3d(int n, const double* lambda, const double* const* t, const double* const* in, double mu,
                           double* const* out) {
    Vec l = mk(lambda, n), T[9], I[3], O[3];
    for (int k = 0; k < 9; k++) T[k] = mk(t[k], n);
    for (int k = 0; k < 3; k++) { I[k] = mk(in[k], n); O[k] = mk(out[k], n); }
    dg::blas1::subroutine(dg::TensorMultiply3d(), l, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8], I[0], I[1], I[2], mu,
                          O[0], O[1], O[2]);
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < n; i++) out[k][i] = O[k][i];
}
}

// dg::blas2::parallel_for with the library's CSR stencil functors (inc/dg/topology/filter.h:174-266, blas2.h:413-454):
// kind 0 CSRMedianFilter, 1 CSRSWMFilter(alpha), 2 CSRAverageFilter, 3 CSRSymvFilter, 4 CSRSlopeLimiter(alpha) (filter.h:288-336)
extern "C" void ref_csr_stencil(int kind, int num_rows, int num_cols, const int* pos, const int* idx, const double* val,
                                double alpha, const double* x, double* y) {
    thrust::host_vector<int> p(pos, pos + num_rows + 1), c(idx, idx + pos[num_rows]);
    Vec v(val, val + pos[num_rows]), vx(x, x + num_cols), vy(y, y + num_rows);
    switch (kind) {
        case 0: dg::blas2::parallel_for(dg::CSRMedianFilter(), num_rows, p, c, v, vx, vy); break;
        case 1: dg::blas2::parallel_for(dg::CSRSWMFilter<double>(alpha), num_rows, p, c, v, vx, vy); break;
        case 2: dg::blas2::parallel_for(dg::CSRAverageFilter(), num_rows, p, c, v, vx, vy); break;
        case 3: dg::blas2::parallel_for(dg::CSRSymvFilter(), num_rows, p, c, v, vx, vy); break;
        case 4: dg::blas2::parallel_for(dg::CSRSlopeLimiter<double>(alpha), num_rows, p, c, v, vx, vy); break;
    }
    for (int i = 0; i < num_rows; i++) y[i] = vy[i];
}

// dg::SparseMatrix::operator* = dg::detail::spgemm_cpu_kernel (inc/dg/backend/sparsematrix.h:549-566, sparsematrix_cpu.h:19-95).
// Returns the number of entries; pass A_idx == NULL to learn it first (A_pos is filled either way).
extern "C" long long ref_spgemm(int B_rows, int B_cols, int C_cols, const int* B_pos, const int* B_idx, const double* B_val, const int* C_pos,
                                const int* C_idx, const double* C_val, int* A_pos, int* A_idx, double* A_val) {
    thrust::host_vector<int> bp(B_pos, B_pos + B_rows + 1), bi(B_idx, B_idx + B_pos[B_rows]), cp(C_pos, C_pos + B_cols + 1), ci(C_idx, C_idx + C_pos[B_cols]);
    Vec bv(B_val, B_val + B_pos[B_rows]), cv(C_val, C_val + C_pos[B_cols]);
    dg::IHMatrix B(B_rows, B_cols, bp, bi, bv), Cm(B_cols, C_cols, cp, ci, cv);
    dg::IHMatrix A = B * Cm;
    for (int i = 0; i <= B_rows; i++) A_pos[i] = A.row_offsets()[i];
    if (A_idx)
        for (size_t k = 0; k < A.values().size(); k++) { A_idx[k] = A.column_indices()[k]; A_val[k] = A.values()[k]; }
    return (long long)A.values().size();
}

// dg::create::limiter_stencil on a 1-d grid / along `direction` (0 x, 1 y) of a 2-d grid (inc/dg/topology/stencil.h:199-256)
extern "C" int ref_limiter_stencil(int ndim, const double* x0, const double* x1, int n, const int* N, const int* bc, int direction, int bound,
                                   int* pos, int* idx, double* val) {
    dg::IHMatrix m;
    if (ndim == 1) m = dg::create::limiter_stencil(dg::Grid1d(x0[0], x1[0], n, N[0], (dg::bc)bc[0]), (dg::bc)bound);
    else m = dg::create::limiter_stencil(direction == 0 ? dg::coo3d::x : dg::coo3d::y,
                                         dg::Grid2d(x0[0], x1[0], x0[1], x1[1], n, N[0], N[1], (dg::bc)bc[0], (dg::bc)bc[1]), (dg::bc)bound);
    if (pos)
        for (size_t i = 0; i < m.row_offsets().size(); i++) pos[i] = m.row_offsets()[i];
    if (idx)
        for (size_t i = 0; i < m.column_indices().size(); i++) { idx[i] = m.column_indices()[i]; val[i] = m.values()[i]; }
    return (int)m.values().size();
}

// dg::create::window_stencil on a 1-d / 2-d grid (inc/dg/topology/stencil.h:177-237); returns nnz, arrays caller-allocated
extern "C" int ref_window_stencil(int ndim, const double* x0, const double* x1, int n, const int* N, const int* bc, const int* window,
                                  int* pos, int* idx, double* val) {
    dg::IHMatrix m;
    if (ndim == 1) m = dg::create::window_stencil((unsigned)window[0], dg::Grid1d(x0[0], x1[0], n, N[0], (dg::bc)bc[0]), (dg::bc)bc[0]);
    else m = dg::create::window_stencil(std::array<int, 2>{window[0], window[1]},
                                        dg::Grid2d(x0[0], x1[0], x0[1], x1[1], n, N[0], N[1], (dg::bc)bc[0], (dg::bc)bc[1]), (dg::bc)bc[0], (dg::bc)bc[1]);
    for (size_t i = 0; i < m.row_offsets().size(); i++) pos[i] = m.row_offsets()[i];
    for (size_t i = 0; i < m.column_indices().size(); i++) { idx[i] = m.column_indices()[i]; val[i] = m.values()[i]; }
    return (int)m.values().size();
}
