/* TEST INFRASTRUCTURE ONLY -- not part of the product.
 *
 * Plain-C restatement of the reference's (feltor-dev/feltor v8.2.2) algorithms on the hot path:
 * blas1 functors, exblas superaccumulator dot, Ell/Coo sparse block symv, CSR spmv, Elliptic2d apply,
 * PCG and the nested-iteration multigrid.  Every function cites the reference file:line it follows
 * (paths relative to /root/reference/).  Compiled by oracle/Makefile into oracle/libdgoracle.so with
 * `gcc -O2 -mfma -ffp-contract=off` so that every fma() below is a real fused multiply-add and nothing
 * else is contracted -- exactly the arithmetic the reference performs when DG_FMA is active
 * (inc/dg/backend/config.h:21-26).
 *
 * Parity pin: tests/test_oracle.py checks this file against the reference's own golden vectors
 * (inc/dg/blas1_t.cpp:102-184, inc/dg/topology/evaluation_t.cpp:56-175,
 * inc/dg/topology/derivatives_t.cpp:54-133), against the committed fixtures in tests/golden/ that were
 * produced by the unmodified reference (oracle/_ref, script tests/golden/make_golden.py), and -- when
 * oracle/_ref/libdgref.so is present -- live against the reference on random inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * blas1 functors: inc/dg/subroutines.h:231-384 (explicit DG_FMA order), dispatch shortcuts of
 * inc/dg/blas1.h:243-566 are applied by the caller (tests/host layer), not here.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_copy(int n, const double* x, double* y) { for (int i = 0; i < n; i++) y[i] = x[i]; }
/* Scal subroutines.h:233-244 */
ORC_API void orc_scal(int n, double* x, double a) { for (int i = 0; i < n; i++) x[i] *= a; }
/* Plus subroutines.h:247-258 */
ORC_API void orc_plus(int n, double* x, double a) { for (int i = 0; i < n; i++) x[i] += a; }
/* Axpby subroutines.h:260-274 */
ORC_API void orc_axpby(int n, double a, const double* x, double b, double* y) {
    for (int i = 0; i < n; i++) { double t = y[i] * b; y[i] = fma(a, x[i], t); }
}
/* z = a x + b y via Evaluate<equals,PairSum> blas1.h:382-385, subroutines.h:124-143: fma(a,x, b*y) */
ORC_API void orc_axpbyz(int n, double a, const double* x, double b, const double* y, double* z) {
    for (int i = 0; i < n; i++) z[i] = fma(a, x[i], b * y[i]);
}
/* Axpbypgz subroutines.h:294-310 */
ORC_API void orc_axpbypgz(int n, double a, const double* x, double b, const double* y, double g, double* z) {
    for (int i = 0; i < n; i++) { double t = z[i] * g; t = fma(a, x[i], t); z[i] = fma(b, y[i], t); }
}
/* PointwiseDot 3-arg subroutines.h:313-323 */
ORC_API void orc_pointwiseDot(int n, double a, const double* x, const double* y, double b, double* z) {
    for (int i = 0; i < n; i++) { double t = z[i] * b; z[i] = fma(a * x[i], y[i], t); }
}
/* AxyPby subroutines.h:276-292 (pointwiseDot with y aliasing an input, blas1.h:413-421) */
ORC_API void orc_axypby(int n, double a, const double* x, double b, double* y) {
    for (int i = 0; i < n; i++) { double tmp = y[i]; double t = tmp * b; y[i] = fma(a * x[i], tmp, t); }
}
/* z = x*y via Evaluate<equals,PairSum>(x,y) blas1.h:441-444: PairSum::sum(alpha=x, x=y) = x*y */
ORC_API void orc_pointwiseDot_xy(int n, const double* x, const double* y, double* z) {
    for (int i = 0; i < n; i++) z[i] = x[i] * y[i];
}
/* PointwiseDot 4-arg subroutines.h:325-330 */
ORC_API void orc_pointwiseDot3(int n, double a, const double* x1, const double* x2, const double* x3, double b, double* y) {
    for (int i = 0; i < n; i++) { double t = y[i] * b; y[i] = fma(a * x1[i], x2[i] * x3[i], t); }
}
/* PointwiseDot2 subroutines.h:336-352 */
ORC_API void orc_pointwiseDot2(int n, double a, const double* x1, const double* y1, double b, const double* x2,
                               const double* y2, double g, double* z) {
    for (int i = 0; i < n; i++) {
        double t = z[i] * g;
        t = fma(a * x1[i], y1[i], t);
        z[i] = fma(b * x2[i], y2[i], t);
    }
}
/* PointwiseDivide subroutines.h:365-384 */
ORC_API void orc_pointwiseDivide(int n, double a, const double* x, const double* y, double b, double* z) {
    for (int i = 0; i < n; i++) { double t = z[i] * b; z[i] = fma(a, x[i] / y[i], t); }
}
ORC_API void orc_pointwiseDivide_alias(int n, double a, const double* y, double b, double* z) {
    for (int i = 0; i < n; i++) { double tmp = z[i]; double t = tmp * b; z[i] = fma(a, tmp / y[i], t); }
}
ORC_API void orc_pointwiseDivide_xy(int n, const double* x, const double* y, double* z) {
    for (int i = 0; i < n; i++) z[i] = x[i] / y[i];
}
/* TensorMultiply2d inc/dg/topology/multiply.h:18-32; NULL tensor component pointers mean the constant
 * value the SparseTensor supplies for them (1 on the diagonal, 0 off-diagonal, tensor.h) */
ORC_API void orc_tensor_multiply2d(int n, const double* lambda, double lambda_s, const double* t00, const double* t01,
                                   const double* t10, const double* t11, const double* in0, const double* in1,
                                   double mu, double* out0, double* out1) {
    for (int i = 0; i < n; i++) {
        double l = lambda ? lambda[i] : lambda_s;
        double a = t00 ? t00[i] : 1., b = t01 ? t01[i] : 0., c = t10 ? t10[i] : 0., d = t11 ? t11[i] : 1.;
        double i0 = in0[i], i1 = in1[i];
        double tmp0 = fma(a, i0, b * i1);
        double tmp1 = fma(c, i0, d * i1);
        double temp = out1[i] * mu;
        out1[i] = fma(l, tmp1, temp);
        temp = out0[i] * mu;
        out0[i] = fma(l, tmp0, temp);
    }
}
/* TensorMultiply3d inc/dg/topology/multiply.h:34-58; t = 9 row-major component pointers (NULL = identity's constant) */
ORC_API void orc_tensor_multiply3d(int n, const double* lambda, double lambda_s, const double* const* t,
                                   const double* const* in, double mu, double* const* out) {
    for (int i = 0; i < n; i++) {
        double l = lambda ? lambda[i] : lambda_s, T[9];
        for (int k = 0; k < 9; k++) T[k] = (t && t[k]) ? t[k][i] : (k % 4 == 0 ? 1. : 0.);
        double i0 = in[0][i], i1 = in[1][i], i2 = in[2][i];
        double tmp0 = fma(T[0], i0, fma(T[1], i1, T[2] * i2));
        double tmp1 = fma(T[3], i0, fma(T[4], i1, T[5] * i2));
        double tmp2 = fma(T[6], i0, fma(T[7], i1, T[8] * i2));
        double temp = out[2][i] * mu;
        out[2][i] = fma(l, tmp2, temp);
        temp = out[1][i] * mu;
        out[1][i] = fma(l, tmp1, temp);
        temp = out[0][i] * mu;
        out[0][i] = fma(l, tmp0, temp);
    }
}
/* The parallel-derivative formulas inc/geometries/ds.h:743-1000 (+ functors :67-135) written left to right with
 * separately rounded operations (the reference's are user lambdas whose contraction is up to its compiler, so parity
 * with a compiled reference is to ~1e-14, not bitwise).  kind and operand order as dgb_ds_apply / dgb_ds_apply_vol:
 * 0 forward (a=f,b=fp) 1 backward (a=f,b=fm) 2 centered (a=fm,b=fp) 3 forward2 (f,fp,fpp) 4 backward2 (f,fm,fmm)
 * 5 dss_centered (fm,f,fp) 6 dssd_centered (fm,f,fp) 7 divBackward (fm,f) 8 divForward (f,fp) 9 divCentered (fm,fp)
 * 10 average (fm,fp) */
ORC_API void orc_ds_apply(int kind, int n, double alpha, const double* a, const double* b, const double* c,
                          const double* Gm, const double* G0, const double* Gp, const double* bm, const double* b0,
                          const double* bp, double delta, double beta, double* g) {
    for (int i = 0; i < n; i++) {
        double v = 0.;
        switch (kind) {
            case 0: v = alpha * b0[i] * (b[i] - a[i]) / delta; break;
            case 1: v = alpha * b0[i] * (a[i] - b[i]) / delta; break;
            case 2: v = alpha * b0[i] * (b[i] - a[i]) / 2. / delta; break;
            case 3: v = alpha * b0[i] * (-3. * a[i] + 4. * b[i] - c[i]) / 2. / delta; break;
            case 4: v = alpha * b0[i] * (3. * a[i] - 4. * b[i] + c[i]) / 2. / delta; break;
            case 5: case 6: {
                double bP2 = (bp[i] + b0[i]) / 2., bM2 = (bm[i] + b0[i]) / 2.;
                double fm2 = (b[i] - a[i]) / delta, fp2 = (c[i] - b[i]) / delta;
                if (kind == 5) v = alpha * b0[i] * (bP2 * fp2 - bM2 * fm2) / delta;
                else {
                    double gp2 = (Gp[i] + G0[i]) / G0[i] / 2., gm2 = (Gm[i] + G0[i]) / G0[i] / 2.;
                    v = alpha * (gp2 * fp2 * bP2 * bP2 - bM2 * bM2 * gm2 * fm2) / delta;
                }
                break;
            }
            case 7: v = alpha * (b0[i] * G0[i] * b[i] - bm[i] * Gm[i] * a[i]) / G0[i] / delta; break;
            case 8: v = alpha * (bp[i] * Gp[i] * b[i] - b0[i] * G0[i] * a[i]) / G0[i] / delta; break;
            case 9: v = alpha * (b[i] * Gp[i] * bp[i] - a[i] * Gm[i] * bm[i]) / G0[i] / 2. / delta; break;
            case 10: v = alpha * (b[i] + a[i]) / 2.; break;
        }
        g[i] = beta == 0. ? v : v + beta * g[i];
    }
}
/* assign_bc_along_field_2nd / _1st (inc/geometries/ds.h:169-296): ghost values fmg, fpg of the minus / plus neighbours where the
 * field line leaves the domain (bbm, bbo, bbp = masks, hbm / hbp = distances to the wall); order 2 uses (fm, f, fp), order 1
 * (fm, fp); neu != 0: Neumann values (dbm, dbp) = (bv0, bv1), else Dirichlet values (fbm, fbp).  Left-to-right, separately
 * rounded (the reference's are user lambdas: parity ~1e-14) */
ORC_API void orc_assign_bc_along_field(int order, int neu, int n, double delta, const double* fm_, const double* f_,
                                       const double* fp_, const double* hbm_, const double* hbp_, const double* bbm_,
                                       const double* bbo_, const double* bbp_, double bv0, double bv1, double* fmg, double* fpg) {
    for (int i = 0; i < n; i++) {
        double fm = fm_[i], fp = fp_[i], fo = f_ ? f_[i] : 0., hm = delta, hp = delta;
        double hbm = hbm_ ? hbm_[i] : 0., hbp = hbp_ ? hbp_[i] : 0., bbm = bbm_[i], bbo = bbo_ ? bbo_[i] : 0., bbp = bbp_[i];
        double plus, minus, bothP = 0., bothM = 0.;
        if (order == 2 && neu) {
            double dbm = bv0, dbp = bv1;
            plus = dbp * hp * (hm + hp) / (2. * hbp + hm) + fo * (2. * hbp + hm - hp) * (hm + hp) / hm / (2. * hbp + hm) +
                   fm * hp * (-2. * hbp + hp) / hm / (2. * hbp + hm);
            minus = fp * hm * (-2. * hbm + hm) / hp / (2. * hbm + hp) - dbm * hm * (hm + hp) / (2. * hbm + hp) +
                    fo * (2. * hbm - hm + hp) * (hm + hp) / hp / (2. * hbm + hp);
            bothM = fo + dbp * hm * (-2. * hbm + hm) / 2. / (hbm + hbp) - dbm * hm * (2. * hbp + hm) / 2. / (hbm + hbp);
            bothP = fo + dbp * hp * (2. * hbm + hp) / 2. / (hbm + hbp) + dbm * hp * (2. * hbp - hp) / 2. / (hbm + hbp);
        } else if (order == 2) {
            double fbm = bv0, fbp = bv1;
            plus = fm * hp * (-hbp + hp) / hm / (hbp + hm) + fo * (hbp - hp) * (hm + hp) / hbp / hm + fbp * hp * (hm + hp) / hbp / (hbp + hm);
            minus = +fo * (hbm - hm) * (hm + hp) / hbm / hp + fbm * hm * (hm + hp) / hbm / (hbm + hp) + fp * hm * (-hbm + hm) / hp / (hbm + hp);
            bothM = fbp * hm * (-hbm + hm) / hbp / (hbm + hbp) + fo * (hbm - hm) * (hbp + hm) / hbm / hbp + fbm * hm * (hbp + hm) / hbm / (hbm + hbp);
            bothP = fo * (hbp - hp) * (hbm + hp) / hbm / hbp + fbp * hp * (hbm + hp) / hbp / (hbm + hbp) + fbm * hp * (-hbp + hp) / hbm / (hbm + hbp);
        } else if (neu) {
            double dbm = bv0, dbp = bv1;
            plus = fm + dbp * (hp + hm);
            minus = fp - dbm * (hp + hm);
            fmg[i] = (1. - bbm) * fm + bbm * minus;
            fpg[i] = (1. - bbp) * fp + bbp * plus;
            continue;
        } else {
            double fbm = bv0, fbp = bv1;
            plus = fm + (fbp - fm) / (hbp + hm) * (hp + hm);
            minus = fp - (hp + hm) * (fp - fbm) / (hp + hbm);
            bothM = fbp + (fbp - fbm) / (hbp + hbm) * (hp + hbm);
            bothP = fbp - (fbp - fbm) / (hbp + hbm) * (hbp + hm);
        }
        fmg[i] = (1. - bbo - bbm) * fm + bbm * minus + bbo * bothM;
        fpg[i] = (1. - bbo - bbp) * fp + bbp * plus + bbo * bothP;
    }
}
/* dg::blas2::stencil / parallel_for with the library's CSR stencil functors, inc/dg/topology/filter.h:84-266.
 * The (lower) median is an order statistic -- rank (n+1)/2 of the stencil values -- so it is restated here by sorting;
 * kind 0 CSRMedianFilter, 1 CSRSWMFilter(alpha), 2 CSRAverageFilter, 3 CSRSymvFilter, 4 CSRSlopeLimiter(alpha) */
static double orc_minmod2(double a, double b) {   /* dg::MinMod, functors.h:255-285 */
    if (a > 0 && b > 0) return a < b ? a : b;
    if (a < 0 && b < 0) return a > b ? a : b;
    return 0.;
}
/* CSRSlopeLimiter::operator() (filter.h:291-333) on row i: rows with 3 n entries (one per cell) limit the cell's n values, the
 * others do nothing.  The sums are written as `a += b*c` in the reference; both of its compilers (gcc -mfma with the default
 * -ffp-contract=fast, nvcc with -fmad=true) contract them, so they are FMAs here. */
static void orc_slope_limiter_row(int i, const int* pos, const int* idx, const double* val, double mod, const double* x, double* y) {
    int k = pos[i], n = (pos[i + 1] - pos[i]) / 3;
    if (n == 0) return;
    for (int u = 0; u < n; u++) y[idx[k + n + u]] = x[idx[k + n + u]];
    double uM = 0, u0 = 0, uP = 0, u1 = 0;
    for (int u = 0; u < n; u++) {
        uM = fma(x[idx[k + u]], fabs(val[k + u]), uM);
        u0 = fma(x[idx[k + n + u]], fabs(val[k + u]), u0);
        u1 = fma(x[idx[k + n + u]], val[k + n + u], u1);
        uP = fma(x[idx[k + 2 * n + u]], fabs(val[k + u]), uP);
    }
    if (val[k] < 0) uM *= -1;
    if (val[k + 2 * n] > 0) uP *= -1;
    if (fabs(u1) <= mod) return;
    double m = orc_minmod2(orc_minmod2(u1, uP - u0), u0 - uM);
    if (m == u1) return;
    for (int u = 0; u < n; u++)
    {   /* gcc shares the product m*v between the two arms of the conditional, so it is rounded on its own (pinned on the
         * golden vectors of the reference's OpenMP build, tests/test_limiter.py) */
        double t = m * val[k + 2 * n + u];
        y[idx[k + n + u]] = val[k + 2 * n] > 0 ? u0 - t : u0 + t;
    }
}
static int orc_cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return x < y ? -1 : x > y;
}
static double orc_row_median(int b, int e, const int* idx, const double* x, int use_dev, double center, double* scratch) {
    int n = e - b;
    for (int k = 0; k < n; k++) scratch[k] = use_dev ? fabs(x[idx[b + k]] - center) : x[idx[b + k]];
    qsort(scratch, n, sizeof(double), orc_cmp_double);
    return scratch[(n + 1) / 2 - 1];
}
ORC_API void orc_csr_stencil(int kind, int num_rows, const int* pos, const int* idx, const double* val, double alpha,
                             const double* x, double* y) {
    int maxn = 1;
    for (int i = 0; i < num_rows; i++) if (pos[i + 1] - pos[i] > maxn) maxn = pos[i + 1] - pos[i];
    double* scratch = (double*)malloc(sizeof(double) * maxn);
    for (int i = 0; i < num_rows; i++) {
        int b = pos[i], e = pos[i + 1], n = e - b;
        if (kind == 4) { orc_slope_limiter_row(i, pos, idx, val, alpha, x, y); continue; }
        if (kind == 0) y[i] = orc_row_median(b, e, idx, x, 0, 0., scratch);
        else if (kind == 1) {
            double med = orc_row_median(b, e, idx, x, 0, 0., scratch);
            double amd = orc_row_median(b, e, idx, x, 1, med, scratch);
            y[i] = fabs(x[i] - med) > alpha * amd ? med : x[i];
        } else if (kind == 2) {
            double t = 0;
            for (int k = b; k < e; k++) t += x[idx[k]] / (double)n;
            y[i] = t;
        } else {
            double t = 0;
            for (int k = b; k < e; k++) t += x[idx[k]] * val[k];   /* y += x*v (test filter; contraction is up to the reference's compiler) */
            y[i] = t;
        }
    }
    free(scratch);
}
/* dg::detail::spgemm_cpu_kernel, inc/dg/backend/sparsematrix_cpu.h:19-95: A = B C.  Pass 1 (A_idx == NULL) fills A_pos only and
 * returns the number of entries; pass 2 fills A_idx (sorted per row) and A_val.  The workspace update `w[j] += b*c` is an FMA in
 * the reference's build (gcc -mfma contracts it; pinned on the live reference in tests/test_spgemm.py). */
static int orc_cmp_int(const void* a, const void* b) { int x = *(const int*)a, y = *(const int*)b; return x < y ? -1 : x > y; }
ORC_API long long orc_spgemm(int B_rows, int C_cols, const int* B_pos, const int* B_idx, const double* B_val, const int* C_pos,
                             const int* C_idx, const double* C_val, int* A_pos, int* A_idx, double* A_val) {
    char* seen = (char*)calloc(C_cols > 0 ? C_cols : 1, 1);
    double* w = (double*)calloc(C_cols > 0 ? C_cols : 1, sizeof(double));
    int* wlist = (int*)malloc(sizeof(int) * (C_cols > 0 ? C_cols : 1));
    long long run = 0;
    for (int i = 0; i < B_rows; i++) {
        int m = 0;
        for (int pB = B_pos[i]; pB < B_pos[i + 1]; pB++) {
            int k = B_idx[pB];
            for (int pC = C_pos[k]; pC < C_pos[k + 1]; pC++) {
                int j = C_idx[pC];
                if (!seen[j]) { seen[j] = 1; wlist[m++] = j; }
                if (A_idx) w[j] = fma(B_val[pB], C_val[pC], w[j]);
            }
        }
        qsort(wlist, m, sizeof(int), orc_cmp_int);
        if (!A_idx) A_pos[i] = (int)run;
        for (int q = 0; q < m; q++) {
            int j = wlist[q];
            if (A_idx) { A_idx[run + q] = j; A_val[run + q] = w[j]; w[j] = 0; }
            seen[j] = 0;
        }
        run += m;
    }
    if (!A_idx) A_pos[B_rows] = (int)run;
    free(seen); free(w); free(wlist);
    return run;
}
/* EmbeddedPairSum subroutines.h:179-204: y = b0*y + sum b_i k_i ; yt likewise.  k = array of pointers */
ORC_API void orc_embedded_pair_sum(int n, double* y, double* yt, double b0, double bt0, int nk, const double* b,
                                   const double* bt, const double* const* k) {
    for (int i = 0; i < n; i++) {
        double a = b0 * y[i], at = bt0 * yt[i];
        for (int s = 0; s < nk; s++) { a = fma(b[s], k[s][i], a); at = fma(bt[s], k[s][i], at); }
        y[i] = a; yt[i] = at;
    }
}

/* ------------------------------------------------------------------------------------------------
 * exblas superaccumulator: inc/dg/backend/exblas/config.h:86-92, accumulate.h:171-349, mylibm.hpp
 * ---------------------------------------------------------------------------------------------- */
#define KRX 8
#define DIGITS 56
#define F_WORDS 20
#define E_WORDS 19
#define BIN_COUNT 39
static const double DELTASCALE = 72057594037927936.0; /* 2^56 */

/* mylibm.hpp:168-183 (portable branch): returns old word, sets signed-overflow flag */
static int64_t xadd(int64_t* mem, int64_t x, unsigned char* of) {
    int64_t y = *mem;
    uint64_t r = (uint64_t)y + (uint64_t)x;
    *mem = (int64_t)r;
    int64_t x63 = (x >> 63) & 1, y63 = (y >> 63) & 1, r63 = ((int64_t)r >> 63) & 1;
    int64_t c62 = r63 ^ x63 ^ y63;
    int64_t c63 = (x63 & y63) | (c62 & (x63 | y63));
    *of = (unsigned char)(c63 ^ c62);
    return y;
}
/* accumulate.h:171-208 */
static void AccumulateWord(int64_t* acc, int i, int64_t x) {
    unsigned char overflow;
    int64_t carry = x, carrybit;
    int64_t oldword = xadd(&acc[i], x, &overflow);
    while (overflow) {
        carry = (oldword + carry) >> DIGITS;
        int s = oldword > 0;
        carrybit = (s ? (int64_t)(1ll << KRX) : (int64_t)((unsigned long long)(-1ll) << KRX));
        xadd(&acc[i], (int64_t)(-(uint64_t)((uint64_t)carry << DIGITS)), &overflow);
        carry += carrybit;
        ++i;
        if (i >= BIN_COUNT) return;
        oldword = xadd(&acc[i], carry, &overflow);
    }
}
/* mylibm.hpp:72-81 exponent, :95-107 myldexp */
static int exponent_of(double x) {
    union { double d; uint64_t i; } c; c.d = x;
    uint64_t e = ((c.i >> 52) & 0x7ff) - 0x3ff;
    return (int)e;
}
static double myldexp(double x, int e) {
    union { double d; uint64_t i; } c; c.d = x;
    c.i += (uint64_t)e << 52;
    return c.d;
}
/* accumulate.h:217-236 */
ORC_API void orc_accumulate(int64_t* acc, double x) {
    if (x == 0) return;
    int e = exponent_of(x);
    int exp_word = e / DIGITS;
    int iup = exp_word + F_WORDS;
    double xscaled = myldexp(x, -DIGITS * exp_word);
    for (int i = iup; i >= 0 && xscaled != 0; --i) {
        double xrounded = rint(xscaled);
        int64_t xint = llrint(xscaled);
        AccumulateWord(acc, i, xint);
        xscaled -= xrounded;
        xscaled *= DELTASCALE;
    }
}
/* accumulate.h:267-285; returns sign */
ORC_API int orc_normalize(int64_t* acc) {
    int imin = 0;
    int64_t carry_in = acc[imin] >> DIGITS;
    acc[imin] -= (int64_t)((uint64_t)carry_in << DIGITS);
    int i;
    for (i = imin + 1; i < BIN_COUNT; ++i) {
        acc[i] += carry_in;
        int64_t carry_out = acc[i] >> DIGITS;
        acc[i] -= (int64_t)((uint64_t)carry_out << DIGITS);
        carry_in = carry_out;
    }
    int imax = i - 1;
    acc[imax] += (int64_t)((uint64_t)carry_in << DIGITS);
    return carry_in < 0;
}
/* mylibm.hpp:118-134 */
static double OddRoundSumNonnegative(double th, double tl) {
    union { double d; int64_t l; } thdb;
    thdb.d = th + tl;
    thdb.l |= (tl != 0.0);
    return thdb.d;
}
/* accumulate.h:297-349 (modifies acc: it is normalised in place) */
ORC_API double orc_round(int64_t* acc) {
    int imin = 0, imax = BIN_COUNT - 1;
    int negative = orc_normalize(acc);
    int i;
    for (i = imax; i >= imin && acc[i] == 0; --i) {}
    if (negative) {
        for (; i >= imin && (acc[i] & ((1ll << DIGITS) - 1)) == ((1ll << DIGITS) - 1); --i) {}
    }
    if (i < 0) return 0.0;
    int64_t hiword = negative ? ((1ll << DIGITS) - 1) - acc[i] : acc[i];
    double rounded = (double)hiword;
    double hi = ldexp(rounded, (i - F_WORDS) * DIGITS);
    if (i == 0) return negative ? -hi : hi;
    hiword -= llrint(rounded);
    double mid = ldexp((double)hiword, (i - F_WORDS) * DIGITS);
    int64_t sticky = 0;
    for (int j = imin; j != i - 1; ++j) sticky |= negative ? ((1ll << DIGITS) - acc[j]) : acc[j];
    int64_t loword = negative ? ((1ll << DIGITS) - acc[i - 1]) : acc[i - 1];
    loword |= !!sticky;
    double lo = ldexp((double)loword, (i - 1 - F_WORDS) * DIGITS);
    if (mid != 0) lo = OddRoundSumNonnegative(mid, lo);
    hi = hi + lo;
    return negative ? -hi : hi;
}
/* exdot: exdot_serial.h:62-135 / exdot_omp.h:95-245.  The FPE cache in front of the superaccumulator is an
 * optimisation only (ExSUM.FPE.hpp:100-116): the accumulated VALUE is the exact sum of the individually
 * rounded products, so accumulating each product directly gives the same normalised accumulator.
 * Returns status (1 if a product is non-finite, blas1.h:161).  acc is returned NORMALISED. */
ORC_API int orc_exdot2(int n, const double* x, const double* y, int64_t* acc) {
    int status = 0;
    memset(acc, 0, BIN_COUNT * sizeof(int64_t));
    for (int i = 0; i < n; i++) {
        double p = x[i] * y[i];
        if (!isfinite(p)) status = 1;
        else orc_accumulate(acc, p);
    }
    orc_normalize(acc);
    return status;
}
/* 3 operands: round(round(x*w)*y), exdot_serial.h:126-128 */
ORC_API int orc_exdot3(int n, const double* x, const double* w, const double* y, int64_t* acc) {
    int status = 0;
    memset(acc, 0, BIN_COUNT * sizeof(int64_t));
    for (int i = 0; i < n; i++) {
        double p1 = x[i] * w[i];
        double p = p1 * y[i];
        if (!isfinite(p)) status = 1;
        else orc_accumulate(acc, p);
    }
    orc_normalize(acc);
    return status;
}
/* blas1::dot / blas2::dot value (blas1.h:152-170, blas2.h:94-115); *status as above */
ORC_API double orc_dot2(int n, const double* x, const double* y, int* status) {
    int64_t acc[BIN_COUNT];
    *status = orc_exdot2(n, x, y, acc);
    return orc_round(acc);
}
ORC_API double orc_dot3(int n, const double* x, const double* w, const double* y, int* status) {
    int64_t acc[BIN_COUNT];
    *status = orc_exdot3(n, x, w, y, acc);
    return orc_round(acc);
}
/* word-wise sum of two NORMALISED accumulators followed by Normalize: the cross-rank combine of
 * inc/dg/backend/exblas/mpi_accumulate.h:94-125 */
ORC_API void orc_superacc_add(int64_t* acc, const int64_t* other) {
    for (int i = 0; i < BIN_COUNT; i++) acc[i] += other[i];
    orc_normalize(acc);
}

/* ------------------------------------------------------------------------------------------------
 * EllSparseBlockMat symv: generic kernel inc/dg/backend/sparseblockmat_omp_kernels.h:10-55
 * (the specialised kernels :57-290 perform the same arithmetic per output element; the only difference
 * is that for right_size==1 an invalid column contributes fma(alpha, 0, y) instead of being skipped,
 * which changes nothing but the sign of a zero).
 * meta = {num_rows, num_cols, blocks_per_line, n, left_size, right_size, nblocks, rr0, rr1}
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_ell_symv(const int* meta, const double* data, const int* cols_idx, const int* data_idx, double alpha,
                          const double* x, double beta, double* y) {
    const int num_rows = meta[0], num_cols = meta[1], bpl = meta[2], n = meta[3], left = meta[4], right = meta[5];
    const int rr0 = meta[7], rr1 = meta[8];
    for (int s = 0; s < left; s++)
        for (int i = 0; i < num_rows; i++)
            for (int k = 0; k < n; k++)
                for (int j = rr0; j < rr1; j++) {
                    size_t I = ((size_t)(s * num_rows + i) * n + k) * right + j;
                    double yy = beta == 0 ? 0. : y[I] * beta;
                    for (int d = 0; d < bpl; d++) {
                        int C = cols_idx[i * bpl + d];
                        if (C == -1) continue;
                        size_t J = (size_t)(s * num_cols + C) * n;
                        int B = (data_idx[i * bpl + d] * n + k) * n;
                        double temp = 0;
                        for (int q = 0; q < n; q++) temp = fma(data[B + q], x[(J + q) * right + j], temp);
                        yy = fma(alpha, temp, yy);
                    }
                    y[I] = yy;
                }
}
/* CooSparseBlockMat symv: sparseblockmat_omp_kernels.h:356-377; x = array of chunk pointers, each chunk laid
 * out [q][s][j]; beta == 1 implied.  meta = {num_rows, num_cols, num_entries, n, left_size, right_size} */
ORC_API void orc_coo_symv(const int* meta, const double* data, const int* rows_idx, const int* cols_idx,
                          const int* data_idx, double alpha, const double* const* x, double* y) {
    const int num_rows = meta[0], num_entries = meta[2], n = meta[3], left = meta[4], right = meta[5];
    for (int s = 0; s < left; s++)
        for (int k = 0; k < n; k++)
            for (int j = 0; j < right; j++)
                for (int i = 0; i < num_entries; i++) {
                    size_t I = ((size_t)(s * num_rows + rows_idx[i]) * n + k) * right + j;
                    double temp = 0;
                    for (int q = 0; q < n; q++)
                        temp = fma(data[(data_idx[i] * n + k) * n + q], x[cols_idx[i]][((size_t)q * left + s) * right + j], temp);
                    y[I] = fma(alpha, temp, y[I]);
                }
}
/* CSR spmv: inc/dg/backend/sparsematrix_omp.h:17-52 */
ORC_API void orc_csr_spmv(int nrows, const int* pos, const int* idx, const double* val, double alpha, const double* x,
                          double beta, double* y) {
    if (beta == 1.) {
        for (int i = 0; i < nrows; i++)
            for (int jj = pos[i]; jj < pos[i + 1]; jj++) y[i] = fma(alpha * val[jj], x[idx[jj]], y[i]);
    } else {
        for (int i = 0; i < nrows; i++) {
            double temp = 0;
            for (int jj = pos[i]; jj < pos[i + 1]; jj++) temp = fma(alpha * val[jj], x[idx[jj]], temp);
            /* reference: y = fma(beta, y, temp); the CUDA backend (cuSPARSE) does not read y for beta==0 and
             * the product follows that (NaN in y is overwritten) */
            y[i] = beta == 0. ? temp : fma(beta, y[i], temp);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Elliptic2d::symv inc/dg/elliptic.h:428-458 as a composition of the kernels above.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    const int* meta; const double* data; const int* cols; const int* didx;
} orc_ell;
typedef struct {
    orc_ell leftx, lefty, rightx, righty, jumpx, jumpy;
    const double* sigma;                        /* m_sigma = chi_scalar * vol (elliptic.h:324-333) */
    const double* vol;                          /* m_vol, NULL = all ones (Cartesian) */
    const double* chi_xx; const double* chi_xy; const double* chi_yx; const double* chi_yy; /* NULL = identity */
    double jfactor; int chi_weight_jump; int size;
} orc_elliptic2d;

static void ell(const orc_ell* m, double alpha, const double* x, double beta, double* y) {
    orc_ell_symv(m->meta, m->data, m->cols, m->didx, alpha, x, beta, y);
}
/* work = 3*size doubles */
ORC_API void orc_elliptic2d_symv(const orc_elliptic2d* e, double alpha, const double* x, double beta, double* y,
                                 double* work) {
    int n = e->size;
    double *tempx = work, *tempy = work + n, *temp = work + 2 * n;
    ell(&e->rightx, 1., x, 0., tempx);                                   /* elliptic.h:431 */
    ell(&e->righty, 1., x, 0., tempy);                                   /* :432 */
    orc_tensor_multiply2d(n, e->sigma, 1., e->chi_xx, e->chi_xy, e->chi_yx, e->chi_yy, tempx, tempy, 0., tempx,
                          tempy);                                        /* :435 */
    ell(&e->lefty, 1., tempy, 0., temp);                                 /* :438 */
    ell(&e->leftx, -1., tempx, -1., temp);                               /* :439 */
    if (0.0 != e->jfactor) {                                             /* :442 */
        if (e->chi_weight_jump) {
            ell(&e->jumpx, e->jfactor, x, 0., tempx);
            ell(&e->jumpy, e->jfactor, x, 0., tempy);
            orc_tensor_multiply2d(n, e->sigma, 1., e->chi_xx, e->chi_xy, e->chi_yx, e->chi_yy, tempx, tempy, 0.,
                                  tempx, tempy);
            orc_axpbypgz(n, 1.0, tempx, 1.0, tempy, 1.0, temp);
        } else {
            ell(&e->jumpx, e->jfactor, x, 1., temp);                     /* :454 */
            ell(&e->jumpy, e->jfactor, x, 1., temp);                     /* :455 */
        }
    }
    /* :458 pointwiseDivide(alpha, temp, vol, beta, y): z*=b; z = fma(a, x/y, z) */
    for (int i = 0; i < n; i++) {
        double v = e->vol ? e->vol[i] : 1.;
        double t = beta == 0 ? 0. : y[i] * beta; /* product does not read y for beta==0 */
        y[i] = fma(alpha, temp[i] / v, t);
    }
}

/* ------------------------------------------------------------------------------------------------
 * PCG::solve inc/dg/pcg.h:136-195 with vector preconditioner P: symv(P,r,z) ==
 * blas1::pointwiseDot(P,r,z) == P*r (blas2_dispatch_shared.h:126-134, blas1.h:441-444).
 * A is the Elliptic2d above.  work = 6*size.  Returns iterations; max_iter if not converged.
 * ---------------------------------------------------------------------------------------------- */
ORC_API int orc_pcg_solve_elliptic2d(const orc_elliptic2d* A, double* x, const double* b, const double* P,
                                     const double* W, double eps, double nrmb_correction, int test_frequency,
                                     int max_iter, double* work, double* residuals /* optional, max_iter */) {
    int n = A->size, st;
    double *r = work, *p = work + n, *ap = work + 2 * n, *awork = work + 3 * n;
    double nrmb = sqrt(orc_dot3(n, b, W, b, &st));                          /* pcg.h:140 */
    double tol = eps * (nrmb + nrmb_correction);
    if (nrmb == 0) { for (int i = 0; i < n; i++) x[i] = 0; return 0; }    /* :150-154 */
    orc_elliptic2d_symv(A, 1., x, 0., r, awork);                            /* :155 */
    orc_axpby(n, 1., b, -1., r);                                            /* :156 */
    if (sqrt(orc_dot3(n, r, W, r, &st)) < tol) return 0;                    /* :157 */
    for (int i = 0; i < n; i++) p[i] = P[i] * r[i];                      /* :159 symv(P,r,p) == pointwiseDot(P,r,p) = P*r */
    double nrmzr_old = orc_dot3(n, p, W, r, &st);                           /* :160 */
    for (int it = 1; it < max_iter; it++) {
        orc_elliptic2d_symv(A, 1., p, 0., ap, awork);                       /* :165 */
        double alpha = nrmzr_old / orc_dot3(n, p, W, ap, &st);              /* :166 */
        orc_axpby(n, alpha, p, 1., x);                                      /* :167 */
        orc_axpby(n, -alpha, ap, 1., r);                                    /* :168 */
        if (0 == it % test_frequency) {                                     /* :169 */
            double res = sqrt(orc_dot3(n, r, W, r, &st));
            if (residuals) residuals[it] = res;
            if (res < tol) return it;                                       /* :177 */
        }
        for (int i = 0; i < n; i++) ap[i] = P[i] * r[i];                  /* :180 */
        double nrmzr_new = orc_dot3(n, ap, W, r, &st);                      /* :181 */
        orc_axpby(n, 1., ap, nrmzr_new / nrmzr_old, p);                     /* :182 */
        nrmzr_old = nrmzr_new;
    }
    return max_iter;
}
