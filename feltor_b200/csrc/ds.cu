// dg::geo::Fieldaligned::ePlus / eMinus (inc/geometries/fieldaligned.h:850-912) and the parallel-derivative
// formulas ds_forward/backward/centered, ds_forward2/backward2, dss_centered, dssd_centered, ds_div*, ds_average
// (inc/geometries/ds.h:744-1000,84-135).
//  * dgb_fa_eplus / dgb_fa_eminus: the 2-d interpolation matrix is applied to ALL planes in one launch
//    (csr_planes_kernel), then the ghost-cell fix-up of the last / first plane for non-periodic z with the same
//    three blas1 steps the reference performs.
//  * dgb_ds_centered_fused (periodic z): gathers f^+ = I^+ f[k+1] and f^- = I^- f[k-1] and evaluates
//    g = alpha bphi (f^+ - f^-)/2/dphi + beta g in ONE kernel: f is gathered, bphi read, g written -- the temporaries
//    m_tempP / m_tempM of DS::centered (ds.h:481-485) never touch HBM.
// The elementwise formulas are user lambdas in the reference (their contraction is compiler-dependent); we evaluate them
// left to right with separately rounded operations; parity with the reference is 1e-14 relative.
#include "common.cuh"

namespace dgb {

extern "C" int dgb_csr_spmv_planes(int, int, const int*, const int*, const double*, double, const double*, double, double*, int, int, dgb_stream_t);
extern "C" int dgb_axpbyz(size_t, double, const double*, double, const double*, double*, dgb_stream_t);
extern "C" int dgb_axpby(size_t, double, const double*, double, double*, dgb_stream_t);
extern "C" int dgb_pointwise_dot(size_t, double, const double*, const double*, double, double*, dgb_stream_t);

enum { DS_FORWARD = 0, DS_BACKWARD = 1, DS_CENTERED = 2, DS_FORWARD2 = 3, DS_BACKWARD2 = 4, DSS_CENTERED = 5,
       DSSD_CENTERED = 6, DS_DIV_BACKWARD = 7, DS_DIV_FORWARD = 8, DS_DIV_CENTERED = 9, DS_AVERAGE = 10 };

// a, b, c: the shifted fields in the argument order of the reference lambdas; bm, b0, bp: bphi on the minus/own/plus plane
template <int KIND>
__device__ __forceinline__ double ds_formula(double alpha, double beta, double delta, double g, double a, double b, double c,
                                              double bm, double b0, double bp) {
    double v;
    if (KIND == DS_FORWARD) v = __ddiv_rn(__dmul_rn(__dmul_rn(alpha, b0), __dsub_rn(b, a)), delta);            // a = f, b = fp
    else if (KIND == DS_BACKWARD) v = __ddiv_rn(__dmul_rn(__dmul_rn(alpha, b0), __dsub_rn(a, b)), delta);      // a = f, b = fm
    else if (KIND == DS_CENTERED) v = __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(alpha, b0), __dsub_rn(b, a)), 2.), delta);  // a = fm, b = fp
    else if (KIND == DS_FORWARD2)   // a = f, b = fp, c = fpp:  alpha*bphi*(-3 f + 4 fp - fpp)/2/delta
        v = __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(alpha, b0), __dsub_rn(__dadd_rn(__dmul_rn(-3., a), __dmul_rn(4., b)), c)), 2.), delta);
    else if (KIND == DS_BACKWARD2)  // a = f, b = fm, c = fmm:  alpha*bphi*(3 f - 4 fm + fmm)/2/delta
        v = __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(alpha, b0), __dadd_rn(__dsub_rn(__dmul_rn(3., a), __dmul_rn(4., b)), c)), 2.), delta);
    else {                          // DSSCentered (ds.h:97-106): a = fm, b = f, c = fp
        double bP2 = __ddiv_rn(__dadd_rn(bp, b0), 2.), bM2 = __ddiv_rn(__dadd_rn(bm, b0), 2.);
        double fm2 = __ddiv_rn(__dsub_rn(b, a), delta), fp2 = __ddiv_rn(__dsub_rn(c, b), delta);
        v = __ddiv_rn(__dmul_rn(__dmul_rn(alpha, b0), __dsub_rn(__dmul_rn(bP2, fp2), __dmul_rn(bM2, fm2))), delta);
    }
    return beta == 0. ? v : __dadd_rn(v, __dmul_rn(beta, g));
}

// the formulas that also need the volume form sqrtG on the three planes (ds.h:111-135, 903-1000)
template <int KIND>
__device__ __forceinline__ double ds_vol_formula(double alpha, double beta, double delta, double g, double a, double b, double c,
                                                  double Gm, double G0, double Gp, double bm, double b0, double bp) {
    double v;
    if (KIND == DSSD_CENTERED) {    // DSSDCentered: a = fm, b = f, c = fp
        double bP2 = __ddiv_rn(__dadd_rn(bp, b0), 2.), bM2 = __ddiv_rn(__dadd_rn(bm, b0), 2.);
        double fm2 = __ddiv_rn(__dsub_rn(b, a), delta), fp2 = __ddiv_rn(__dsub_rn(c, b), delta);
        double gp2 = __ddiv_rn(__ddiv_rn(__dadd_rn(Gp, G0), G0), 2.), gm2 = __ddiv_rn(__ddiv_rn(__dadd_rn(Gm, G0), G0), 2.);
        double t1 = __dmul_rn(__dmul_rn(__dmul_rn(gp2, fp2), bP2), bP2), t2 = __dmul_rn(__dmul_rn(__dmul_rn(bM2, bM2), gm2), fm2);
        v = __ddiv_rn(__dmul_rn(alpha, __dsub_rn(t1, t2)), delta);
    } else if (KIND == DS_DIV_BACKWARD)  // a = fm, b = f: alpha*(bP0*G0*f - bPm*Gm*fm)/G0/delta
        v = __ddiv_rn(__ddiv_rn(__dmul_rn(alpha, __dsub_rn(__dmul_rn(__dmul_rn(b0, G0), b), __dmul_rn(__dmul_rn(bm, Gm), a))), G0), delta);
    else if (KIND == DS_DIV_FORWARD)     // a = f, b = fp: alpha*(bPp*Gp*fp - bP0*G0*f)/G0/delta
        v = __ddiv_rn(__ddiv_rn(__dmul_rn(alpha, __dsub_rn(__dmul_rn(__dmul_rn(bp, Gp), b), __dmul_rn(__dmul_rn(b0, G0), a))), G0), delta);
    else if (KIND == DS_DIV_CENTERED)    // a = fm, b = fp: alpha*(fp*Gp*bPp - fm*Gm*bPm)/G0/2/delta
        v = __ddiv_rn(__ddiv_rn(__ddiv_rn(__dmul_rn(alpha, __dsub_rn(__dmul_rn(__dmul_rn(b, Gp), bp), __dmul_rn(__dmul_rn(a, Gm), bm))), G0), 2.), delta);
    else                                 // ds_average: a = fm, b = fp: alpha*(fp+fm)/2
        v = __ddiv_rn(__dmul_rn(alpha, __dadd_rn(b, a)), 2.);
    return beta == 0. ? v : __dadd_rn(v, __dmul_rn(beta, g));
}

template <int KIND>
__global__ void __launch_bounds__(256)
ds_vol_kernel(size_t n, double alpha, double beta, double delta, const double* __restrict__ a, const double* __restrict__ b,
              const double* __restrict__ c, const double* __restrict__ Gm, const double* __restrict__ G0,
              const double* __restrict__ Gp, const double* __restrict__ bm, const double* __restrict__ b0,
              const double* __restrict__ bp, double* __restrict__ g) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double cc = c ? c[i] : 0.;
        double go = beta == 0. ? 0. : g[i];
        g[i] = ds_vol_formula<KIND>(alpha, beta, delta, go, a[i], b[i], cc, Gm ? Gm[i] : 0., G0 ? G0[i] : 1., Gp ? Gp[i] : 0.,
                                    bm ? bm[i] : 0., b0 ? b0[i] : 0., bp ? bp[i] : 0.);
    }
}

template <int KIND>
__global__ void __launch_bounds__(256)
ds_kernel(size_t n, double alpha, double beta, double delta, const double* __restrict__ a, const double* __restrict__ b,
          const double* __restrict__ c, const double* __restrict__ bm, const double* __restrict__ b0,
          const double* __restrict__ bp, double* __restrict__ g) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double cc = c ? c[i] : 0., vm = bm ? bm[i] : 0., vp = bp ? bp[i] : 0.;
        double go = beta == 0. ? 0. : g[i];
        g[i] = ds_formula<KIND>(alpha, beta, delta, go, a[i], b[i], cc, vm, b0[i], vp);
    }
}

// fused DS::centered for periodic z: thread = one row of the 2-d plane and PL consecutive planes
template <int PL>
__global__ void __launch_bounds__(128)
ds_centered_fused_kernel(int num_rows, int nplanes, const int* __restrict__ ppos, const int* __restrict__ pidx,
                         const double* __restrict__ pval, const int* __restrict__ mpos, const int* __restrict__ midx,
                         const double* __restrict__ mval, double alpha, const double* __restrict__ f,
                         const double* __restrict__ bphi, double delta, double beta, double* __restrict__ g) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int p0 = blockIdx.y * PL;
    if (row >= num_rows) return;
    double fp[PL], fm[PL];
    const double *xp[PL], *xm[PL];
#pragma unroll
    for (int p = 0; p < PL; p++) {
        int pl = min(p0 + p, nplanes - 1);
        int up = pl + 1 == nplanes ? 0 : pl + 1, dn = pl == 0 ? nplanes - 1 : pl - 1;
        xp[p] = f + (size_t)up * num_rows;
        xm[p] = f + (size_t)dn * num_rows;
        fp[p] = 0.;
        fm[p] = 0.;
    }
    // same summation order as the CSR symv with alpha = 1, beta = 0 (sparsematrix_omp.h:39-48)
    for (int jj = ppos[row], e = ppos[row + 1]; jj < e; jj++) {
        const double v = __dmul_rn(1., __ldg(pval + jj));
        const int j = __ldg(pidx + jj);
#pragma unroll
        for (int p = 0; p < PL; p++) fp[p] = __fma_rn(v, __ldg(xp[p] + j), fp[p]);
    }
    for (int jj = mpos[row], e = mpos[row + 1]; jj < e; jj++) {
        const double v = __dmul_rn(1., __ldg(mval + jj));
        const int j = __ldg(midx + jj);
#pragma unroll
        for (int p = 0; p < PL; p++) fm[p] = __fma_rn(v, __ldg(xm[p] + j), fm[p]);
    }
#pragma unroll
    for (int p = 0; p < PL; p++) {
        if (p0 + p >= nplanes) break;
        const size_t i = (size_t)(p0 + p) * num_rows + row;
        double go = beta == 0. ? 0. : g[i];
        g[i] = ds_formula<DS_CENTERED>(alpha, beta, delta, go, fm[p], fp[p], 0., 0., bphi[i], 0.);
    }
}

static unsigned ew_grid(size_t n) {
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    return (unsigned)(want < cap ? want : cap);
}

// assign_bc_along_field_2nd / _1st and swap_bc_perp (ds.h:169-330): ghost values of the minus / plus neighbours where the field
// line leaves the domain.  The reference's functors are user lambdas; the expressions below are theirs, evaluated left to
// right with separately rounded operations (the library is compiled with -fmad=false).  fmg may alias fm, fpg may alias fp.
template <int ORDER, bool NEU>
__global__ void __launch_bounds__(256)
assign_bc_kernel(size_t n, double delta, const double* fm_, const double* __restrict__ f_, const double* fp_,
                 const double* __restrict__ hbm_, const double* __restrict__ hbp_, const double* __restrict__ bbm_,
                 const double* __restrict__ bbo_, const double* __restrict__ bbp_, double bv0, double bv1, double* fmg, double* fpg) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double fm = fm_[i], fp = fp_[i], fo = ORDER == 2 ? f_[i] : 0., hm = delta, hp = delta;
        const double bbm = bbm_[i], bbp = bbp_[i];
        double plus, minus, bothP = 0., bothM = 0.;
        if (ORDER == 1 && NEU) {  // ds.h:247-258
            const double dbm = bv0, dbp = bv1;
            plus = fm + dbp * (hp + hm);
            minus = fp - dbm * (hp + hm);
            fmg[i] = (1. - bbm) * fm + bbm * minus;
            fpg[i] = (1. - bbp) * fp + bbp * plus;
            continue;
        }
        const double hbm = hbm_[i], hbp = hbp_[i], bbo = bbo_[i];
        if (ORDER == 2 && NEU) {  // ds.h:178-200
            const double dbm = bv0, dbp = bv1;
            plus = dbp * hp * (hm + hp) / (2. * hbp + hm) + fo * (2. * hbp + hm - hp) * (hm + hp) / hm / (2. * hbp + hm) +
                   fm * hp * (-2. * hbp + hp) / hm / (2. * hbp + hm);
            minus = fp * hm * (-2. * hbm + hm) / hp / (2. * hbm + hp) - dbm * hm * (hm + hp) / (2. * hbm + hp) +
                    fo * (2. * hbm - hm + hp) * (hm + hp) / hp / (2. * hbm + hp);
            bothM = fo + dbp * hm * (-2. * hbm + hm) / 2. / (hbm + hbp) - dbm * hm * (2. * hbp + hm) / 2. / (hbm + hbp);
            bothP = fo + dbp * hp * (2. * hbm + hp) / 2. / (hbm + hbp) + dbm * hp * (2. * hbp - hp) / 2. / (hbm + hbp);
        } else if (ORDER == 2) {  // ds.h:204-225
            const double fbm = bv0, fbp = bv1;
            plus = fm * hp * (-hbp + hp) / hm / (hbp + hm) + fo * (hbp - hp) * (hm + hp) / hbp / hm + fbp * hp * (hm + hp) / hbp / (hbp + hm);
            minus = +fo * (hbm - hm) * (hm + hp) / hbm / hp + fbm * hm * (hm + hp) / hbm / (hbm + hp) + fp * hm * (-hbm + hm) / hp / (hbm + hp);
            bothM = fbp * hm * (-hbm + hm) / hbp / (hbm + hbp) + fo * (hbm - hm) * (hbp + hm) / hbm / hbp + fbm * hm * (hbp + hm) / hbm / (hbm + hbp);
            bothP = fo * (hbp - hp) * (hbm + hp) / hbm / hbp + fbp * hp * (hbm + hp) / hbp / (hbm + hbp) + fbm * hp * (-hbp + hp) / hbm / (hbm + hbp);
        } else {  // ORDER == 1, DIR: ds.h:262-277
            const double fbm = bv0, fbp = bv1;
            plus = fm + (fbp - fm) / (hbp + hm) * (hp + hm);
            minus = fp - (hp + hm) * (fp - fbm) / (hp + hbm);
            bothM = fbp + (fbp - fbm) / (hbp + hbm) * (hp + hbm);
            bothP = fbp - (fbp - fbm) / (hbp + hbm) * (hbp + hm);
        }
        fmg[i] = (1. - bbo - bbm) * fm + bbm * minus + bbo * bothM;
        fpg[i] = (1. - bbo - bbp) * fp + bbp * plus + bbo * bothP;
    }
}
// swap_bc_perp (ds.h:307-318): the values outside the box change sign
__global__ void __launch_bounds__(256)
swap_bc_perp_kernel(size_t n, const double* fm_, const double* fp_, const double* __restrict__ bbm_, const double* __restrict__ bbo_,
                    const double* __restrict__ bbp_, double* fmg, double* fpg) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double fm = fm_[i], fp = fp_[i], bbm = bbm_[i], bbo = bbo_[i], bbp = bbp_[i];
        fmg[i] = (1. - bbo - bbm) * fm + (bbm + bbo) * (-fm);
        fpg[i] = (1. - bbo - bbp) * fp + (bbp + bbo) * (-fp);
    }
}

}  // namespace dgb

using namespace dgb;

extern "C" {
// kind: 0 forward (a=f,b=fp) 1 backward (a=f,b=fm) 2 centered (a=fm,b=fp) 3 forward2 (a=f,b=fp,c=fpp)
//       4 backward2 (a=f,b=fm,c=fmm) 5 dss_centered (a=fm,b=f,c=fp; bphi_m, bphi, bphi_p)
int dgb_ds_apply(int kind, size_t n, double alpha, const double* a, const double* b, const double* c, const double* bphi_m,
                 const double* bphi, const double* bphi_p, double delta_phi, double beta, double* g, dgb_stream_t s) {
    if (n == 0) return 0;
    if (!a || !b || !bphi || !g || (kind >= 3 && !c) || (kind == 5 && (!bphi_m || !bphi_p))) {
        set_error("dgb_ds_apply: missing operand for kind %d", kind);
        return DGB_ERR_INVALID;
    }
    cudaStream_t st = as_stream(s);
    unsigned grid = ew_grid(n);
    switch (kind) {
        case DS_FORWARD: ds_kernel<DS_FORWARD><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, bphi_m, bphi, bphi_p, g); break;
        case DS_BACKWARD: ds_kernel<DS_BACKWARD><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, bphi_m, bphi, bphi_p, g); break;
        case DS_CENTERED: ds_kernel<DS_CENTERED><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, bphi_m, bphi, bphi_p, g); break;
        case DS_FORWARD2: ds_kernel<DS_FORWARD2><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, bphi_m, bphi, bphi_p, g); break;
        case DS_BACKWARD2: ds_kernel<DS_BACKWARD2><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, bphi_m, bphi, bphi_p, g); break;
        case DSS_CENTERED: ds_kernel<DSS_CENTERED><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, bphi_m, bphi, bphi_p, g); break;
        default: set_error("dgb_ds_apply: unknown kind %d", kind); return DGB_ERR_INVALID;
    }
    DGB_LAUNCHED();
    return 0;
}

// kind: 6 dssd_centered (a=fm,b=f,c=fp) 7 ds_divBackward (a=fm,b=f) 8 ds_divForward (a=f,b=fp) 9 ds_divCentered (a=fm,b=fp)
//       10 ds_average (a=fm,b=fp; no metric fields);  ds.h:881-1000
int dgb_ds_apply_vol(int kind, size_t n, double alpha, const double* a, const double* b, const double* c, const double* sqrtG_m,
                     const double* sqrtG, const double* sqrtG_p, const double* bphi_m, const double* bphi, const double* bphi_p,
                     double delta_phi, double beta, double* g, dgb_stream_t s) {
    if (n == 0) return 0;
    bool ok = a && b && g;
    if (kind == DSSD_CENTERED) ok = ok && c && sqrtG_m && sqrtG && sqrtG_p && bphi_m && bphi && bphi_p;
    else if (kind == DS_DIV_BACKWARD) ok = ok && sqrtG_m && sqrtG && bphi_m && bphi;
    else if (kind == DS_DIV_FORWARD) ok = ok && sqrtG_p && sqrtG && bphi_p && bphi;
    else if (kind == DS_DIV_CENTERED) ok = ok && sqrtG_m && sqrtG && sqrtG_p && bphi_m && bphi_p;
    else if (kind != DS_AVERAGE) { set_error("dgb_ds_apply_vol: unknown kind %d", kind); return DGB_ERR_INVALID; }
    if (!ok) { set_error("dgb_ds_apply_vol: missing operand for kind %d", kind); return DGB_ERR_INVALID; }
    cudaStream_t st = as_stream(s);
    unsigned grid = ew_grid(n);
#define DGB_DSV(K) case K: ds_vol_kernel<K><<<grid, 256, 0, st>>>(n, alpha, beta, delta_phi, a, b, c, sqrtG_m, sqrtG, sqrtG_p, bphi_m, bphi, bphi_p, g); break;
    switch (kind) { DGB_DSV(DSSD_CENTERED) DGB_DSV(DS_DIV_BACKWARD) DGB_DSV(DS_DIV_FORWARD) DGB_DSV(DS_DIV_CENTERED) DGB_DSV(DS_AVERAGE) }
#undef DGB_DSV
    DGB_LAUNCHED();
    return 0;
}

// Fieldaligned::ePlus (plus != 0) / eMinus: out[k] = M f[k+1] resp. M f[k-1] on all planes, then for bcz != PER the
// ghost-cell fix-up of the last / first plane (fieldaligned.h:868-879, 898-909).  bnd = m_right resp. m_left,
// limiter, ghost: 2-d fields of num_rows elements (ghost is scratch); unused for bcz == DGB_PER.
int dgb_fa_shift(int plus, int num_rows, int nplanes, const int* pos, const int* idx, const double* val, const double* f,
                 double* out, int bcz, const double* bnd, const double* limiter, double* ghost, double delta_phi, dgb_stream_t s) {
    int e = dgb_csr_spmv_planes(num_rows, num_rows, pos, idx, val, 1., f, 0., out, nplanes, plus ? 1 : -1, s);
    if (e || bcz == DGB_PER || nplanes == 0) return e;
    if (!bnd || !limiter || !ghost) { set_error("dgb_fa_shift: boundary fields required for non-periodic z"); return DGB_ERR_INVALID; }
    const size_t n = num_rows;
    const int i0 = plus ? nplanes - 1 : 0;
    const double* fi = f + (size_t)i0 * n;
    double* ti = out + (size_t)i0 * n;
    const bool dir = plus ? (bcz == DGB_DIR || bcz == DGB_NEU_DIR) : (bcz == DGB_DIR || bcz == DGB_DIR_NEU);
    if (dir) e = dgb_axpbyz(n, 2., bnd, -1., fi, ghost, s);
    else e = dgb_axpbyz(n, plus ? delta_phi : -delta_phi, bnd, 1., fi, ghost, s);
    if (e) return e;
    if ((e = dgb_axpby(n, -1., ti, 1., ghost, s))) return e;       // ghost = 1*ghost - 1*temp  (axpby(1,ghost,-1,temp,ghost))
    return dgb_pointwise_dot(n, 1., limiter, ghost, 1., ti, s);    // temp += limiter * ghost
}

int dgb_assign_bc_along_field(int order, int bc, size_t n, double delta_phi, const double* fm, const double* f, const double* fp,
                              const double* hbm, const double* hbp, const double* bbm, const double* bbo, const double* bbp, double bv_minus,
                              double bv_plus, double* fmg, double* fpg, dgb_stream_t s) {
    if (order != 1 && order != 2) { set_error("dgb_assign_bc_along_field: order must be 1 or 2"); return DGB_ERR_INVALID; }
    if (bc != DGB_NEU && bc != DGB_DIR) { set_error("dgb_assign_bc_along_field: only dg::NEU and dg::DIR exist (ds.h:173,201)"); return DGB_ERR_UNSUPPORTED; }
    const bool neu = bc == DGB_NEU;
    if (!fm || !fp || !fmg || !fpg || !bbm || !bbp || (order == 2 && !f) || (!(order == 1 && neu) && (!hbm || !hbp || !bbo))) {
        set_error("dgb_assign_bc_along_field: missing operand");
        return DGB_ERR_INVALID;
    }
    if (n == 0) return 0;
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    cudaStream_t st = as_stream(s);
    if (order == 2 && neu) assign_bc_kernel<2, true><<<grid, 256, 0, st>>>(n, delta_phi, fm, f, fp, hbm, hbp, bbm, bbo, bbp, bv_minus, bv_plus, fmg, fpg);
    else if (order == 2) assign_bc_kernel<2, false><<<grid, 256, 0, st>>>(n, delta_phi, fm, f, fp, hbm, hbp, bbm, bbo, bbp, bv_minus, bv_plus, fmg, fpg);
    else if (neu) assign_bc_kernel<1, true><<<grid, 256, 0, st>>>(n, delta_phi, fm, f, fp, hbm, hbp, bbm, bbo, bbp, bv_minus, bv_plus, fmg, fpg);
    else assign_bc_kernel<1, false><<<grid, 256, 0, st>>>(n, delta_phi, fm, f, fp, hbm, hbp, bbm, bbo, bbp, bv_minus, bv_plus, fmg, fpg);
    DGB_LAUNCHED();
    return 0;
}
int dgb_swap_bc_perp(size_t n, const double* fm, const double* fp, const double* bbm, const double* bbo, const double* bbp, double* fmg,
                     double* fpg, dgb_stream_t s) {
    if (n == 0) return 0;
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    swap_bc_perp_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(s)>>>(n, fm, fp, bbm, bbo, bbp, fmg, fpg);
    DGB_LAUNCHED();
    return 0;
}

// DS::centered(alpha, f, beta, g) for periodic z in one launch (ds.h:481-485)
int dgb_ds_centered_fused(int num_rows, int nplanes, const int* plus_pos, const int* plus_idx, const double* plus_val,
                          const int* minus_pos, const int* minus_idx, const double* minus_val, double alpha, const double* f,
                          const double* bphi, double delta_phi, double beta, double* g, dgb_stream_t s) {
    if (num_rows <= 0 || nplanes <= 0) return 0;
    if (f == g) { set_error("dgb_ds_centered_fused: f must not alias g"); return DGB_ERR_INVALID; }
    dim3 block(128), grid((num_rows + 127) / 128, (nplanes + 3) / 4);
    ds_centered_fused_kernel<4><<<grid, block, 0, as_stream(s)>>>(num_rows, nplanes, plus_pos, plus_idx, plus_val, minus_pos,
                                                                minus_idx, minus_val, alpha, f, bphi, delta_phi, beta, g);
    DGB_LAUNCHED();
    return 0;
}
}
