// dg::Elliptic2d: y = alpha/vol [ -Lx sigma (chi_xx Rx + chi_xy Ry) - Ly sigma (chi_yx Rx + chi_yy Ry) + jfactor (Jx + Jy) ] x + beta y
// Replaces the 8-launch composition of Elliptic2d::symv (inc/dg/elliptic.h:428-458).  Three implementations, chosen per call:
//  * elliptic2d_walker_kernel (elliptic_walker.cu), the default from ~400^2 cells on: one warp per strip of cell columns
//    walking along y, warp-private TMA rings, fluxes in registers / shuffles, TMA store; ONE pass, 24 B/dof of HBM traffic
//    (+8 if beta != 0, +8 vol).  Also carries the fused exact dot of PCG, the Helmholtz epilogue and the slab (multi-GPU) mode.
//  * elliptic2d_fused_kernel (elliptic_fused.cu), small grids and n = 4: a CTA owns a tile of 32 x 8 cells staged by TMA,
//    fluxes through shared memory, one thread per cell; same traffic, same options.
//  * elliptic2d_symv_unfused (this file): the reference's composition on our kernels (6 Ell symv + tensor multiply +
//    divide) -- the general path (any chi tensor, chi-weighted jumps, any matrix structure) and the test comparator
//    (every fused result is checked bitwise against it).
// The n x n blocks of interior rows are kernel parameters (uniform / constant-bank operands of the DFMAs); boundary cells
// look their blocks up in global memory.  This file also holds the GeneralHelmholtz mode (helmholtz.h:74-80) and
// Elliptic::variation (elliptic.h:497-502).
#include "elliptic.cuh"
#include <cstdlib>

namespace dgb {

// ------------------------------------------------------------------------------------------------ unfused
extern "C" int dgb_tensor_multiply2d(size_t, const double*, double, const double*, const double*, const double*,
                                     const double*, const double*, const double*, double, double*, double*, dgb_stream_t);
extern "C" int dgb_axpbypgz(size_t, double, const double*, double, const double*, double, double*, dgb_stream_t);
extern "C" int dgb_tensor_dot2d(size_t, double, const double*, double, const double*, const double*, const double*, const double*,
                                const double*, const double*, const double*, double, const double*, const double*, double, double*,
                                dgb_stream_t);

__global__ void __launch_bounds__(256)
elliptic_finish_kernel(size_t n, double alpha, const double* __restrict__ temp, const double* __restrict__ vol,
                       double beta, double* __restrict__ y) {
    // pointwiseDivide(alpha, temp, vol, beta, y) (elliptic.h:458, functor subroutines.h:376); beta == 0 does not
    // read y (the only deliberate deviation -- NaN in y is overwritten instead of propagated)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double q = vol ? __ddiv_rn(temp[i], vol[i]) : temp[i];
        double t = beta == 0. ? 0. : __dmul_rn(y[i], beta);
        y[i] = __fma_rn(alpha, q, t);
    }
}

static int ensure_temps(Elliptic2dPlan& p) {
    if (p.tx) return 0;
    DGB_CUDA(cudaMalloc(&p.tx, p.size * sizeof(double)));
    DGB_CUDA(cudaMalloc(&p.ty, p.size * sizeof(double)));
    DGB_CUDA(cudaMalloc(&p.t, p.size * sizeof(double)));
    return 0;
}

static int elliptic2d_unfused(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    int e = ensure_temps(p);
    if (e) return e;
    dgb_stream_t s = reinterpret_cast<dgb_stream_t>(st);
    const size_t n = p.size;
    if ((e = ell_symv(p.rightx, 1., x, 0., p.tx, st, false))) return e;                    // elliptic.h:431
    if ((e = ell_symv(p.righty, 1., x, 0., p.ty, st, false))) return e;                    // :432
    if ((e = dgb_tensor_multiply2d(n, p.sigma, 1., p.chi[0], p.chi[1], p.chi[2], p.chi[3], p.tx, p.ty, 0., p.tx, p.ty, s))) return e;  // :435
    if ((e = ell_symv(p.lefty, 1., p.ty, 0., p.t, st, false))) return e;                   // :438
    if ((e = ell_symv(p.leftx, -1., p.tx, -1., p.t, st, false))) return e;                 // :439
    if (p.jfactor != 0.) {                                                                 // :442
        if (p.chi_weight_jump) {
            if ((e = ell_symv(p.jumpx, p.jfactor, x, 0., p.tx, st, false))) return e;
            if ((e = ell_symv(p.jumpy, p.jfactor, x, 0., p.ty, st, false))) return e;
            if ((e = dgb_tensor_multiply2d(n, p.sigma, 1., p.chi[0], p.chi[1], p.chi[2], p.chi[3], p.tx, p.ty, 0., p.tx, p.ty, s))) return e;
            if ((e = dgb_axpbypgz(n, 1., p.tx, 1., p.ty, 1., p.t, s))) return e;
        } else {
            if ((e = ell_symv(p.jumpx, p.jfactor, x, 1., p.t, st, false))) return e;       // :454
            if ((e = ell_symv(p.jumpy, p.jfactor, x, 1., p.t, st, false))) return e;       // :455
        }
    }
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    elliptic_finish_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(n, alpha, p.t, p.vol, beta, y);  // :458
    DGB_LAUNCHED();
    return 0;
}

extern "C" int dgb_pointwise_dot(size_t, double, const double*, const double*, double, double*, dgb_stream_t);
extern "C" int dgb_axpby(size_t, double, const double*, double, double*, dgb_stream_t);
static int elliptic2d_symv_plain(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                                 bool force_unfused);

static bool env_unfused_helm() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DGB_ELLIPTIC_UNFUSED"); v = (e && atoi(e)) ? 1 : 0; }
    return v != 0;
}
// GeneralHelmholtz::symv (helmholtz.h:74-80): only the two-operand form exists in the reference
int elliptic2d_symv(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                    bool force_unfused) {
    if (!p.helm) return elliptic2d_symv_plain(p, alpha, x, beta, y, st, force_unfused);
    if (alpha != 1. || beta != 0.) {
        set_error("dgb_elliptic2d_symv: a Helmholtz plan supports symv(x, y) only (alpha = 1, beta = 0), as the reference");
        return DGB_ERR_UNSUPPORTED;
    }
    if (p.kernel_mode == DGB_ELLIPTIC_KERNEL_UNFUSED && !p.slab) force_unfused = true;
    if (!force_unfused && !env_unfused_helm() && p.fusable && p.helm_alpha != 0. && !p.chi[0] && !p.chi[1] && !p.chi[2] && !p.chi[3] &&
        !p.chi_weight_jump && p.sigma && x != y)
        return elliptic2d_fused_launch(p, 1., x, 0., y, st);  // both fused kernels apply the Helmholtz epilogue themselves
    int e = 0;
    if (p.helm_alpha != 0.) { if ((e = elliptic2d_symv_plain(p, 1., x, 0., y, st, true))) return e; }
    dgb_stream_t s = reinterpret_cast<dgb_stream_t>(st);
    const size_t n = p.slab ? (size_t)p.slab_rows * p.n * p.Nx * p.n : p.size;
    if (p.helm_chi) return dgb_pointwise_dot(n, 1., p.helm_chi, x, -p.helm_alpha, y, s);
    // chi = 1: z *= b; z = fma(1*1, x, z)
    return dgb_axpby(n, 1., x, -p.helm_alpha, y, s);
}

static int elliptic2d_symv_plain(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                                 bool force_unfused) {
    if (!p.sigma) { set_error("dgb_elliptic2d_symv: sigma has not been set"); return DGB_ERR_INVALID; }
    if (x == y) { set_error("dgb_elliptic2d_symv: x must not alias y"); return DGB_ERR_INVALID; }
    static int env_unfused = -1;
    if (env_unfused < 0) { const char* e = getenv("DGB_ELLIPTIC_UNFUSED"); env_unfused = (e && atoi(e)) ? 1 : 0; }
    bool identity_chi = !p.chi[0] && !p.chi[1] && !p.chi[2] && !p.chi[3];
    if (p.slab) {
        if (!p.fusable || !identity_chi || p.chi_weight_jump) {
            set_error("dgb_elliptic2d_symv: a slab (multi-GPU) plan needs the fused kernel (dx.h matrices, identity chi tensor)");
            return DGB_ERR_UNSUPPORTED;
        }
        return elliptic2d_fused_launch(p, alpha, x, beta, y, st);
    }
    if (p.kernel_mode == DGB_ELLIPTIC_KERNEL_UNFUSED) force_unfused = true;
    if (!force_unfused && !env_unfused && p.fusable && identity_chi && !p.chi_weight_jump)
        return elliptic2d_fused_launch(p, alpha, x, beta, y, st);
    return elliptic2d_unfused(p, alpha, x, beta, y, st);
}

// structure check for the fused kernel: near-diagonal pattern with |offset| <= 1, boundary slots either padding or
// the (periodically wrapped) neighbour
static bool dx_like(const EllDev& m, int expect_bpl, bool& wrap) {
    if (!m.has_pattern || m.bpl != expect_bpl) return false;
    wrap = false;
    for (int d = 0; d < m.bpl; d++)
        if (m.off[d] < -1 || m.off[d] > 1) return false;
    if (m.num_rows != m.num_cols) return false;
    for (int i = 0; i < m.num_rows; i++) {
        if (i >= m.i_lo && i < m.i_hi) continue;
        for (int d = 0; d < m.bpl; d++) {
            int c = m.h_cols[(size_t)i * m.bpl + d];
            if (c == -1) continue;
            int o = c - i;  // boundary rows may assign slots to neighbours differently from interior rows
            if (o > 1) { o -= m.num_rows; wrap = true; }
            else if (o < -1) { o += m.num_rows; wrap = true; }
            if (o < -1 || o > 1) return false;
        }
    }
    return true;
}

static void analyse(Elliptic2dPlan& p) {
    p.fusable = false;
    const int n = p.rightx.n;
    if (n < 2 || n > 4) return;
    EllDev* xm[3] = {&p.rightx, &p.leftx, &p.jumpx};
    EllDev* ym[3] = {&p.righty, &p.lefty, &p.jumpy};
    int bder = p.rightx.bpl;
    if (bder != 2 && bder != 3) return;
    bool wx[3], wy[3];
    for (int k = 0; k < 3; k++) {
        int eb = k == 2 ? 3 : bder;
        if (!dx_like(*xm[k], eb, wx[k]) || !dx_like(*ym[k], eb, wy[k])) return;
        if (xm[k]->n != n || ym[k]->n != n) return;
        if (xm[k]->right != 1 || ym[k]->left != 1) return;
        if (xm[k]->rr0 != 0 || xm[k]->rr1 != 1 || ym[k]->rr0 != 0 || ym[k]->rr1 != ym[k]->right) return;
    }
    const int Nx = p.rightx.num_rows, Ny = p.righty.num_rows;
    for (int k = 0; k < 3; k++) {
        if (xm[k]->num_rows != Nx || ym[k]->num_rows != Ny) return;
        if (xm[k]->left != Ny * n || ym[k]->right != Nx * n) return;
    }
    if (Nx < 5 || Ny < 5) return;
    // a periodic family wraps in all three matrices; a non-periodic one in none.  (A matrix whose stencil never
    // leaves the domain at a boundary row -- e.g. forward dx at row 0 -- reports no wrap there, so compare on the
    // jump matrix, which reaches both sides.)
    p.wrapx = wx[2]; p.wrapy = wy[2];
    for (int k = 0; k < 2; k++) {
        if (wx[k] && !p.wrapx) return;
        if (wy[k] && !p.wrapy) return;
    }
    // the fused kernel hard-wires the interior stencil offsets of dx.h: right {0,+1} / left {-1,0} (forward),
    // right {-1,0} / left {0,+1} (backward), {-1,0,+1} for centered derivatives and for the jumps
    auto offs = [](const EllDev& m, int first) {
        for (int d = 0; d < m.bpl; d++)
            if (m.off[d] != first + d) return false;
        return true;
    };
    int dirk;
    if (bder == 3) dirk = 2;
    else if (offs(p.rightx, 0)) dirk = 0;
    else dirk = 1;
    const int rfirst = dirk == 0 ? 0 : -1, lfirst = dirk == 1 ? 0 : -1;
    if (!offs(p.rightx, rfirst) || !offs(p.righty, rfirst) || !offs(p.leftx, lfirst) || !offs(p.lefty, lfirst) ||
        !offs(p.jumpx, -1) || !offs(p.jumpy, -1))
        return;
    p.n = n; p.Nx = Nx; p.Ny = Ny; p.bder = bder; p.dirk = dirk;
    p.fusable = true;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_elliptic2d_create(dgb_elliptic2d** out, const dgb_ell_host* leftx, const dgb_ell_host* lefty,
                          const dgb_ell_host* rightx, const dgb_ell_host* righty, const dgb_ell_host* jumpx,
                          const dgb_ell_host* jumpy, double jfactor, int chi_weight_jump) {
    Elliptic2dPlan* p = new Elliptic2dPlan();
    const dgb_ell_host* hs[6] = {leftx, lefty, rightx, righty, jumpx, jumpy};
    EllDev* ds[6] = {&p->leftx, &p->lefty, &p->rightx, &p->righty, &p->jumpx, &p->jumpy};
    int e = 0;
    for (int k = 0; k < 6 && !e; k++) e = ell_upload(*ds[k], hs[k]);
    if (!e) {
        p->size = p->rightx.total_cols();
        for (int k = 0; k < 6; k++)
            if (ds[k]->total_rows() != p->size || ds[k]->total_cols() != p->size) {
                set_error("dgb_elliptic2d_create: matrix %d is not square of size %zu", k, p->size);
                e = DGB_ERR_INVALID;
            }
    }
    if (e) { for (int k = 0; k < 6; k++) ell_release(*ds[k]); delete p; return e; }
    p->jfactor = jfactor;
    p->chi_weight_jump = chi_weight_jump != 0;
    analyse(*p);
    *out = reinterpret_cast<dgb_elliptic2d*>(p);
    return 0;
}
int dgb_elliptic2d_destroy(dgb_elliptic2d* h) {
    Elliptic2dPlan* p = reinterpret_cast<Elliptic2dPlan*>(h);
    if (!p) return 0;
    EllDev* ds[6] = {&p->leftx, &p->lefty, &p->rightx, &p->righty, &p->jumpx, &p->jumpy};
    for (int k = 0; k < 6; k++) ell_release(*ds[k]);
    cudaFree(p->tx); cudaFree(p->ty); cudaFree(p->t);
    elliptic2d_walker_release(*p);
    delete p;
    return 0;
}
int dgb_elliptic2d_set_sigma(dgb_elliptic2d* h, const double* sigma) { reinterpret_cast<Elliptic2dPlan*>(h)->sigma = sigma; return 0; }
int dgb_elliptic2d_set_vol(dgb_elliptic2d* h, const double* vol) { reinterpret_cast<Elliptic2dPlan*>(h)->vol = vol; return 0; }
int dgb_elliptic2d_set_chi(dgb_elliptic2d* h, const double* xx, const double* xy, const double* yx, const double* yy) {
    Elliptic2dPlan* p = reinterpret_cast<Elliptic2dPlan*>(h);
    p->chi[0] = xx; p->chi[1] = xy; p->chi[2] = yx; p->chi[3] = yy;
    return 0;
}
int dgb_elliptic2d_set_helmholtz(dgb_elliptic2d* h, int enable, double alpha, const double* chi) {
    Elliptic2dPlan* p = reinterpret_cast<Elliptic2dPlan*>(h);
    p->helm = enable != 0; p->helm_alpha = alpha; p->helm_chi = chi;
    return 0;
}
int dgb_elliptic2d_variation(dgb_elliptic2d* h, double alpha, const double* lambda, const double* phi, double beta, double* sigma,
                             dgb_stream_t s) {
    Elliptic2dPlan& p = *reinterpret_cast<Elliptic2dPlan*>(h);
    if (p.slab) { set_error("dgb_elliptic2d_variation: not available on a slab plan"); return DGB_ERR_UNSUPPORTED; }
    int e = ensure_temps(p);
    if (e) return e;
    cudaStream_t st = as_stream(s);
    if ((e = ell_symv(p.rightx, 1., phi, 0., p.tx, st, false))) return e;   // elliptic.h:499
    if ((e = ell_symv(p.righty, 1., phi, 0., p.ty, st, false))) return e;   // :500
    return dgb_tensor_dot2d(p.size, alpha, lambda, 1., p.tx, p.ty, p.chi[0], p.chi[1], p.chi[2], p.chi[3], lambda, 1., p.tx, p.ty,
                            beta, sigma, s);                                 // :501
}
int dgb_elliptic2d_set_jfactor(dgb_elliptic2d* h, double jfactor) { reinterpret_cast<Elliptic2dPlan*>(h)->jfactor = jfactor; return 0; }
int dgb_elliptic2d_set_slab(dgb_elliptic2d* h, int yoff, int rows, int ghost) {
    Elliptic2dPlan* p = reinterpret_cast<Elliptic2dPlan*>(h);
    if (!p->fusable) { set_error("dgb_elliptic2d_set_slab: the plan's matrices are not supported by the fused kernel"); return DGB_ERR_UNSUPPORTED; }
    if (yoff < 0 || rows < 1 || yoff + rows > p->Ny || ghost < 0) { set_error("dgb_elliptic2d_set_slab: invalid slab"); return DGB_ERR_INVALID; }
    p->slab = true; p->slab_yoff = yoff; p->slab_rows = rows; p->slab_ghost = ghost;
    return 0;
}
int dgb_elliptic2d_set_kernel(dgb_elliptic2d* h, int kernel) {
    Elliptic2dPlan* p = reinterpret_cast<Elliptic2dPlan*>(h);
    if (kernel < DGB_ELLIPTIC_KERNEL_AUTO || kernel > DGB_ELLIPTIC_KERNEL_UNFUSED) { set_error("dgb_elliptic2d_set_kernel: unknown kernel %d", kernel); return DGB_ERR_INVALID; }
    if ((kernel == DGB_ELLIPTIC_KERNEL_TILE || kernel == DGB_ELLIPTIC_KERNEL_WALKER) && !p->fusable) {
        set_error("dgb_elliptic2d_set_kernel: the plan's matrices do not have the dx.h structure of the fused kernels"); return DGB_ERR_UNSUPPORTED;
    }
    if (kernel == DGB_ELLIPTIC_KERNEL_WALKER && !(p->n == 2 || p->n == 3)) {
        set_error("dgb_elliptic2d_set_kernel: the walker kernel exists for n = 2, 3 (n = %d)", p->n); return DGB_ERR_UNSUPPORTED;
    }
    if (kernel == DGB_ELLIPTIC_KERNEL_WALKER && (p->Nx < 5 || p->Ny < 5)) {
        set_error("dgb_elliptic2d_set_kernel: the walker kernel needs at least 5 x 5 cells"); return DGB_ERR_UNSUPPORTED;
    }
    p->kernel_mode = kernel;
    return 0;
}
int dgb_elliptic2d_set_ordering(dgb_elliptic2d* h, int ordering) {
    Elliptic2dPlan* p = reinterpret_cast<Elliptic2dPlan*>(h);
    if (ordering != DGB_ORDER_REFERENCE && ordering != DGB_ORDER_RELAXED) { set_error("dgb_elliptic2d_set_ordering: unknown ordering %d", ordering); return DGB_ERR_INVALID; }
    p->relaxed = ordering == DGB_ORDER_RELAXED;
    return 0;
}
int dgb_elliptic2d_get_kernel(const dgb_elliptic2d* h, int with_dot, int* kernel) {
    const Elliptic2dPlan* p = reinterpret_cast<const Elliptic2dPlan*>(h);
    const bool identity_chi = !p->chi[0] && !p->chi[1] && !p->chi[2] && !p->chi[3];
    const bool fused = p->fusable && identity_chi && !p->chi_weight_jump && p->kernel_mode != DGB_ELLIPTIC_KERNEL_UNFUSED &&
                       (p->slab || !getenv("DGB_ELLIPTIC_UNFUSED"));
    *kernel = !fused ? DGB_ELLIPTIC_KERNEL_UNFUSED
                     : (elliptic2d_walker_supported(*p, with_dot != 0) ? DGB_ELLIPTIC_KERNEL_WALKER : DGB_ELLIPTIC_KERNEL_TILE);
    return 0;
}
int dgb_elliptic2d_size(const dgb_elliptic2d* h, size_t* size, int* fused) {
    const Elliptic2dPlan* p = reinterpret_cast<const Elliptic2dPlan*>(h);
    if (size) *size = p->slab ? (size_t)p->slab_rows * p->n * p->Nx * p->n : p->size;
    if (fused) *fused = p->fusable ? 1 : 0;
    return 0;
}
int dgb_elliptic2d_symv(dgb_elliptic2d* h, double alpha, const double* x, double beta, double* y, dgb_stream_t s) {
    return elliptic2d_symv(*reinterpret_cast<Elliptic2dPlan*>(h), alpha, x, beta, y, as_stream(s), false);
}
// Elliptic3d with set_compute_in_2d(true) (elliptic.h:557-797, the mode src/feltor/feltor.h uses): the 3-d operator is the 2-d
// one on every plane; sigma (= chi*vol) and the Helmholtz chi are 3-d fields, vol and the chi tensor are shared by the planes
int dgb_elliptic2d_symv_planes(dgb_elliptic2d* h, int nplanes, const double* sigma3d, double alpha, const double* x, double beta,
                               double* y, dgb_stream_t s) {
    Elliptic2dPlan& p = *reinterpret_cast<Elliptic2dPlan*>(h);
    if (p.slab) { set_error("dgb_elliptic2d_symv_planes: not available on a slab plan"); return DGB_ERR_UNSUPPORTED; }
    if (nplanes < 0 || !sigma3d || !x || !y) { set_error("dgb_elliptic2d_symv_planes: invalid argument"); return DGB_ERR_INVALID; }
    const double *sigma0 = p.sigma, *helm0 = p.helm_chi;
    const size_t n = (size_t)p.size;
    int e = 0;
    // all planes in one launch of the tile kernel when the plan qualifies for it (DGB_ELLIPTIC_PLANES_LOOP=1: per-plane loop)
    static int env_loop = -1;
    if (env_loop < 0) { const char* v = getenv("DGB_ELLIPTIC_PLANES_LOOP"); env_loop = (v && atoi(v)) ? 1 : 0; }
    const bool identity_chi = !p.chi[0] && !p.chi[1] && !p.chi[2] && !p.chi[3];
    const bool helm_ok = !p.helm || (alpha == 1. && beta == 0. && p.helm_alpha != 0.);
    if (!env_loop && nplanes > 1 && p.fusable && identity_chi && !p.chi_weight_jump && helm_ok && x != y &&
        (long long)nplanes * p.Ny * p.n < (1ll << 31)) {
        p.sigma = sigma3d;
        e = elliptic2d_fused_launch_planes(p, nplanes, alpha, x, beta, y, as_stream(s));
        p.sigma = sigma0;
        return e;
    }
    for (int k = 0; k < nplanes && !e; k++) {
        p.sigma = sigma3d + k * n;
        if (helm0) p.helm_chi = helm0 + k * n;
        e = elliptic2d_symv(p, alpha, x + k * n, beta, y + k * n, as_stream(s), false);
    }
    p.sigma = sigma0;
    p.helm_chi = helm0;
    return e;
}
int dgb_elliptic2d_symv_unfused(dgb_elliptic2d* h, double alpha, const double* x, double beta, double* y, dgb_stream_t s) {
    return elliptic2d_symv(*reinterpret_cast<Elliptic2dPlan*>(h), alpha, x, beta, y, as_stream(s), true);
}
}
