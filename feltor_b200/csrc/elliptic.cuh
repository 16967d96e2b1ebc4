// Internal: Elliptic2d plan shared by elliptic.cu, pcg.cu and multigrid.cu.
#pragma once
#include "ell.cuh"

namespace dgb {

struct Elliptic2dPlan {
    EllDev leftx, lefty, rightx, righty, jumpx, jumpy;
    double jfactor = 1.;
    bool chi_weight_jump = false;
    const double* sigma = nullptr;  // borrowed
    const double* vol = nullptr;    // borrowed, nullptr = 1
    const double* chi[4] = {nullptr, nullptr, nullptr, nullptr};  // xx, xy, yx, yy; nullptr = identity entry
    double *tx = nullptr, *ty = nullptr, *t = nullptr;  // owned temporaries of the unfused path
    size_t size = 0;
    int n = 0, Nx = 0, Ny = 0;
    bool fusable = false;  // matrices have the dx.h structure the fused kernel was written for
    int bder = 0;          // blocks per line of the four derivative matrices (2 or 3)
    int dirk = 0;          // stencil kind of the right derivatives: 0 {0,+1} forward, 1 {-1,0} backward, 2 {-1,0,+1}
    bool wrapx = false, wrapy = false;
    // slab of a y-decomposed global operator: rows [slab_yoff, slab_yoff + slab_rows) of the Ny global cell rows; x and
    // sigma operands carry slab_ghost ghost cell rows on either side (filled by the halo exchange)
    bool slab = false;
    int slab_yoff = 0, slab_rows = 0, slab_ghost = 0;
    // GeneralHelmholtz mode (helmholtz.h:74-80): symv(x, y) = chi x - helm_alpha (Elliptic x)
    bool helm = false;
    double helm_alpha = 0.;
    const double* helm_chi = nullptr;  // borrowed, nullptr = 1
    bool relaxed = false;  // dgb_elliptic2d_set_ordering: interior rows of the walker kernel in the relaxed operation order
    int kernel_mode = 0;  // dgb_elliptic2d_set_kernel: 0 auto, 1 tile kernel, 2 walker kernel, 3 unfused (reference launch sequence)
    void* walk_part[2] = {nullptr, nullptr};  // work partitions of the walker kernel (plain / fused-dot variant)
};
void elliptic2d_walker_release(Elliptic2dPlan& p);
bool elliptic2d_walker_supported(const Elliptic2dPlan& p, bool with_dot = false);

int elliptic2d_symv(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                    bool force_unfused);
int elliptic2d_fused_launch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st);
int elliptic2d_fused_launch_planes(Elliptic2dPlan& p, int nplanes, double alpha, const double* x, double beta, double* y, cudaStream_t st);

}  // namespace dgb
