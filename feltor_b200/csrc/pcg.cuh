// Internal: device-resident PCG scalars and the hooks fused kernels call when a dot product completes.
#pragma once
#include "superacc.cuh"

namespace dgb {

struct PcgState {
    double nrmzr_old, pAp, alpha, beta, res, tol;
    int done;      // set by the device when ||r||_W < tol (pcg.h:177) or on a non-finite dot
    int iter;      // iteration at which `done` was raised
    int status;    // 1: a dot product met NaN/Inf (blas1.h:161)
    int cur;       // iteration the launches in flight belong to (written by the update kernel's finisher)
};

struct FusedDot {
    const double* w;
    sa::DotSlot slot;
    PcgState* pcg;
};

// pcg.h:166  alpha = nrmzr_old / dot(p, W, ap)
__device__ __forceinline__ void pcg_after_pAp(PcgState* st, const dgb_dot_result* r) {
    st->pAp = r->value;
    st->alpha = __ddiv_rn(st->nrmzr_old, r->value);
    if (r->status) { st->status = 1; st->done = 1; }
}

struct Elliptic2dPlan;
int elliptic2d_fused_launch_dot(Elliptic2dPlan& p, const double* x, double* y, cudaStream_t st, const FusedDot& fd);

}  // namespace dgb
