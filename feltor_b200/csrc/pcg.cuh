// Internal: device-resident PCG scalars and the hooks fused kernels call when a dot product completes.
#pragma once
#include "superacc.cuh"
#include "comm.cuh"

namespace dgb {

struct PcgState {
    double nrmzr_old, pAp, alpha, beta, res, tol;
    int done;      // set by the device when ||r||_W < tol (pcg.h:177) or on a non-finite dot
    int iter;      // iteration at which `done` was raised
    int status;    // 1: a dot product met NaN/Inf (blas1.h:161)
    int cur;       // iteration the launches in flight belong to (written by the update kernel's finisher)
    int dist;      // 1: dots are completed by an integer allreduce + pcg_scalar_kernel, the fused finishers only publish
                   //    the local accumulator
};

struct FusedDot {
    const double* w;
    sa::DotSlot slot;
    PcgState* pcg;
    // multi-GPU with peer memory: the finishing block of the kernel exchanges the record itself (comm.cuh) and runs the hook
    P2pView p2p;
    unsigned long long epoch;
    int pdl = 0;  // launch with programmatic stream serialization (common.cuh); single-GPU solves only
};

// normalise + round a summed record in place (status = number of ranks that met NaN/Inf)
__device__ inline void finalize_record(dgb_dot_result* r) {
    long long acc[sa::BINS];
    for (int i = 0; i < sa::BINS; i++) acc[i] = r->acc[i];
    int neg = sa::normalize(acc, 1);
    for (int i = 0; i < sa::BINS; i++) r->acc[i] = acc[i];
    r->value = sa::round_normalized(acc, neg);
    r->status = r->status != 0 || r->pad != 0;
    r->pad = 0;
}

// pcg.h:166  alpha = nrmzr_old / dot(p, W, ap)
__device__ __forceinline__ void pcg_after_pAp(PcgState* st, const dgb_dot_result* r) {
    st->pAp = r->value;
    st->alpha = __ddiv_rn(st->nrmzr_old, r->value);
    if (r->status) { st->status = 1; st->done = 1; }
}

// pcg.h:171-177  res = sqrt(dot(r,W,r)); converged if res < tol
__device__ __forceinline__ void pcg_after_rr(PcgState* st, const dgb_dot_result* rr, int iter) {
    double res = __dsqrt_rn(rr->value);
    st->res = res;
    if (rr->status) { st->status = 1; st->done = 1; st->iter = iter; }
    else if (res < st->tol) { st->done = 1; st->iter = iter; }
}
// pcg.h:181-183  beta = nrmzr_new / nrmzr_old; nrmzr_old = nrmzr_new
__device__ __forceinline__ void pcg_after_zr(PcgState* st, const dgb_dot_result* zr, int iter) {
    double nw = zr->value;
    st->beta = __ddiv_rn(nw, st->nrmzr_old);
    st->nrmzr_old = nw;
    st->cur = iter;
    if (zr->status) { st->status = 1; st->done = 1; st->iter = iter; }
}

// what the finishing block of a fused apply + dot(p,W,Ap) kernel does once the local record is published.  `last` is true
// in ALL threads of that block (>= 64 threads).
__device__ inline void fused_dot_finish(bool last, PcgState* st, dgb_dot_result* results, const P2pView& pv, unsigned long long epoch) {
    if (!last) return;
    if (!st->dist) { if (threadIdx.x == 0) pcg_after_pAp(st, results); return; }
    if (!pv.enabled) return;  // NCCL path: the host enqueues the allreduce and the scalar kernel
    p2p_allreduce_records(pv, reinterpret_cast<long long*>(results), 0, 1, epoch);
    if (threadIdx.x == 0) { finalize_record(results); pcg_after_pAp(st, results); }
}

struct Elliptic2dPlan;
int elliptic2d_fused_launch_dot(Elliptic2dPlan& p, const double* x, double* y, cudaStream_t st, const FusedDot& fd);
// PCG iteration with the direction update folded into the operator kernel (elliptic_walker_fold.cu): possible when the plan
// runs on the walker kernel and every operand is TMA-describable (no periodic seam in x, 16-byte aligned rows)
bool elliptic2d_walker_fold_possible(const Elliptic2dPlan& p, const double* w);
int elliptic2d_walker_launch_fold(Elliptic2dPlan& p, const double* z, const double* p_old, double* p_new, double* ap, cudaStream_t st,
                                  const FusedDot& fd);

}  // namespace dgb
