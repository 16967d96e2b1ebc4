// Walker kernel with the PCG direction update folded into its loader (elliptic_walker.cuh, FOLD): own translation unit so
// that the two sets of instantiations compile in parallel.
#include "elliptic_walker.cuh"

namespace dgb {

// ap = A p_new with p_new = z + beta p_old (beta from the device-side PCG state), fused with dot(p_new, W, ap).
// z, p_old, p_new carry the ghost rows of a slab plan like the operand of elliptic2d_symv does.
int elliptic2d_walker_launch_fold(Elliptic2dPlan& p, const double* z, const double* p_old, double* p_new, double* ap, cudaStream_t st,
                                  const FusedDot& fd) {
    const bool plain = !p.helm && !p.vol;
#define DGB_WCASE(NN, DD)                                                                                                     \
    case NN * 10 + DD:                                                                                                        \
        return plain ? wlaunch<NN, DD, true, true, true>(p, 1., z, 0., ap, st, &fd, p_old, p_new)                              \
                     : wlaunch<NN, DD, true, false, true>(p, 1., z, 0., ap, st, &fd, p_old, p_new);
    switch (p.n * 10 + p.dirk) {
        DGB_WCASE(2, 0) DGB_WCASE(2, 1) DGB_WCASE(2, 2)
        DGB_WCASE(3, 0) DGB_WCASE(3, 1) DGB_WCASE(3, 2)
    }
#undef DGB_WCASE
    set_error("elliptic2d walker kernel: unsupported n=%d direction kind=%d", p.n, p.dirk);
    return DGB_ERR_UNSUPPORTED;
}

}  // namespace dgb
