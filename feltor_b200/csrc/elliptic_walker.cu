// Host side of the walker kernel (elliptic_walker.cuh): work partition, tensor maps, launch.
#include "elliptic_walker.cuh"

namespace dgb {

// ------------------------------------------------------------------------------------------------ host
// Cut the (warp column, cell row) space, column after column, into one contiguous piece of equal COST per warp.  A row of
// a column whose warp takes the general path (domain boundary in x, periodic seam) costs `wslow` interior rows.
// measured on B200 (n = 3, 1024^2): a boundary column costs ~2.3 interior ones, a column filled by LDGSTS ~3
static double walk_weight(bool manual, bool dot = false, bool centered = false) {
    static double wb = -1., wm = -1.;
    if (wb < 0.) {
        const char* e = getenv("DGB_WALK_SLOW_WEIGHT");
        wb = e ? atof(e) : 2.3;
        if (wb < 1.) wb = 1.;
        e = getenv("DGB_WALK_MANUAL_WEIGHT");
        wm = e ? atof(e) : 3.0;
        if (wm < 1.) wm = 1.;
    }
    // the fused-dot variants (8 warps), measured optima inside PCG with the folded direction update (profiles/fold_r02.md)
    if (!manual && dot && !getenv("DGB_WALK_SLOW_WEIGHT")) return centered ? 4.0 : 2.0;
    return manual ? wm : wb;
}
// host part (no CUDA): tasks[] = (warp column, first row, end row, 0), tbegin[g] .. tbegin[g+1] = the tasks of warp g
static void partition_host(int Nx, int Ny, int UL, int HL, int nwarps, int fx_lo, int fx_hi, bool wrapx, bool tma, int min_rows,
                           bool dot, std::vector<int4>& tasks, std::vector<int>& tbegin) {
    const int ncols = (Nx + UL - 1) / UL;
    std::vector<double> wcol(ncols);
    double total = 0.;
    for (int c = 0; c < ncols; c++) {
        const int cl = c * UL - HL;
        const bool fast = cl >= fx_lo && cl + 32 <= fx_hi;
        const bool manual = !tma || (wrapx && (cl < 0 || cl + 32 > Nx));
        wcol[c] = manual ? walk_weight(true) : (fast ? 1. : walk_weight(false, dot, HL == 2));
        total += wcol[c] * Ny;
    }
    // Every warp gets a budget B of cost units (one unit = one interior cell row); a piece costs rows * weight plus a fixed
    // overhead (GY-only step, leading rows, pipeline fill).  Pieces are laid along the columns in order; the smallest B
    // for which nwarps warps hold everything is found by bisection.  A piece is at least min_rows rows long (the ring
    // depth) unless the whole column is shorter.
    static double ovh = -1.;
    if (ovh < 0.) { const char* e = getenv("DGB_WALK_TASK_OVERHEAD"); ovh = e ? atof(e) : 2.0; if (ovh < 0.) ovh = 0.; }
    auto assign = [&](double B, bool keep) -> int {
        if (keep) { tasks.clear(); tbegin.assign(1, 0); }
        int g = 0, c = 0, row = 0;
        double budget = B;
        while (c < ncols) {
            const int left = Ny - row;
            int fit = (int)((budget - ovh) / wcol[c] + 1e-9);
            if (fit < min_rows && fit < left) {  // not worth starting a piece here: next warp
                if (budget >= B - 1e-9) fit = std::min(left, min_rows);  // a fresh warp always takes at least the minimum
                else { g++; budget = B; if (keep) tbegin.push_back((int)tasks.size()); continue; }
            }
            int take = std::min(fit, left);
            if (left - take > 0 && left - take < min_rows)  // do not leave a stub shorter than the ring depth behind
                take = (left - min_rows >= min_rows) ? left - min_rows : left;
            if (take < 1) take = std::min(left, min_rows);
            if (keep) tasks.push_back(make_int4(c, row, row + take, 0));
            budget -= take * wcol[c] + ovh;
            row += take;
            if (row >= Ny) { c++; row = 0; }
        }
        return g + 1;
    };
    double lo = total / nwarps, hi = 2. * total / nwarps + ovh * 4. + (double)min_rows * walk_weight(true);
    while (assign(hi, false) > nwarps) hi *= 1.5;
    for (int it = 0; it < 40; it++) {
        const double mid = 0.5 * (lo + hi);
        if (assign(mid, false) <= nwarps) hi = mid; else lo = mid;
    }
    assign(hi, true);
    while ((int)tbegin.size() < nwarps + 1) tbegin.push_back((int)tasks.size());
    tbegin[nwarps] = (int)tasks.size();
}

int build_partition(WalkPartition& P, int Nx, int Ny, int UL, int HL, int nwarps, int fx_lo, int fx_hi, bool wrapx,
                           bool tma, int min_rows, bool dot, cudaStream_t st) {
    const int key = (tma ? 1 : 0) | (wrapx ? 2 : 0) | (fx_lo << 2) | (fx_hi << 12);
    if (P.Nx == Nx && P.Ny == Ny && P.nwarps == nwarps && P.UL == UL && P.key == key && P.d_tasks) return 0;
    std::vector<int4> tasks;
    std::vector<int> tbegin;
    partition_host(Nx, Ny, UL, HL, nwarps, fx_lo, fx_hi, wrapx, tma, min_rows, dot, tasks, tbegin);
    cudaFree(P.d_tasks); cudaFree(P.d_tbegin);
    P.d_tasks = nullptr; P.d_tbegin = nullptr;
    DGB_CUDA(cudaMalloc(&P.d_tasks, (tasks.size() + 1) * sizeof(int4)));
    DGB_CUDA(cudaMalloc(&P.d_tbegin, tbegin.size() * sizeof(int)));
    // synchronous copies from pageable memory: the partition is built once per (grid, kernel) pair
    DGB_CUDA(cudaMemcpyAsync(P.d_tasks, tasks.data(), tasks.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    DGB_CUDA(cudaMemcpyAsync(P.d_tbegin, tbegin.data(), tbegin.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    DGB_CUDA(cudaStreamSynchronize(st));
    P.Nx = Nx; P.Ny = Ny; P.nwarps = nwarps; P.UL = UL; P.key = key; P.ntasks = (int)tasks.size();
    return 0;
}

}  // namespace dgb
// test hook (host only, no device needed): the work partition of the walker kernel for an Nx x Ny cell grid whose cells
// [fx_lo, fx_hi) are interior in x.  tasks_out: 3 ints per piece (warp column, first row, end row), tbegin_out: nwarps + 1.
extern "C" int dgb_debug_walker_partition(int Nx, int Ny, int centered, int nwarps, int fx_lo, int fx_hi, int wrapx, int tma, int dot,
                                          int* tasks_out, int max_tasks, int* ntasks, int* tbegin_out) {
    const int HL = centered ? 2 : 1, UL = 32 - 2 * HL, min_rows = (centered ? 2 : 1) + 2 + dgb::WALK_PD;
    std::vector<int4> tasks;
    std::vector<int> tbegin;
    dgb::partition_host(Nx, Ny, UL, HL, nwarps, fx_lo, fx_hi, wrapx != 0, tma != 0, min_rows, dot != 0, tasks, tbegin);
    *ntasks = (int)tasks.size();
    if ((int)tasks.size() > max_tasks) { dgb::set_error("dgb_debug_walker_partition: %zu pieces, room for %d", tasks.size(), max_tasks); return DGB_ERR_INVALID; }
    for (size_t k = 0; k < tasks.size(); k++) { tasks_out[3 * k] = tasks[k].x; tasks_out[3 * k + 1] = tasks[k].y; tasks_out[3 * k + 2] = tasks[k].z; }
    for (int g = 0; g <= nwarps; g++) tbegin_out[g] = tbegin[g];
    return 0;
}
namespace dgb {

void elliptic2d_walker_release(Elliptic2dPlan& p) {
    for (int k = 0; k < 2; k++) {
        WalkPartition* P = reinterpret_cast<WalkPartition*>(p.walk_part[k]);
        if (!P) continue;
        cudaFree(P->d_tasks);
        cudaFree(P->d_tbegin);
        delete P;
        p.walk_part[k] = nullptr;
    }
}

// structural support only; whether the walker is also the FASTER kernel is decided by elliptic2d_walker_supported
bool elliptic2d_walker_possible(const Elliptic2dPlan& p) {
    return p.fusable && (p.n == 2 || p.n == 3) && p.Nx >= 5 && (p.slab ? p.slab_rows : p.Ny) >= 5 && !(p.helm && p.helm_alpha == 0.);
}
bool elliptic2d_walker_supported(const Elliptic2dPlan& p, bool with_dot) {
    (void)with_dot;  // measured with the budget partition: the walker wins from ~400^2 cells on for every variant
    if (p.kernel_mode == DGB_ELLIPTIC_KERNEL_WALKER) return elliptic2d_walker_possible(p);
    if (p.kernel_mode != DGB_ELLIPTIC_KERNEL_AUTO) return false;
    // process-wide defaults for plans in auto mode (read once); a plan's own mode (dgb_elliptic2d_set_kernel) wins
    static int off = -1, force = -1;
    if (off < 0) { const char* e = getenv("DGB_ELLIPTIC_TILE"); off = (e && atoi(e)) ? 1 : 0; }
    if (force < 0) { const char* e = getenv("DGB_ELLIPTIC_WALKER"); force = (e && atoi(e)) ? 1 : 0; }
    if (off || !elliptic2d_walker_possible(p)) return false;
    if (force) return true;
    // measured on B200 (n = 3): the walker wins for the one-sided discretisations from ~512^2 cells on (87 vs 119 us at
    // 1024^2); below that each warp gets too few rows to amortise its pipeline fill.  The centered stencil (28 useful
    // lanes, 250 registers, 8 warps) gains less: 118 vs 156 us for the plain apply, +3..6 % in PCG with the fused dot
    const long long cells = (long long)p.Nx * (p.slab ? p.slab_rows : p.Ny);
    return cells >= 400 * 400;
}

bool elliptic2d_walker_fold_possible(const Elliptic2dPlan& p, const double* w) {
    const char* e = getenv("DGB_NO_TMA");
    if (e && atoi(e)) return false;
    if (!encode_fn()) return false;
    const size_t gh = p.slab ? (size_t)p.slab_ghost * p.n * p.Nx * p.n : 0;
    return elliptic2d_walker_supported(p, true) && !p.wrapx && ((size_t)p.Nx * p.n) % 2 == 0 && gh % 2 == 0 && aligned16(p.sigma) &&
           aligned16(w);
}

int elliptic2d_walker_launch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                             const FusedDot* fd) {
    const bool plain = !p.helm && !p.vol && (fd || beta == 0.);
#define DGB_WCASE(NN, DD)                                                                                          \
    case NN * 10 + DD:                                                                                             \
        if (fd) return plain ? wlaunch<NN, DD, true, true>(p, 1., x, 0., y, st, fd) : wlaunch<NN, DD, true, false>(p, 1., x, 0., y, st, fd); \
        return plain ? wlaunch<NN, DD, false, true>(p, alpha, x, beta, y, st, nullptr) : wlaunch<NN, DD, false, false>(p, alpha, x, beta, y, st, nullptr);
    switch (p.n * 10 + p.dirk) {
        DGB_WCASE(2, 0) DGB_WCASE(2, 1) DGB_WCASE(2, 2)
        DGB_WCASE(3, 0) DGB_WCASE(3, 1) DGB_WCASE(3, 2)
    }
#undef DGB_WCASE
    set_error("elliptic2d walker kernel: unsupported n=%d direction kind=%d", p.n, p.dirk);
    return DGB_ERR_UNSUPPORTED;
}

}  // namespace dgb
