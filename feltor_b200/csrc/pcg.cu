// dg::PCG<DVec>::solve (inc/dg/pcg.h:136-195) for A = Elliptic2d plan, vector preconditioner P, weights W.
// Same recurrences and the same exact (superaccumulator) dots as the reference, restructured so that one iteration
// is three launches and 128 B/dof of HBM traffic instead of ~46 vector passes:
//   K1  ap = A p  fused with dot(p,W,ap);   its last block computes alpha = nrmzr_old / pAp        (pcg.h:165-166)
//   K2  x += alpha p; r -= alpha ap; z = P r -> ap; dot(r,W,r) and dot(z,W,r) in the same pass;
//       its last blocks test ||r||_W < tol and compute beta = nrmzr_new / nrmzr_old               (pcg.h:167-181)
//   K3  p = z + beta p                                                                             (pcg.h:182)
// When the operator runs on the walker kernel with TMA operands, K3 disappears: from the second iteration on K1 reads z and
// the old direction, forms the new direction in its shared-memory ring (same Axpby arithmetic), applies the operator to it
// and writes it to the other direction buffer -- two launches and 120 B/dof per iteration (DGB_PCG_NO_FOLD=1 switches back).
// All scalars live in a PcgState record on the device.  The host enqueues batches of iterations and reads the
// record once per batch; after convergence the remaining launches of a batch exit at their first instruction, so
// x, r and the iteration count are exactly those of the reference's loop exit.
#include "elliptic.cuh"
#include "pcg.cuh"
#include "pcg_internal.cuh"
#include "comm.cuh"
#include <cmath>
#include <cstdlib>

namespace dgb {

constexpr int PCG_THREADS = 256;
constexpr int PCG_WARPS = PCG_THREADS / 32;

struct Pcg {
    size_t n = 0;
    double *r = nullptr, *p = nullptr, *ap = nullptr;
    double* p_base = nullptr;  // allocation behind p (p may be offset by the ghost rows of a slab)
    size_t p_cap = 0;
    // peer-memory halo (multi-GPU): the p buffers of all ranks mapped into this process
    void* p_peers[P2P_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    Comm* p_peers_comm = nullptr;
    bool p_mapped = false;
    // folded direction update (K3 inside K1's loader, elliptic_walker.cuh FOLD): the direction ping-pongs between p and p2,
    // z = P r gets its own buffer (K1 overwrites ap while neighbouring strips still read z); both laid out like p
    double *p2_base = nullptr, *p2 = nullptr, *z_base = nullptr, *z = nullptr;
    size_t fold_cap = 0;
    void* z_peers[P2P_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    Comm* z_peers_comm = nullptr;
    bool z_mapped = false;
    bool last_folded = false;
    PcgState* st = nullptr;       // device
    PcgState* st_host = nullptr;  // pinned
    sa::DotSlot slot;             // 4 slots: 0 pAp, 1 rWr, 2 zWr, 3 setup dots
    dgb_dot_result* results = nullptr;
    dgb_dot_result* results_host = nullptr;
    int check_every = 8;
    // optional per-kernel timing (bench.py roofline): CUDA events on the launching stream around K1/K2/K3
    bool profile = false;
    static constexpr int PROF_MAX = 64;
    cudaEvent_t ev[PROF_MAX][4];
    bool ev_ready = false;
    int prof_n = 0;
    double prof_ms[3] = {0., 0., 0.};
    long long prof_count = 0;
};

// -------------------------------------------------------------------------------------------- setup kernels
// generic 3-operand exact dot into slot `si`; finisher hook selected by MODE
//   MODE 0: none (host reads results[si])
//   MODE 1: nrmzr_old = value                       (pcg.h:160)
//   MODE 2: pAp/alpha                               (generic-operator path of K1)
template <int MODE>
__global__ void __launch_bounds__(PCG_THREADS)
pcg_dot3_kernel(size_t n, const double* __restrict__ x, const double* __restrict__ w, const double* __restrict__ y,
                sa::DotSlot slot, int si, PcgState* st) {
    __shared__ long long smem[sa::BINS];
    if (MODE == 2 && st->done) return;
    sa::block_init<1>(smem);
    long long* my = smem;
    sa::Fpe fpe;
    fpe.clear();
    int bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double pr = __dmul_rn(__dmul_rn(x[i], w[i]), y[i]);
        if (!isfinite(pr)) { bad = 1; pr = 0.; }
        fpe.add(pr, my);
    }
    fpe.flush_warp(my);
    if (sa::block_finish<1>(smem, bad, slot, si) && threadIdx.x == 0) {
        const dgb_dot_result* r = slot.result + si;
        if (st->dist) return;  // completed by the allreduce + pcg_scalar_kernel
        if (MODE == 1) { st->nrmzr_old = r->value; if (r->status) { st->status = 1; st->done = 1; } }
        if (MODE == 2) pcg_after_pAp(st, r);
    }
}

// K2 (pcg.h:167-181).  PREFETCH: the six 128-bit loads of the next trip are in flight while this trip is reduced (2 CTAs
// per SM); otherwise the kernel relies on occupancy (BPS CTAs per SM).
template <bool CHECK, bool PREFETCH, int BPS, bool P2P>  // P2P: the finishing block exchanges the dots over peer memory
__global__ void __launch_bounds__(PCG_THREADS, BPS)
pcg_update_kernel(size_t n, const double* __restrict__ p, const double* ap, double* __restrict__ x,
                  double* __restrict__ r, const double* __restrict__ P, const double* __restrict__ W, PcgState* st,
                  sa::DotSlot slot, int iter, P2pView peer, unsigned long long epoch, double* zout, double* rem_lo,
                  double* rem_up, size_t gcnt) {
    // zout: where z = P r goes -- the ap buffer itself (classic three-kernel iteration) or the z buffer of the folded
    // iteration, whose bottom / top rows are ALSO stored into the neighbours' ghost rows (rem_lo / rem_up, peer memory)
    __shared__ long long smem[2 * sa::BINS];  // one accumulator per dot and block: [0] rr (slot 1), [1] zr (slot 2)
    sa::block_init<2>(smem);  // programmatic dependent launch (common.cuh): nothing of the predecessor is touched before pdl_wait()
    pdl_wait();
    if (st->done) return;
    const double alpha = st->alpha, malpha = -alpha;
    long long* my_rr = smem;
    long long* my_zr = smem + sa::BINS;
    sa::Fpe frr, fzr, frr1, fzr1;  // two independent expansions per dot: the add cascades interleave in the FP64 pipe
    frr.clear();
    fzr.clear();
    frr1.clear();
    fzr1.clear();
    int bad = 0;
    const size_t T = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nvec = n / 2;
    size_t i = tid;
    double2 pv, av, xv, rv, Pv, Wv;
    if (PREFETCH && i < nvec) { pv = ld2(p + 2 * i); av = ld2(ap + 2 * i); xv = ld2(x + 2 * i); rv = ld2(r + 2 * i); Pv = ld2(P + 2 * i); Wv = ld2(W + 2 * i); }
    while (i < nvec) {
        const size_t inext = i + T;
        double2 pn, an, xn, rn, Pn, Wn;
        if (PREFETCH) {
            if (inext < nvec) { pn = ld2(p + 2 * inext); an = ld2(ap + 2 * inext); xn = ld2(x + 2 * inext); rn = ld2(r + 2 * inext); Pn = ld2(P + 2 * inext); Wn = ld2(W + 2 * inext); }
        } else {
            pv = ld2(p + 2 * i); av = ld2(ap + 2 * i); xv = ld2(x + 2 * i); rv = ld2(r + 2 * i); Pv = ld2(P + 2 * i); Wv = ld2(W + 2 * i);
        }
        // Axpby(alpha,1): y = y*1; y = fma(alpha, x, y)   (subroutines.h:260-274)
        xv.x = __fma_rn(alpha, pv.x, __dmul_rn(xv.x, 1.));
        xv.y = __fma_rn(alpha, pv.y, __dmul_rn(xv.y, 1.));
        rv.x = __fma_rn(malpha, av.x, __dmul_rn(rv.x, 1.));
        rv.y = __fma_rn(malpha, av.y, __dmul_rn(rv.y, 1.));
        double2 z = make_double2(__dmul_rn(Pv.x, rv.x), __dmul_rn(Pv.y, rv.y));  // symv(P, r, ap) == P*r
        st2(x + 2 * i, xv);
        st2(r + 2 * i, rv);
        st2(zout + 2 * i, z);
        if (rem_lo && 2 * i < gcnt) st2(rem_lo + 2 * i, z);
        if (rem_up && 2 * i >= n - gcnt) st2(rem_up + (2 * i - (n - gcnt)), z);
        double ra0 = 0., ra1 = 0.;
        if (CHECK) {
            double a0 = __dmul_rn(__dmul_rn(rv.x, Wv.x), rv.x), a1 = __dmul_rn(__dmul_rn(rv.y, Wv.y), rv.y);
            if (!isfinite(a0)) { bad = 1; a0 = 0.; }
            if (!isfinite(a1)) { bad = 1; a1 = 0.; }
            ra0 = frr.add_lazy(a0);
            ra1 = frr1.add_lazy(a1);
        }
        double b0 = __dmul_rn(__dmul_rn(z.x, Wv.x), rv.x), b1 = __dmul_rn(__dmul_rn(z.y, Wv.y), rv.y);
        if (!isfinite(b0)) { bad = 1; b0 = 0.; }
        if (!isfinite(b1)) { bad = 1; b1 = 0.; }
        const double rb0 = fzr.add_lazy(b0), rb1 = fzr1.add_lazy(b1);
        if (ra0 != 0. || ra1 != 0. || rb0 != 0. || rb1 != 0.) {  // rare: residues the expansions cannot hold
            sa::accumulate(my_rr, ra0, 1); sa::accumulate(my_rr, ra1, 1);
            sa::accumulate(my_zr, rb0, 1); sa::accumulate(my_zr, rb1, 1);
        }
        if (PREFETCH) { pv = pn; av = an; xv = xn; rv = rn; Pv = Pn; Wv = Wn; }
        i = inext;
    }
    if ((n & 1) && tid == 0) {
        size_t i = n - 1;
        double xs = __fma_rn(alpha, p[i], __dmul_rn(x[i], 1.));
        double rs = __fma_rn(malpha, ap[i], __dmul_rn(r[i], 1.));
        double z = __dmul_rn(P[i], rs);
        x[i] = xs; r[i] = rs; zout[i] = z;
        if (CHECK) {
            double a0 = __dmul_rn(__dmul_rn(rs, W[i]), rs);
            if (!isfinite(a0)) { bad = 1; a0 = 0.; }
            frr.add(a0, my_rr);
        }
        double b0 = __dmul_rn(__dmul_rn(z, W[i]), rs);
        if (!isfinite(b0)) { bad = 1; b0 = 0.; }
        fzr.add(b0, my_zr);
    }
    pdl_trigger();  // the streaming part is done: the next kernel's launch and prologue may overlap the exact-dot tail
    // both dots leave together: one fence / ticket sequence
    fzr.merge(fzr1, my_zr);
    if (CHECK) {
        frr.merge(frr1, my_rr);
        sa::flush_warp2(frr, my_rr, fzr, my_zr);
    } else {
        fzr.flush_warp(my_zr);
    }
    const bool last = CHECK ? sa::block_finish_multi<2>(smem, bad, slot, 1) : sa::block_finish_multi<1>(my_zr, bad, slot, 2);
    if (!last) return;
    if (st->dist) {
        if (!P2P || !peer.enabled) return;  // NCCL path: allreduce + pcg_scalar_kernel follow on the stream
        // peer-memory path: this (the finishing) block exchanges the records with the other ranks right here
        p2p_allreduce_records(peer, reinterpret_cast<long long*>(slot.result), CHECK ? 1 : 2, CHECK ? 2 : 1, epoch);
        if (threadIdx.x == 0) {
            if (CHECK) finalize_record(slot.result + 1);
            finalize_record(slot.result + 2);
        }
    }
    if (threadIdx.x == 0) {
        if (CHECK) pcg_after_rr(st, slot.result + 1, iter);
        pcg_after_zr(st, slot.result + 2, iter);
    }
}

// multi-GPU: after the integer allreduce of the local accumulators, one thread normalises, rounds and runs the hook
//   MODE 0: res[3] only   MODE 1: nrmzr_old = dot (pcg.h:160)   MODE 2: alpha (pcg.h:166)   MODE 3: K2 hooks
template <int MODE>
__global__ void __launch_bounds__(64) pcg_scalar_kernel(PcgState* st, dgb_dot_result* res, int iter, int check, P2pView pv, int first,
                                                        int count, unsigned long long epoch) {
    // peer-memory path: the exchange of the local records happens here (comm.cuh); NCCL path: already summed
    if (pv.enabled) p2p_allreduce_records(pv, reinterpret_cast<long long*>(res), first, count, epoch);
    if (threadIdx.x != 0) return;
    if (MODE >= 2 && st->done) return;
    if (MODE == 0) finalize_record(res + 3);
    if (MODE == 1) { finalize_record(res + 0); st->nrmzr_old = res[0].value; if (res[0].status) { st->status = 1; st->done = 1; } }
    if (MODE == 2) { finalize_record(res + 0); pcg_after_pAp(st, res + 0); }
    if (MODE == 3) {
        if (check) { finalize_record(res + 1); pcg_after_rr(st, res + 1, iter); }
        finalize_record(res + 2);
        pcg_after_zr(st, res + 2, iter);
    }
}

// K3 (pcg.h:182): axpby(1, ap, beta, p): p = p*beta; p = fma(1, ap, p).  Pure stream (24 B/dof): full occupancy, four
// front-batched 128-bit loads per operand and thread, 8 CTAs per SM (the layout of the blas1 kernels).
// Multi-GPU: the bottom / top `gcnt` doubles (the rows the neighbours need) are ALSO stored straight into the upper ghost
// rows of the lower neighbour (`rem_lo`) / the lower ghost rows of the upper neighbour (`rem_up`) through peer memory --
// the halo exchange of the next operator application rides in this kernel, a neighbour barrier follows.
__global__ void __launch_bounds__(PCG_THREADS, 4)
pcg_direction_kernel(size_t n, const double* __restrict__ z, double* __restrict__ p, const PcgState* st, double* rem_lo,
                     double* rem_up, size_t gcnt) {
    pdl_wait();     // programmatic dependent launch (common.cuh); no prologue worth overlapping, but the lighter kernel boundary counts
    pdl_trigger();  // the successor (K1) still waits for this grid to complete before it reads p
    if (st->done) return;
    constexpr int U = 4;
    const double beta = st->beta;
    const size_t T = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nvec = n / 2;
    for (size_t base = tid; base < nvec; base += (size_t)U * T) {
        double2 zv[U], pv[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = base + (size_t)u * T;
            if (i < nvec) { zv[u] = ld2(z + 2 * i); pv[u] = ld2(p + 2 * i); }
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = base + (size_t)u * T;
            if (i < nvec) {
                pv[u].x = __fma_rn(1., zv[u].x, __dmul_rn(pv[u].x, beta));
                pv[u].y = __fma_rn(1., zv[u].y, __dmul_rn(pv[u].y, beta));
                st2(p + 2 * i, pv[u]);
                if (rem_lo && 2 * i < gcnt) st2(rem_lo + 2 * i, pv[u]);
                if (rem_up && 2 * i >= n - gcnt) st2(rem_up + (2 * i - (n - gcnt)), pv[u]);
            }
        }
    }
    if ((n & 1) && tid == 0) p[n - 1] = __fma_rn(1., z[n - 1], __dmul_rn(p[n - 1], beta));
}

// r = b - r (axpby(1, b, -1, r), pcg.h:156) ; p = P*r (pcg.h:159)
__global__ void __launch_bounds__(PCG_THREADS)
pcg_residual_kernel(size_t n, const double* __restrict__ b, double* __restrict__ r) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        r[i] = __fma_rn(1., b[i], __dmul_rn(r[i], -1.));
}
__global__ void __launch_bounds__(PCG_THREADS)
pcg_precond_kernel(size_t n, const double* __restrict__ P, const double* __restrict__ r, double* __restrict__ p) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = __dmul_rn(P[i], r[i]);
}

static unsigned grid_for(size_t n, int per_thread) {
    size_t want = (n + (size_t)PCG_THREADS * per_thread - 1) / ((size_t)PCG_THREADS * per_thread);
    if (want == 0) want = 1;
    size_t cap = (size_t)sm_count() * 2;  // the reduction kernels hold ~100 registers: two resident CTAs per SM, one wave
    return (unsigned)(want < cap ? want : cap);
}

static int fetch_state(Pcg& s, cudaStream_t st) {
    DGB_CUDA(cudaMemcpyAsync(s.st_host, s.st, sizeof(PcgState), cudaMemcpyDeviceToHost, st));
    DGB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
static int fetch_result(Pcg& s, int si, cudaStream_t st) {
    DGB_CUDA(cudaMemcpyAsync(s.results_host + si, s.results + si, sizeof(dgb_dot_result), cudaMemcpyDeviceToHost, st));
    DGB_CUDA(cudaStreamSynchronize(st));
    if (s.results_host[si].status) {
        set_error("dot product failed since one of the inputs contains NaN or Inf");
        return DGB_ERR_NOTFINITE;
    }
    return 0;
}

// dist: complete the dot(s) in slots [first, first+count) across ranks, then run the scalar hook MODE
template <int MODE>
static int dist_finish(Pcg& s, Comm* comm, int first, int count, int iter, int check, cudaStream_t st) {
    const P2pView pv = comm_p2p_view(comm);
    unsigned long long epoch = 0;
    if (pv.enabled) epoch = comm_p2p_next_epoch(comm, first, count);
    else {
        int e = comm_allreduce_i64(comm, reinterpret_cast<long long*>(s.results + first), (size_t)count * sizeof(dgb_dot_result) / 8, st);
        if (e) return e;
    }
    pcg_scalar_kernel<MODE><<<1, 64, 0, st>>>(s.st, s.results, iter, check, pv, first, count, epoch);
    DGB_LAUNCHED();
    return 0;
}

int pcg_solve_impl(Pcg& s, Comm* comm, Elliptic2dPlan& A, double* x, const double* b, const double* P, const double* W,
                   double eps, double nrmb_correction, int test_frequency, int max_iter, int* iterations, cudaStream_t st) {
    const size_t n = s.n;
    const bool dist = comm != nullptr;
    const size_t nA = A.slab ? (size_t)A.slab_rows * A.n * A.Nx * A.n : A.size;
    if (nA != n) { set_error("dgb_pcg_solve: operator size %zu != workspace size %zu", nA, n); return DGB_ERR_INVALID; }
    if (dist != A.slab) { set_error("dgb_pcg_solve: a slab plan needs the distributed solve and vice versa"); return DGB_ERR_INVALID; }
    if (test_frequency < 1) { set_error("dgb_pcg_solve: test_frequency must be >= 1"); return DGB_ERR_INVALID; }
    // K2 / K3 move every operand with 128-bit accesses.  Device vectors of the reference (thrust::device_vector) and of
    // cudaMalloc are 256-byte aligned; an offset view is rejected here instead of faulting inside the kernel
    if (!aligned16(x) || !aligned16(b) || !aligned16(P) || !aligned16(W)) {
        set_error("dgb_pcg_solve: x, b, P and W must be 16-byte aligned");
        return DGB_ERR_INVALID;
    }
    int e;
    // search direction: in slab mode it carries ghost rows that the halo exchange fills.  The ghost block in front is
    // padded to an even number of doubles so that s.p itself stays 16-byte aligned for odd n * Nx * ghost
    const size_t row_len = dist ? (size_t)A.Nx * A.n : 0;
    const size_t ghost_rows = dist ? (size_t)A.slab_ghost * A.n : 0, gh_rows = ghost_rows * row_len, gh = gh_rows + (gh_rows & 1);
    if (s.p_cap < n + 2 * gh) {
        if (s.p_mapped) { comm_p2p_unmap(s.p_peers_comm, s.p_peers); s.p_mapped = false; }
        cudaFree(s.p_base);
        s.p_base = nullptr;
        DGB_CUDA(cudaMalloc(&s.p_base, (n + 2 * gh) * sizeof(double)));
        DGB_CUDA(cudaMemsetAsync(s.p_base, 0, (n + 2 * gh) * sizeof(double), st));
        s.p_cap = n + 2 * gh;
    }
    s.p = s.p_base + gh;
    // peer-memory halo: map the p buffers of all ranks once per allocation (collective)
    double *rem_lo = nullptr, *rem_up = nullptr;
    int nb_lower = -1, nb_upper = -1;
    bool p2p_halo = false;
    if (dist && comm_size(comm) > 1) {
        if (!s.p_mapped || s.p_peers_comm != comm) {
            if (s.p_mapped) comm_p2p_unmap(s.p_peers_comm, s.p_peers);
            DGB_CUDA(cudaStreamSynchronize(st));
            int mapped = 0;
            if ((e = comm_p2p_map(comm, s.p_base, s.p_peers, &mapped))) return e;
            s.p_mapped = mapped != 0;
            s.p_peers_comm = comm;
        }
        const int rank = comm_rank(comm), size = comm_size(comm);
        nb_lower = rank - 1; nb_upper = rank + 1;
        if (nb_lower < 0) nb_lower = A.wrapy ? size - 1 : -1;
        if (nb_upper >= size) nb_upper = A.wrapy ? 0 : -1;
        // every rank must take the same decision: equal slab heights are not required, but the neighbour's layout is
        // [gh | n_nb | gh] with ITS n -- its upper ghost starts at gh + n_nb, which we do not know; so only the LOWER ghost
        // of the upper neighbour (offset 0) and, for equal n, the upper ghost of the lower neighbour are addressable.
        // bench/solver slabs are equal-sized whenever Ny divides by the rank count; otherwise NCCL does the exchange.
        p2p_halo = s.p_mapped && gh > 0 && (n % 2 == 0) && (gh_rows % 2 == 0) && A.Ny % size == 0 && A.slab_rows == A.Ny / size;
        if (p2p_halo) {
            if (nb_lower >= 0) rem_lo = reinterpret_cast<double*>(s.p_peers[nb_lower]) + gh + n;  // its upper ghost rows
            if (nb_upper >= 0) rem_up = reinterpret_cast<double*>(s.p_peers[nb_upper]);           // its lower ghost rows
        }
    }
    auto halo = [&](double* v) -> int {
        if (!dist) return 0;
        return comm_halo_rows(comm, v, row_len, (size_t)A.slab_rows * A.n, ghost_rows, A.wrapy, st);
    };
    PcgState init{};
    init.dist = dist ? 1 : 0;
    DGB_CUDA(cudaMemcpyAsync(s.st, &init, sizeof(PcgState), cudaMemcpyHostToDevice, st));
    const unsigned g1 = grid_for(n, 4);
    // pcg.h:140  nrmb = sqrt(dot(b, W, b))
    pcg_dot3_kernel<0><<<g1, PCG_THREADS, 0, st>>>(n, b, W, b, s.slot, 3, s.st);
    DGB_LAUNCHED();
    if (dist && (e = dist_finish<0>(s, comm, 3, 1, 0, 0, st))) return e;
    if ((e = fetch_result(s, 3, st))) return e;
    const double nrmb = std::sqrt(s.results_host[3].value);
    const double tol = eps * (nrmb + nrmb_correction);
    if (nrmb == 0) {  // pcg.h:150-154
        DGB_CUDA(cudaMemsetAsync(x, 0, n * sizeof(double), st));
        *iterations = 0;
        return 0;
    }
    init.tol = tol;
    DGB_CUDA(cudaMemcpyAsync(s.st, &init, sizeof(PcgState), cudaMemcpyHostToDevice, st));
    if (dist) {  // the operator needs the ghost rows of its argument: stage x in the padded buffer
        DGB_CUDA(cudaMemcpyAsync(s.p, x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if ((e = halo(s.p))) return e;
        if ((e = elliptic2d_symv(A, 1., s.p, 0., s.r, st, false))) return e;
    } else {
        if ((e = elliptic2d_symv(A, 1., x, 0., s.r, st, false))) return e;  // pcg.h:155
    }
    pcg_residual_kernel<<<grid_for(n, 1), PCG_THREADS, 0, st>>>(n, b, s.r);  // pcg.h:156
    DGB_LAUNCHED();
    pcg_dot3_kernel<0><<<g1, PCG_THREADS, 0, st>>>(n, s.r, W, s.r, s.slot, 3, s.st);
    DGB_LAUNCHED();
    if (dist && (e = dist_finish<0>(s, comm, 3, 1, 0, 0, st))) return e;
    pcg_precond_kernel<<<grid_for(n, 1), PCG_THREADS, 0, st>>>(n, P, s.r, s.p);  // pcg.h:159
    DGB_LAUNCHED();
    pcg_dot3_kernel<1><<<g1, PCG_THREADS, 0, st>>>(n, s.p, W, s.r, s.slot, 0, s.st);  // pcg.h:160
    DGB_LAUNCHED();
    if (dist && (e = dist_finish<1>(s, comm, 0, 1, 0, 0, st))) return e;
    if ((e = halo(s.p))) return e;
    if ((e = fetch_result(s, 3, st))) return e;
    if (std::sqrt(s.results_host[3].value) < tol) { *iterations = 0; return 0; }  // pcg.h:157
    const bool identity_chi = !A.chi[0] && !A.chi[1] && !A.chi[2] && !A.chi[3];
    const bool fused = A.fusable && identity_chi && !A.chi_weight_jump && (dist || !getenv("DGB_ELLIPTIC_UNFUSED")) &&
                       (dist || A.kernel_mode != DGB_ELLIPTIC_KERNEL_UNFUSED) &&
                       (!A.helm || A.helm_alpha != 0.);
    // ---- folded direction update: buffers (and, multi-GPU, the peer mapping of z)
    bool fold = fused && !getenv("DGB_PCG_NO_FOLD") && elliptic2d_walker_fold_possible(A, W) && n % 2 == 0 &&
                true;
    double *zrem_lo = nullptr, *zrem_up = nullptr;
    if (fold) {
        if (s.fold_cap < n + 2 * gh) {
            if (s.z_mapped) { comm_p2p_unmap(s.z_peers_comm, s.z_peers); s.z_mapped = false; }
            cudaFree(s.p2_base); cudaFree(s.z_base);
            s.p2_base = s.z_base = nullptr;
            s.fold_cap = 0;
            DGB_CUDA(cudaMalloc(&s.p2_base, (n + 2 * gh) * sizeof(double)));
            DGB_CUDA(cudaMalloc(&s.z_base, (n + 2 * gh) * sizeof(double)));
            DGB_CUDA(cudaMemsetAsync(s.p2_base, 0, (n + 2 * gh) * sizeof(double), st));
            DGB_CUDA(cudaMemsetAsync(s.z_base, 0, (n + 2 * gh) * sizeof(double), st));
            s.fold_cap = n + 2 * gh;
        }
        s.p2 = s.p2_base + gh;
        s.z = s.z_base + gh;
        if (p2p_halo) {  // the same collective decision on every rank: p2p_halo is derived from agreed quantities
            if (!s.z_mapped || s.z_peers_comm != comm) {
                if (s.z_mapped) comm_p2p_unmap(s.z_peers_comm, s.z_peers);
                DGB_CUDA(cudaStreamSynchronize(st));
                int mapped = 0;
                if ((e = comm_p2p_map(comm, s.z_base, s.z_peers, &mapped))) return e;
                s.z_mapped = mapped != 0;
                s.z_peers_comm = comm;
            }
            if (s.z_mapped) {
                if (nb_lower >= 0) zrem_lo = reinterpret_cast<double*>(s.z_peers[nb_lower]) + gh + n;
                if (nb_upper >= 0) zrem_up = reinterpret_cast<double*>(s.z_peers[nb_upper]);
            }
        }
    }
    s.last_folded = fold;
    const bool z_by_peer = fold && p2p_halo && s.z_mapped;
    const P2pView pview = comm_p2p_view(comm);
    // The finishing block of K1/K2 can run the peer-memory exchange itself (one launch less per dot); measured at 2 GPUs this
    // is 3 % SLOWER than the separate 64-thread exchange kernel (the serial tail of a 300-CTA kernel gets longer), so it is
    // opt-in: DGB_P2P_IN_KERNEL=1
    static int in_kernel = -1;
    if (in_kernel < 0) { const char* ev = getenv("DGB_P2P_IN_KERNEL"); in_kernel = (ev && atoi(ev)) ? 1 : 0; }
    const bool p2p_dots = dist && pview.enabled && in_kernel;
    FusedDot fd{W, s.slot, s.st, pview, 0ull};
    if (!p2p_dots) fd.p2p.enabled = 0;
    // programmatic dependent launches between the two kernels of the iteration (single GPU, walker kernel): opt-out DGB_PDL=0
    static int pdl_env = -1;
    if (pdl_env < 0) { const char* ev = getenv("DGB_PDL"); pdl_env = ev ? atoi(ev) : 3; }  // bit 0: K2 launches, bit 1: K1 launches
    const bool pdl_ok = !dist && !s.profile;
    fd.pdl = (pdl_ok && (pdl_env & 2)) ? 1 : 0;
    // K2 as a dependent launch pays off on the latency-bound stages (+4-5 % at <= 256^2 cells) and costs 5-9 % behind the walker
    // kernel on large grids (measured, profiles/pdl_r02.md; cause not established): size gate, DGB_PDL_K2_MAX overrides
    static long long k2_max = -1;
    if (k2_max < 0) { const char* ev = getenv("DGB_PDL_K2_MAX"); k2_max = ev ? atoll(ev) : 1200000ll; }
    const bool pdl_k2 = pdl_ok && (pdl_env & 1) && (long long)n <= k2_max;
    static int k2_elems = -1;  // experiment knob: elements per thread that size K2's grid
    if (k2_elems < 0) { const char* ev = getenv("DGB_PCG_K2_ELEMS"); k2_elems = ev ? atoi(ev) : 2; if (k2_elems < 2) k2_elems = 2; }
    const unsigned g2 = grid_for(n, k2_elems);
    static int k2_variant = -1;  // experiment knob: 0 register prefetch, 2 CTAs/SM   1 no prefetch, 4 CTAs/SM   2 no prefetch, 3 CTAs/SM
    if (k2_variant < 0) { const char* ev = getenv("DGB_PCG_K2_VARIANT"); k2_variant = ev ? atoi(ev) : 0; }
    unsigned g3;
    {
        size_t want = (n / 2 + (size_t)PCG_THREADS * 4 - 1) / ((size_t)PCG_THREADS * 4), cap = (size_t)sm_count() * 8;
        if (want == 0) want = 1;
        g3 = (unsigned)(want < cap ? want : cap);
    }
    double *pcur = s.p, *palt = s.p2;  // folded iteration: direction of this / the next iteration
    int i = 1;
    while (i < max_iter) {
        // a batch ends where the host looks at the device state.  Convergence can only be raised by an iteration that
        // tests the residual (i % test_frequency == 0), so with test_frequency > 1 the batch ends right after such an
        // iteration instead of at an arbitrary one (coarse multigrid stages test every 10th iteration, multigrid.h:646)
        int stop = i + s.check_every;
        if (test_frequency > 1) {
            const int next_test = (i / test_frequency + 1) * test_frequency;  // first tested iteration >= i (i itself if divisible: handled by +1 below)
            stop = (i % test_frequency == 0 ? i : next_test) + 1;
            while (stop - i < s.check_every / 2) stop += test_frequency;       // keep batches from getting tiny
        }
        if (stop > max_iter) stop = max_iter;
        for (; i < stop; i++) {
            const bool prof = s.profile && s.prof_n < Pcg::PROF_MAX;
            const int check = i % test_frequency == 0;
            if (prof) cudaEventRecord(s.ev[s.prof_n][0], st);
            if (fused) {
                if (p2p_dots) fd.epoch = comm_p2p_next_epoch(comm, 0, 1);
                if (fold && i > 1) {
                    if ((e = elliptic2d_walker_launch_fold(A, s.z, pcur, palt, s.ap, st, fd))) return e;
                    std::swap(pcur, palt);
                } else if ((e = elliptic2d_fused_launch_dot(A, pcur, s.ap, st, fd))) return e;
            } else {
                if ((e = elliptic2d_symv(A, 1., s.p, 0., s.ap, st, false))) return e;
                pcg_dot3_kernel<2><<<g1, PCG_THREADS, 0, st>>>(n, s.p, W, s.ap, s.slot, 0, s.st);
                DGB_LAUNCHED();
            }
            if (dist && !(fused && p2p_dots) && (e = dist_finish<2>(s, comm, 0, 1, i, 0, st))) return e;
            if (prof) cudaEventRecord(s.ev[s.prof_n][1], st);
#define DGB_K2(C_, P_, B_) (p2p_dots ? pcg_update_kernel<C_, P_, B_, true> : pcg_update_kernel<C_, P_, B_, false>)
            P2pView k2v = pview;
            k2v.enabled = p2p_dots ? 1 : 0;
            const unsigned long long k2e = p2p_dots ? comm_p2p_next_epoch(comm, check ? 1 : 2, check ? 2 : 1) : 0ull;
            if (k2_variant == 0 && pdl_k2) {
                double* zo = fold ? s.z : s.ap;
                double *zl = z_by_peer ? zrem_lo : nullptr, *zu = z_by_peer ? zrem_up : nullptr;
                if (check) DGB_CUDA(launch_pdl(pcg_update_kernel<true, true, 2, false>, dim3(g2), dim3(PCG_THREADS), 0, st, n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, zo, zl, zu, gh));
                else DGB_CUDA(launch_pdl(pcg_update_kernel<false, true, 2, false>, dim3(g2), dim3(PCG_THREADS), 0, st, n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, zo, zl, zu, gh));
            } else if (k2_variant == 0) {
                if (check) DGB_K2(true, true, 2)<<<g2, PCG_THREADS, 0, st>>>(n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, fold ? s.z : s.ap, z_by_peer ? zrem_lo : nullptr, z_by_peer ? zrem_up : nullptr, gh);
                else DGB_K2(false, true, 2)<<<g2, PCG_THREADS, 0, st>>>(n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, fold ? s.z : s.ap, z_by_peer ? zrem_lo : nullptr, z_by_peer ? zrem_up : nullptr, gh);
            } else if (k2_variant == 1) {
                if (check) DGB_K2(true, false, 4)<<<2 * g2, PCG_THREADS, 0, st>>>(n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, fold ? s.z : s.ap, z_by_peer ? zrem_lo : nullptr, z_by_peer ? zrem_up : nullptr, gh);
                else DGB_K2(false, false, 4)<<<2 * g2, PCG_THREADS, 0, st>>>(n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, fold ? s.z : s.ap, z_by_peer ? zrem_lo : nullptr, z_by_peer ? zrem_up : nullptr, gh);
            } else {
                if (check) DGB_K2(true, false, 3)<<<g2 / 2 * 3, PCG_THREADS, 0, st>>>(n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, fold ? s.z : s.ap, z_by_peer ? zrem_lo : nullptr, z_by_peer ? zrem_up : nullptr, gh);
                else DGB_K2(false, false, 3)<<<g2 / 2 * 3, PCG_THREADS, 0, st>>>(n, pcur, s.ap, x, s.r, P, W, s.st, s.slot, i, k2v, k2e, fold ? s.z : s.ap, z_by_peer ? zrem_lo : nullptr, z_by_peer ? zrem_up : nullptr, gh);
            }
#undef DGB_K2
            DGB_LAUNCHED();
            if (dist && !p2p_dots && (e = dist_finish<3>(s, comm, check ? 1 : 2, check ? 2 : 1, i, check, st))) return e;
            if (prof) cudaEventRecord(s.ev[s.prof_n][2], st);
            if (fold) {  // the next K1 forms the direction itself; it needs the neighbours' boundary rows of z
                // peer-memory dots: the all-to-all record exchange of pcg_scalar_kernel<3> (enqueued after K2 on every rank,
                // returns only when every rank's record has arrived) already orders the neighbours' K2 -- and with it their
                // stores into our ghost rows -- before our next K1: no separate neighbour barrier
                if (z_by_peer && pview.enabled && !p2p_dots) {}
                else if (z_by_peer) { if ((e = comm_p2p_neighbour_barrier(comm, nb_lower, nb_upper, st))) return e; }
                else if ((e = halo(s.z))) return e;
            } else {
                if (pdl_k2) DGB_CUDA(launch_pdl(pcg_direction_kernel, dim3(g3), dim3(PCG_THREADS), 0, st, n, (const double*)s.ap, s.p, s.st, rem_lo, rem_up, gh));
                else pcg_direction_kernel<<<g3, PCG_THREADS, 0, st>>>(n, s.ap, s.p, s.st, rem_lo, rem_up, gh);
                DGB_LAUNCHED();
                if (p2p_halo) { if ((e = comm_p2p_neighbour_barrier(comm, nb_lower, nb_upper, st))) return e; }
                else if ((e = halo(s.p))) return e;
            }
            if (prof) cudaEventRecord(s.ev[s.prof_n++][3], st);
        }
        if ((e = fetch_state(s, st))) return e;
        if (s.profile) {  // the stream is idle here: harvest the event pairs
            for (int k = 0; k < s.prof_n; k++)
                for (int j = 0; j < 3; j++) {
                    float ms = 0.f;
                    if (cudaEventElapsedTime(&ms, s.ev[k][j], s.ev[k][j + 1]) == cudaSuccess) s.prof_ms[j] += ms;
                }
            s.prof_count += s.prof_n;
            s.prof_n = 0;
        }
        if (s.st_host->status) {
            set_error("dot product failed since one of the inputs contains NaN or Inf");
            return DGB_ERR_NOTFINITE;
        }
        if (s.st_host->done) { *iterations = s.st_host->iter; return 0; }
    }
    *iterations = max_iter;
    set_error("PCG failed to converge within max_iter = %d iterations (residual %g, tolerance %g)", max_iter,
              s.st_host->res, tol);
    return DGB_ERR_NOCONVERGE;
}

int pcg_solve(Pcg& s, Elliptic2dPlan& A, double* x, const double* b, const double* P, const double* W, double eps,
              double nrmb_correction, int test_frequency, int max_iter, int* iterations, cudaStream_t st) {
    return pcg_solve_impl(s, nullptr, A, x, b, P, W, eps, nrmb_correction, test_frequency, max_iter, iterations, st);
}

static int pcg_alloc(Pcg* s, size_t n) {
    s->n = n;
    size_t bytes = (n ? n : 1) * sizeof(double);
    DGB_CUDA(cudaMalloc(&s->r, bytes));
    DGB_CUDA(cudaMalloc(&s->p_base, bytes));
    s->p = s->p_base;
    s->p_cap = n ? n : 1;
    DGB_CUDA(cudaMalloc(&s->ap, bytes));
    DGB_CUDA(cudaMalloc(&s->st, sizeof(PcgState)));
    DGB_CUDA(cudaMemset(s->st, 0, sizeof(PcgState)));
    DGB_CUDA(cudaMallocHost(&s->st_host, sizeof(PcgState)));
    const int ns = 4;
    DGB_CUDA(cudaMalloc(&s->slot.gacc, ns * sa::GACC_WORDS * sizeof(long long)));
    DGB_CUDA(cudaMemset(s->slot.gacc, 0, ns * sa::GACC_WORDS * sizeof(long long)));
    DGB_CUDA(cudaMalloc(&s->slot.gstatus, ns * sizeof(int)));
    DGB_CUDA(cudaMemset(s->slot.gstatus, 0, ns * sizeof(int)));
    DGB_CUDA(cudaMalloc(&s->slot.ticket, ns * sizeof(unsigned int)));
    DGB_CUDA(cudaMemset(s->slot.ticket, 0, ns * sizeof(unsigned int)));
    DGB_CUDA(cudaMalloc(&s->results, ns * sizeof(dgb_dot_result)));
    DGB_CUDA(cudaMemset(s->results, 0, ns * sizeof(dgb_dot_result)));
    DGB_CUDA(cudaMallocHost(&s->results_host, ns * sizeof(dgb_dot_result)));
    s->slot.result = s->results;
    const char* ce = getenv("DGB_PCG_CHECK_EVERY");
    if (ce && atoi(ce) > 0) s->check_every = atoi(ce);
    return 0;
}
Pcg* pcg_new(size_t n, int* err) {
    Pcg* s = new Pcg();
    int e = pcg_alloc(s, n);
    if (e) { if (err) *err = e; pcg_delete(s); return nullptr; }
    return s;
}
void pcg_delete(Pcg* s) {
    if (!s) return;
    if (s->ev_ready)
        for (int k = 0; k < Pcg::PROF_MAX; k++)
            for (int j = 0; j < 4; j++) cudaEventDestroy(s->ev[k][j]);
    if (s->p_mapped) comm_p2p_unmap(s->p_peers_comm, s->p_peers);
    if (s->z_mapped) comm_p2p_unmap(s->z_peers_comm, s->z_peers);
    cudaFree(s->p2_base); cudaFree(s->z_base);
    cudaFree(s->r); cudaFree(s->p_base); cudaFree(s->ap); cudaFree(s->st); cudaFreeHost(s->st_host);
    cudaFree(s->slot.gacc); cudaFree(s->slot.gstatus); cudaFree(s->slot.ticket);
    cudaFree(s->results); cudaFreeHost(s->results_host);
    delete s;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_pcg_create(dgb_pcg** out, size_t n) {
    int e = 0;
    Pcg* s = pcg_new(n, &e);
    if (!s) return e;
    *out = reinterpret_cast<dgb_pcg*>(s);
    return 0;
}
int dgb_pcg_set_profile(dgb_pcg* h, int on) {
    Pcg* s = reinterpret_cast<Pcg*>(h);
    if (on && !s->ev_ready) {
        for (int k = 0; k < Pcg::PROF_MAX; k++)
            for (int j = 0; j < 4; j++) DGB_CUDA(cudaEventCreate(&s->ev[k][j]));
        s->ev_ready = true;
    }
    s->profile = on != 0;
    s->prof_n = 0;
    s->prof_count = 0;
    s->prof_ms[0] = s->prof_ms[1] = s->prof_ms[2] = 0.;
    return 0;
}
int dgb_pcg_get_profile(dgb_pcg* h, double* ms_apply_dot, double* ms_update, double* ms_direction, long long* iterations) {
    Pcg* s = reinterpret_cast<Pcg*>(h);
    if (ms_apply_dot) *ms_apply_dot = s->prof_ms[0];
    if (ms_update) *ms_update = s->prof_ms[1];
    if (ms_direction) *ms_direction = s->prof_ms[2];
    if (iterations) *iterations = s->prof_count;
    return 0;
}
int dgb_pcg_last_folded(dgb_pcg* h, int* folded) {
    if (!h || !folded) { set_error("dgb_pcg_last_folded: NULL argument"); return DGB_ERR_INVALID; }
    *folded = reinterpret_cast<Pcg*>(h)->last_folded ? 1 : 0;
    return 0;
}
int dgb_pcg_destroy(dgb_pcg* h) {
    pcg_delete(reinterpret_cast<Pcg*>(h));
    return 0;
}
int dgb_pcg_solve_elliptic2d_dist(dgb_pcg* h, dgb_comm* comm, dgb_elliptic2d* A, double* x, const double* b, const double* P,
                                  const double* W, double eps, double nrmb_correction, int test_frequency, int max_iter,
                                  int* iterations, dgb_stream_t s) {
    if (!h || !A || !comm) { set_error("dgb_pcg_solve_elliptic2d_dist: NULL handle"); return DGB_ERR_INVALID; }
    return pcg_solve_impl(*reinterpret_cast<Pcg*>(h), reinterpret_cast<Comm*>(comm), *reinterpret_cast<Elliptic2dPlan*>(A), x,
                          b, P, W, eps, nrmb_correction, test_frequency, max_iter, iterations, as_stream(s));
}
int dgb_pcg_solve_elliptic2d(dgb_pcg* h, dgb_elliptic2d* A, double* x, const double* b, const double* P,
                             const double* W, double eps, double nrmb_correction, int test_frequency, int max_iter,
                             int* iterations, dgb_stream_t s) {
    if (!h || !A) { set_error("dgb_pcg_solve_elliptic2d: NULL handle"); return DGB_ERR_INVALID; }
    return pcg_solve(*reinterpret_cast<Pcg*>(h), *reinterpret_cast<Elliptic2dPlan*>(A), x, b, P, W, eps,
                     nrmb_correction, test_frequency, max_iter, iterations, as_stream(s));
}
}
