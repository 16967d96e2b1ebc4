// Fused Elliptic2d apply, second generation: warp-private sliding window ("walker").
//
// Why: the CTA-tile kernel (elliptic_fused.cu) moves the algorithmic minimum through HBM but is latency bound -- four
// CTA-wide barriers per tile, fluxes through shared memory, 25 % occupancy (ncu: profiles/ncu_r01_pcg_kernels_before.md).
// The operator costs ~62 DFMA per dof in the reference's rounding order, so on B200 (64 FP64 lanes/SM/clk) the FP64
// pipe time (~36 us at n=3, 1024^2) equals the HBM time (~35 us): the kernel has to keep the FP64 pipe fed and nothing
// else.  Design:
//   * one WARP owns a strip of UL = 32 - 2 HL cell columns (HL halo lanes on either side) and walks a range of cell
//     rows in y; lane l owns cell column c0 - HL + l, i.e. the n x n nodal values of one cell per step.  The host cuts
//     the (column, row) space into one contiguous, cost-weighted piece per warp (boundary columns are more expensive),
//     so every warp of the persistent grid finishes at the same time;
//   * the warp runs its own TMA pipeline: per cell row one box (n rows x (32 n + 2) doubles) of x, sigma (and the
//     weights of the fused dot) lands in a warp-private ring of shared-memory slots, completion on warp-private
//     mbarriers; lane 0 issues, all lanes wait.  The producer side runs ahead of the consumer across task
//     boundaries (the rings are FIFOs).  No __syncthreads anywhere in the main loop;
//   * the y-fluxes GY of the previous / next cell rows stay in REGISTERS (sliding window), the x-flux GX of the
//     neighbour cell comes by warp shuffle; stencil operands are read from the ring with conflict-free 8-byte LDS;
//   * the result row is staged in shared memory and leaves by TMA store (cp.async.bulk.tensor ... shared -> global),
//     double buffered; the optional exact dot(x, w, y) of PCG rides along in a per-lane floating-point expansion;
//   * warps whose columns touch a periodic seam in x (TMA cannot wrap), or operands TMA cannot describe, fill the
//     same ring with LDGSTS (cp.async) instead -- the compute code is identical;
//   * the arithmetic is the reference's rounding sequence (see elliptic_fused.cu header; inc/dg/elliptic.h:431-458,
//     inc/dg/backend/sparseblockmat_omp_kernels.h:36-50, inc/dg/topology/multiply.h:18-32): results are bitwise those
//     of the tile kernel and of the reference's OpenMP backend.
#pragma once
#include "elliptic_dev.cuh"
#include <vector>
#include <type_traits>

namespace dgb {

#ifndef DGB_WALK_WARPS
#define DGB_WALK_WARPS 12
#endif
#ifndef DGB_WALK_PD
#define DGB_WALK_PD 1
#endif
#ifndef DGB_WALK_WRING
#define DGB_WALK_WRING 1  // 1: the weights of the fused dot come through their own TMA ring; 0: lane-strided global loads
#endif                    // behind an L1 prefetch (frees two ring slots per warp; measured slower: 128 vs 122 us at 12 warps)
constexpr int WALK_MAX_WARPS = DGB_WALK_WARPS, WALK_PD = DGB_WALK_PD;
constexpr bool WALK_WRING = DGB_WALK_WRING != 0;
constexpr int WNOROW = -(1 << 30);
// keeps the compiler from hoisting the next phase's shared-memory loads above this point (register pressure)
#define DGB_PHASE_FENCE() asm volatile("" ::: "memory")
#ifndef DGB_WALK_NOFENCE_DOT
#define DGB_WALK_NOFENCE_DOT 0  // 1: the fused-dot variants (8 warps, 255 registers available) drop the phase fences
#endif
#ifndef DGB_WALK_FPE_PER_CELL
#define DGB_WALK_FPE_PER_CELL 0  // 1: one expansion per nodal value of the cell instead of one per column (one-sided stencils)
#endif

struct WalkArgs {
    MatView rx, ry, lx, ly, jx, jy;
    int Nx, Ny, wrapx, wrapy;        // Ny: rows of this slab (== global rows when not in slab mode); wrapy = 0 in slab mode
    int fx_lo, fx_hi, fy_lo, fy_hi;  // cells that are interior rows of all x- resp. y-matrices
    int slab, ghost, yoff, Nyg, pery;
    int tma_load, tma_store;
    const int4* tasks;               // (warp column, first cell row, end cell row, -)
    const int* tbegin;               // tasks of warp g: [tbegin[g], tbegin[g+1])
    const double* sigma;
    const double* vol;
    const double* x;
    const double* w;  // weights of the fused dot
    double* fold_pn;  // FOLD: new direction buffer (first row of the lower ghost block)
    double* y;
    double alpha, beta, jfactor;
    int helm;                 // GeneralHelmholtz epilogue: y = chi x - helm_alpha y (helmholtz.h:74-80)
    double helm_alpha;
    const double* helm_chi;
    sa::DotSlot slot;
    PcgState* pcg;
    P2pView p2p;               // multi-GPU: the finishing block exchanges the dot record over peer memory (pcg.cuh)
    unsigned long long epoch;
    int pdl;                   // host side only: launch with programmatic stream serialization
};

template <int N, int DIRK, bool DOT, bool FOLD = false>
struct WL {
    static constexpr int RK = DIRK, LK = DIRK == 0 ? 1 : (DIRK == 1 ? 0 : 2);
    static constexpr int HL = DIRK == 2 ? 2 : 1;            // halo lanes / halo cell rows
    static constexpr int UL = 32 - 2 * HL;                  // cell columns a warp produces
    static constexpr int WX = DIRK == 2 ? 2 : 1;            // newest x row a step needs: iy + WX
    static constexpr int LY = DIRK == 0 ? 0 : 1;            // the step at iy computes GY(iy + LY)
    static constexpr int LLO = Offs<LK>::first;             // oldest GY row a step needs: iy + LLO
    static constexpr int NGY = Offs<LK>::BPL;
    static constexpr int NP = NGY - 1;                      // steps before the first row that only build GY
    static constexpr int PD = WALK_PD;                      // prefetch distance in steps
    static constexpr int SX = WX + 2 + PD, SS = LY + 1 + PD, SW = 1 + PD;  // ring sizes (slots)
    // FOLD (PCG direction update in the loader): the old direction lands in its own ring; a row is combined with z the
    // moment its tick arrives, so only the PD rows in flight need a slot
    static constexpr int SP = FOLD ? (DIRK == 2 ? 1 : 2) : 0;  // (the centered variant has no shared memory left for a second one)
    static constexpr int RP = 32 * N + 2;                   // row pitch of a slot (doubles); +2: 16-byte aligned start
    static constexpr int SLOT = (N * RP * 8 + 127) / 128 * 128 / 8;       // doubles per slot (128-B aligned)
    static constexpr int OP = UL * N;                                      // row pitch of the output staging buffer
    static constexpr int OSLOT = (N * OP * 8 + 127) / 128 * 128 / 8;
    static constexpr int XOFF = 0, SOFF = XOFF + SX * SLOT, WOFF = SOFF + SS * SLOT, POFF = WOFF + ((DOT && WALK_WRING) ? SW : 0) * SLOT,
                         OOFF = POFF + SP * SLOT, BOFF = OOFF + OSLOT, WARP_DOUBLES = BOFF + 16;  // 16 doubles: up to 16 mbarriers
    static constexpr int FIT = (227 * 1024 - 1024) / (WARP_DOUBLES * 8);  // warps whose rings fit into one SM
    // warps come in multiples of 4 (one per scheduler, the register file is per scheduler): 12 warps leave 168
    // registers per thread, enough for the one-sided stencils; the centered one needs ~250 -> 8 warps
#ifndef DGB_WALK_DOT_WARPS
#define DGB_WALK_DOT_WARPS 8  // the fused-dot variant needs ~210 registers: 2 warps per scheduler
#endif
    static constexpr int WANT0 = (DIRK == 2 && N > 2 && WALK_MAX_WARPS > 8) ? 8 : WALK_MAX_WARPS;
    static constexpr int WANT = (DOT && WANT0 > DGB_WALK_DOT_WARPS) ? DGB_WALK_DOT_WARPS : WANT0;
    static constexpr int WARPS = FIT < WANT ? FIT : WANT, THREADS = 32 * WARPS;
    static constexpr size_t BYTES = (size_t)WARP_DOUBLES * 8 * WARPS;
    static_assert(SX <= 8, "mbarrier block / wait switch too small");
    static_assert(WARPS >= 1, "ring does not fit into shared memory");
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int K>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K) : "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(unsigned pending) {  // allow `pending` newest groups in flight
    switch (pending) {
        case 0: cp_async_wait<0>(); break;
        case 1: cp_async_wait<1>(); break;
        case 2: cp_async_wait<2>(); break;
        case 3: cp_async_wait<3>(); break;
        case 4: cp_async_wait<4>(); break;
        case 5: cp_async_wait<5>(); break;
        case 6: cp_async_wait<6>(); break;
        default: cp_async_wait<7>(); break;
    }
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_store_1d(double* dst, const double* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int K>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(K) : "memory"); }
template <int K>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(K) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ double shfl_up_d(double v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ double shfl_down_d(double v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// N independent "lines" through one block-row of M at once:
//   O(line, k) = fma(a, sum_q blk_d[k][q] * V_d(line, q), O(line, k))   over the slots d (slot order = reference order)
// V_d is the operand of the neighbour at the slot's offset (-1: vm, 0: v0, +1: vp; vm is never touched for KIND 0, vp
// never for KIND 1).  TR = false: lines are the first array index (x-direction stencils: line = ky, q/k = kx);
// TR = true: lines are the second index (y-direction: line = kx, q/k = ky).  Every coefficient is fetched once and
// used for all N lines.  FAST: `cell` is known to be an interior row, the blocks are uniform (constant-bank) operands.
template <int N, int KIND, bool FAST, bool TR>
__device__ __forceinline__ void stencil_lines(const MatView& M, const double (&C)[Offs<KIND>::BPL][N][N], int cell,
                                              const double (&vm)[N][N], const double (&v0)[N][N], const double (&vp)[N][N],
                                              double a, double (&out)[N][N]) {
    constexpr int B = Offs<KIND>::BPL;
    if (FAST || (cell >= M.i_lo && cell < M.i_hi)) {
#pragma unroll
        for (int d = 0; d < B; d++) {
            const int o = Offs<KIND>::first + d;  // compile-time after unrolling
            // N*N independent FMA chains (k, line) of length N (q): dependent DFMAs are N*N issue slots apart
            double t[N][N];
#pragma unroll
            for (int k = 0; k < N; k++)
#pragma unroll
                for (int l = 0; l < N; l++) t[k][l] = 0.;
#pragma unroll
            for (int q = 0; q < N; q++)
#pragma unroll
                for (int k = 0; k < N; k++) {
                    const double c = C[d][k][q];
#pragma unroll
                    for (int l = 0; l < N; l++) {
                        const double xv = o < 0 ? (TR ? vm[q][l] : vm[l][q]) : (o == 0 ? (TR ? v0[q][l] : v0[l][q]) : (TR ? vp[q][l] : vp[l][q]));
                        t[k][l] = __fma_rn(c, xv, t[k][l]);
                    }
                }
#pragma unroll
            for (int k = 0; k < N; k++)
#pragma unroll
                for (int l = 0; l < N; l++) {
                    if (TR) out[k][l] = __fma_rn(a, t[k][l], out[k][l]);
                    else out[l][k] = __fma_rn(a, t[k][l], out[l][k]);
                }
        }
    } else {
#pragma unroll 1
        for (int d = 0; d < B; d++) {
            const int col = M.cols[cell * B + d];
            if (col < 0) continue;
            int o = col - cell;  // boundary rows keep their own slot -> neighbour assignment (dx.h:85-97)
            if (o > 1) o -= M.num;
            else if (o < -1) o += M.num;
            const double* blk = M.data + (size_t)M.didx[cell * B + d] * N * N;
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t[N];
#pragma unroll
                for (int l = 0; l < N; l++) t[l] = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) {
                    const double c = __ldg(blk + k * N + q);
#pragma unroll
                    for (int l = 0; l < N; l++) {
                        const double xm = TR ? vm[q][l] : vm[l][q], x0 = TR ? v0[q][l] : v0[l][q], xp = TR ? vp[q][l] : vp[l][q];
                        t[l] = __fma_rn(c, o < 0 ? xm : (o == 0 ? x0 : xp), t[l]);
                    }
                }
#pragma unroll
                for (int l = 0; l < N; l++) {
                    if (TR) out[k][l] = __fma_rn(a, t[l], out[k][l]);
                    else out[l][k] = __fma_rn(a, t[l], out[l][k]);
                }
            }
        }
    }
}

// The same with the operands in shared memory: V_d(line, q) = p_d[line * LS + q * QS].  Each operand value is loaded right
// before its N uses, which keeps the live register set small (3 warps per scheduler fit).
template <int N, int KIND, bool FAST, int LS, int QS, bool TR>
__device__ __forceinline__ void stencil_mem(const MatView& M, const double (&C)[Offs<KIND>::BPL][N][N], int cell,
                                            const double* pm, const double* p0, const double* pp, double a, double (&out)[N][N]) {
    constexpr int B = Offs<KIND>::BPL;
    if (FAST || (cell >= M.i_lo && cell < M.i_hi)) {
#pragma unroll
        for (int d = 0; d < B; d++) {
            const int o = Offs<KIND>::first + d;  // compile-time after unrolling
            const double* p = o < 0 ? pm : (o == 0 ? p0 : pp);
            double t[N][N];
#pragma unroll
            for (int k = 0; k < N; k++)
#pragma unroll
                for (int l = 0; l < N; l++) t[k][l] = 0.;
#pragma unroll
            for (int q = 0; q < N; q++) {
                double xv[N];
#pragma unroll
                for (int l = 0; l < N; l++) xv[l] = p[l * LS + q * QS];
#pragma unroll
                for (int k = 0; k < N; k++) {
                    const double c = C[d][k][q];
#pragma unroll
                    for (int l = 0; l < N; l++) t[k][l] = __fma_rn(c, xv[l], t[k][l]);
                }
            }
#pragma unroll
            for (int k = 0; k < N; k++)
#pragma unroll
                for (int l = 0; l < N; l++) {
                    if (TR) out[k][l] = __fma_rn(a, t[k][l], out[k][l]);
                    else out[l][k] = __fma_rn(a, t[k][l], out[l][k]);
                }
        }
    } else {
#pragma unroll 1
        for (int d = 0; d < B; d++) {
            const int col = M.cols[cell * B + d];
            if (col < 0) continue;
            int o = col - cell;  // boundary rows keep their own slot -> neighbour assignment (dx.h:85-97)
            if (o > 1) o -= M.num;
            else if (o < -1) o += M.num;
            const double* blk = M.data + (size_t)M.didx[cell * B + d] * N * N;
            const double* p = o < 0 ? pm : (o == 0 ? p0 : pp);
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t[N];
#pragma unroll
                for (int l = 0; l < N; l++) t[l] = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) {
                    const double c = __ldg(blk + k * N + q);
#pragma unroll
                    for (int l = 0; l < N; l++) t[l] = __fma_rn(c, p[l * LS + q * QS], t[l]);
                }
#pragma unroll
                for (int l = 0; l < N; l++) {
                    if (TR) out[k][l] = __fma_rn(a, t[l], out[k][l]);
                    else out[l][k] = __fma_rn(a, t[l], out[l][k]);
                }
            }
        }
    }
}

// RELAXED ordering (dgb_elliptic2d_set_ordering, opt-in): the same stencils with ONE fused-multiply-add chain per output
// over all blocks and coefficients -- no per-block partial sum, no separate alpha step (the sign rides in the coefficient,
// jump blocks arrive pre-scaled by jfactor).  27 % fewer FP64 operations per cell (405 instead of 558 at n = 3); the result
// differs from the reference's rounding sequence in the last bits (<= 1e-14 relative, the north star allows 1e-12 for symv).
// Interior rows only (the boundary rows keep the reference sequence).
template <int N, int KIND, int LS, int QS, bool TR, int SGN>
__device__ __forceinline__ void stencil_mem_relaxed(const double (&C)[Offs<KIND>::BPL][N][N], const double* pm, const double* p0,
                                                    const double* pp, double (&out)[N][N]) {
    constexpr int B = Offs<KIND>::BPL;
#pragma unroll
    for (int d = 0; d < B; d++) {
        const int o = Offs<KIND>::first + d;
        const double* p = o < 0 ? pm : (o == 0 ? p0 : pp);
#pragma unroll
        for (int q = 0; q < N; q++) {
            double xv[N];
#pragma unroll
            for (int l = 0; l < N; l++) xv[l] = p[l * LS + q * QS];
#pragma unroll
            for (int k = 0; k < N; k++) {
                const double c = SGN > 0 ? C[d][k][q] : -C[d][k][q];
#pragma unroll
                for (int l = 0; l < N; l++) {
                    if (TR) out[k][l] = __fma_rn(c, xv[l], out[k][l]);
                    else out[l][k] = __fma_rn(c, xv[l], out[l][k]);
                }
            }
        }
    }
}
template <int N, int KIND, bool TR, int SGN>
__device__ __forceinline__ void stencil_lines_relaxed(const double (&C)[Offs<KIND>::BPL][N][N], const double (&vm)[N][N],
                                                      const double (&v0)[N][N], const double (&vp)[N][N], double (&out)[N][N]) {
    constexpr int B = Offs<KIND>::BPL;
#pragma unroll
    for (int d = 0; d < B; d++) {
        const int o = Offs<KIND>::first + d;
#pragma unroll
        for (int q = 0; q < N; q++)
#pragma unroll
            for (int k = 0; k < N; k++) {
                const double c = SGN > 0 ? C[d][k][q] : -C[d][k][q];
#pragma unroll
                for (int l = 0; l < N; l++) {
                    const double xv = o < 0 ? (TR ? vm[q][l] : vm[l][q]) : (o == 0 ? (TR ? v0[q][l] : v0[l][q]) : (TR ? vp[q][l] : vp[l][q]));
                    if (TR) out[k][l] = __fma_rn(c, xv, out[k][l]);
                    else out[l][k] = __fma_rn(c, xv, out[l][k]);
                }
            }
    }
}

// local cell row r (may lie outside the slab) -> row used for ADDRESSING x / sigma, WNOROW if no data exists
__device__ __forceinline__ int w_yaddr(int r, const WalkArgs& A) {
    if (!A.slab) { int g = gcell(r, A.Ny, A.wrapy); return g < 0 ? WNOROW : g; }
    if (r < -A.ghost || r >= A.Ny + A.ghost) return WNOROW;
    const int g = r + A.yoff;
    if (!A.pery && (g < 0 || g >= A.Nyg)) return WNOROW;
    return r;
}
// local cell row r -> block-row index of the (global) y-matrices, WNOROW if the row does not exist
__device__ __forceinline__ int w_ymat(int r, const WalkArgs& A) {
    if (!A.slab) { int g = gcell(r, A.Ny, A.wrapy); return g < 0 ? WNOROW : g; }
    int g = r + A.yoff;
    if (g < 0 || g >= A.Nyg) {
        if (!A.pery) return WNOROW;
        g = g < 0 ? g + A.Nyg : g - A.Nyg;
    }
    return g;
}

// LDGSTS loader of one slot: N rows x RP doubles starting at global element column `cs` of cell row r
template <int N, int RP>
__device__ __forceinline__ void fill_slot_ldgsts(double* dst, const double* src, int r, int cs, const WalkArgs& A, int lane) {
    const int LD = A.Nx * N;
    const int gy = w_yaddr(r, A);
    for (int e = lane; e < RP; e += 32) {
        const int ce = cs + e;
        int cell = ce >= 0 ? ce / N : -((-ce + N - 1) / N);
        const int sub = ce - cell * N;
        cell = gcell(cell, A.Nx, A.wrapx);
        const bool ok = cell >= 0 && gy != WNOROW;
#pragma unroll
        for (int k = 0; k < N; k++)
            cp_async8(dst + k * RP + e, ok ? src + ((long long)(gy * N + k) * LD + cell * N + sub) : src, ok);
    }
}

template <int N, int RP>
__device__ __forceinline__ void ld_cell(const double* slot, int e, double (&dst)[N][N]) {
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < N; b++) dst[a][b] = slot[a * RP + e + b];
}

// PLAIN: alpha-only epilogue (beta == 0, no volume form, no Helmholtz term) known at compile time -- the hot variants
// ALLTMA: every load and store of this launch goes through TMA (no periodic seam in x, all operands describable) -- the
// LDGSTS / direct-store alternatives are compiled out
// RELAX: the interior rows use the relaxed operation order above (only instantiated for the PLAIN, ALLTMA variants)
// FOLD (only with DOT, ALLTMA): the PCG direction update p = z + beta p (pcg.h:182) happens in the loader -- map_x describes
// z, map_p the old direction, every row is combined in its ring slot when it lands (the operator then runs on the new
// direction) and leaves through map_pn into the OTHER direction buffer (strip halos are written by both neighbours with
// identical values; the old direction is never modified, so no warp can see a half-updated halo)
template <int N, int DIRK, bool DOT, bool PLAIN, bool ALLTMA, bool RELAX = false, bool FOLD = false>
__global__ void __launch_bounds__((WL<N, DIRK, DOT, FOLD>::THREADS), 1)
elliptic2d_walker_kernel(const __grid_constant__ WalkArgs A, const __grid_constant__ EllipticCoef<N, Offs<DIRK>::BPL> C,
                         const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_s,
                         const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_y,
                         const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_pn) {
    static_assert(!FOLD || (DOT && ALLTMA && !RELAX), "the folded direction update exists for the all-TMA fused-dot variant");
    using L = WL<N, DIRK, DOT, FOLD>;
    constexpr int RK = L::RK, LK = L::LK, HL = L::HL, UL = L::UL, WX = L::WX, LY = L::LY, LLO = L::LLO, NGY = L::NGY,
                  NP = L::NP, SX = L::SX, SS = L::SS, SW = L::SW, SP = FOLD ? L::SP : 1, RP = L::RP, SLOT = L::SLOT, OP = L::OP, WALK_WARPS = L::WARPS;
    constexpr unsigned SLOT_BYTES = N * RP * 8;
    __shared__ long long dsm[DOT ? sa::BINS : 1];  // one accumulator per block
    extern __shared__ __align__(128) double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wb = smem + (size_t)warp * L::WARP_DOUBLES;
    double* Xr = wb + L::XOFF;
    double* Sr = wb + L::SOFF;
    double* Wr = wb + L::WOFF;
    double* Pr = wb + L::POFF;
    double* ob = wb + L::OOFF;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(wb + L::BOFF);
    // N independent two-term expansions per lane: the fused dot rides in the FP64 pipe this kernel is bound by, so it
    // uses the short expansion and keeps the add cascades of one cell row independent of each other
    constexpr int NE = (DGB_WALK_FPE_PER_CELL && DIRK != 2) ? N * N : N;
    sa::FpeT<2> fpe[NE];
#define WPHASE() do { if (!(DGB_WALK_NOFENCE_DOT && DOT)) DGB_PHASE_FENCE(); } while (0)
    int bad = 0;
    // prologue that touches nothing a preceding kernel produces (shared memory, mbarriers, the host-built task tables): with a
    // programmatic dependent launch it runs while the predecessor's finishing block is still at work (common.cuh)
    if (DOT) {
        sa::block_init<1>(dsm);
#pragma unroll
        for (int k = 0; k < NE; k++) fpe[k].clear();
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < SX; s++) mbar_init(bar + s, 1);
    }
    __syncwarp();
    const int LD = A.Nx * N;
    const int gwarp = blockIdx.x * WALK_WARPS + warp;
    const int tb = A.tbegin[gwarp], te = A.tbegin[gwarp + 1];
    pdl_wait();
    if (DOT) {
        if (A.pcg->done) return;  // solver already converged: the remaining launches of the batch are no-ops
    }
    const double fbeta = FOLD ? A.pcg->beta : 0.;

    // ---- producer: the rings are FIFOs; xp/sp/wp count pushed rows, xr/sr/wr released ones, tickp issued ticks.
    // Tick j of a task carries x row iy0 - HL + j and, where they exist, sigma row (x row) - WX + LY and w row (x row) - WX.
    unsigned xp = 0, sp = 0, wp = 0, tickp = 0, xr = 0, sr = 0, wr = 0, pr = 0, phases = 0;
    int pt = tb, pj = 0, p_iy0 = 0, p_iy1 = 0, p_cs = 0, p_manual = 0;
    auto ptask = [&]() {
        if (pt >= te) return;
        const int4 t = A.tasks[pt];
        const int cl = t.x * UL - HL;
        p_iy0 = t.y; p_iy1 = t.z;
        p_cs = cl * N - ((cl * N) & 1);
        p_manual = ALLTMA ? 0 : (!A.tma_load || (A.wrapx && (cl < 0 || cl + 32 > A.Nx)));
    };
    ptask();
    auto produce = [&]() {
        while (pt < te) {
            const int r = p_iy0 - HL + pj, rs = r - WX + LY, rw = r - WX;
            const bool needS = rs >= p_iy0 - NP + LY && rs <= p_iy1 - 1 + LY;
            const bool needW = DOT && WALK_WRING && rw >= p_iy0 && rw <= p_iy1 - 1;
            if (xp - xr >= (unsigned)SX || (needS && sp - sr >= (unsigned)SS) || (needW && wp - wr >= (unsigned)SW)) break;
            if (FOLD && tickp - pr >= (unsigned)SP) break;
            double* dx = Xr + (xp % SX) * SLOT;
            double* ds = Sr + (sp % SS) * SLOT;
            double* dw = Wr + (wp % SW) * SLOT;
            if (p_manual) {
                fill_slot_ldgsts<N, RP>(dx, A.x, r, p_cs, A, lane);
                if (needS) fill_slot_ldgsts<N, RP>(ds, A.sigma, rs, p_cs, A, lane);
                if (needW) fill_slot_ldgsts<N, RP>(dw, A.w, rw, p_cs, A, lane);
            } else if (lane == 0) {
                unsigned long long* b = bar + (tickp % SX);
                mbar_expect_tx(b, SLOT_BYTES * ((FOLD ? 2u : 1u) + (needS ? 1u : 0u) + (needW ? 1u : 0u)));
                // rows outside a non-periodic domain are zero-filled by TMA and never used; periodic rows wrap here
                const int yr = A.slab ? r + A.ghost : (A.wrapy ? (r < 0 ? r + A.Ny : (r >= A.Ny ? r - A.Ny : r)) : r);
                tma_load_2d(dx, &map_x, b, p_cs, yr * N);
                if (FOLD) tma_load_2d(Pr + (tickp % SP) * SLOT, &map_p, b, p_cs, yr * N);
                if (needS) {
                    const int ys = A.slab ? rs + A.ghost : (A.wrapy ? (rs < 0 ? rs + A.Ny : (rs >= A.Ny ? rs - A.Ny : rs)) : rs);
                    tma_load_2d(ds, &map_s, b, p_cs, ys * N);
                }
                if (needW) tma_load_2d(dw, &map_w, b, p_cs, rw * N);  // w has no ghost rows
            }
            if (!ALLTMA) cp_async_commit();  // one (possibly empty) LDGSTS group per tick: group index == tick index
            xp++; sp += needS ? 1u : 0u; wp += needW ? 1u : 0u; tickp++;
            if (++pj == p_iy1 - p_iy0 + HL + WX) { pt++; pj = 0; ptask(); }
        }
    };

    // ---- consumer
    unsigned xc = 0, sc = 0, wc = 0, tickc = 0;  // rows / ticks consumed by the finished tasks
    for (int ct = tb; ct < te; ct++) {
        const int4 task = A.tasks[ct];
        const int c0 = task.x * UL, iy0 = task.y, iy1 = task.z;
        const int cl = c0 - HL;                              // cell column of lane 0
        const int sh = (cl * N) & 1;                         // the box starts at an even element column
        const bool manual = ALLTMA ? false : (!A.tma_load || (A.wrapx && (cl < 0 || cl + 32 > A.Nx)));
        const int fx = iy0 - HL, nticks = iy1 - iy0 + HL + WX, s0row = iy0 - NP + LY;
        const int gx = gcell(cl + lane, A.Nx, A.wrapx);      // my cell column (wrapped), -1 if it does not exist
        const bool outlane = lane >= HL && lane < HL + UL && cl + lane < A.Nx;
        const bool fastx = cl >= A.fx_lo && cl + 32 <= A.fx_hi;
        const int eo = lane * N + sh, el = max(lane - 1, 0) * N + sh, er = min(lane + 1, 31) * N + sh;
        auto xrow = [&](int r) -> const double* { return Xr + ((xc + (unsigned)(r - fx)) % SX) * SLOT; };
        auto srow = [&](int r) -> const double* { return Sr + ((sc + (unsigned)(r - s0row)) % SS) * SLOT; };
        auto wrow = [&](int r) -> const double* { return Wr + ((wc + (unsigned)(r - iy0)) % SW) * SLOT; };
        produce();

        double gy[NGY][N][N];  // gy[i] = GY(iy + LLO + i), [ky][kx]
        int waited = 0;
        for (int iy = iy0 - NP; iy < iy1; iy++) {
            // ---- wait for the ticks up to j(iy) = iy - iy0 + HL + WX
            for (const int need = iy - iy0 + HL + WX; waited <= need; waited++) {
                const unsigned t = tickc + (unsigned)waited;
                if (manual) {
                    cp_async_wait_pending(tickp - 1u - t);
                    __syncwarp();
                } else {
                    const unsigned s = t % SX;
                    mbar_wait(bar + s, (phases >> s) & 1u);
                    phases ^= 1u << s;
                }
                if (FOLD) {
                    // the row that just landed: new direction = fma(1, z, old * beta) (Axpby functor, subroutines.h:260-274),
                    // in place in the x ring; rows this piece owns (and the ghost rows at a slab edge) leave by TMA
                    double* zs = Xr + ((xc + (unsigned)waited) % SX) * SLOT;
                    const double* ps = Pr + (t % SP) * SLOT;
                    constexpr int TRIPS = (N * RP / 2 + 31) / 32;
                    double2 zv[TRIPS], pv[TRIPS];  // all loads first: the ring slots may alias as far as the compiler knows
#pragma unroll
                    for (int u = 0; u < TRIPS; u++) {
                        const int e = 2 * lane + 64 * u;
                        if (u + 1 < TRIPS || e < N * RP) {
                            zv[u] = *reinterpret_cast<const double2*>(zs + e);
                            pv[u] = *reinterpret_cast<const double2*>(ps + e);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < TRIPS; u++) {
                        const int e = 2 * lane + 64 * u;
                        if (u + 1 < TRIPS || e < N * RP) {
                            pv[u].x = __fma_rn(1., zv[u].x, __dmul_rn(pv[u].x, fbeta));
                            pv[u].y = __fma_rn(1., zv[u].y, __dmul_rn(pv[u].y, fbeta));
                            *reinterpret_cast<double2*>(zs + e) = pv[u];
                        }
                    }
                    fence_async_smem();
                    __syncwarp();
                    const int r = fx + waited;
                    const bool mine = (r >= iy0 && r < iy1) || (A.slab && ((iy0 == 0 && r < 0) || (iy1 == A.Ny && r >= A.Ny)));
                    const int pcs = cl * N - sh, yrow = (A.slab ? r + A.ghost : r) * N;
                    if (mine && pcs >= 0) {
                        if (lane == 0) {
                            tma_store_2d(&map_pn, zs, pcs, yrow);
                            bulk_commit();
                        }
                    } else if (mine) {
                        // leftmost strip: the box starts left of the array and a TMA tensor store must not have a negative
                        // start coordinate -- one 1-d bulk copy per row instead (same bulk async-group)
                        if (lane == 0) {
                            const int off = -pcs, len = min(RP, LD - pcs) - off;  // pcs is even: 16-byte aligned pieces
#pragma unroll
                            for (int k = 0; k < N; k++)
                                bulk_store_1d(A.fold_pn + (size_t)(yrow + k) * LD, zs + k * RP + off, (unsigned)len * 8u);
                            bulk_commit();
                        }
                    }
                    pr = t + 1u;
                    produce();  // the freed slot lets the next row of the old direction start
                }
            }
            const bool fasty = iy + A.yoff - HL >= A.fy_lo && iy + A.yoff + WX + 1 <= A.fy_hi;
            const bool fast = fastx && fasty;
            const bool emit = iy >= iy0;

            double acc[N][N];  // [ky][kx] the cell's outputs (valid when emit)
            // the whole stencil part of the step exists twice, FAST (interior rows everywhere, constant-bank blocks) and
            // general, selected by ONE warp-uniform branch
            auto stencils = [&](auto fast_tag) {
                constexpr bool FAST = decltype(fast_tag)::value;
            // ---- GY(R) = sigma(R) * (Ry x)(R), R = iy + LY: operands are the cell's own column in rows R-1, R, R+1
            {
                const int R = iy + LY;
                const double* pa = xrow(RK != 0 ? R - 1 : R) + eo;
                const double* pb = xrow(R) + eo;
                const double* pc = xrow(RK != 1 ? R + 1 : R) + eo;
                double g[N][N];
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int b = 0; b < N; b++) g[a][b] = 0.;
                bool ong = true;
                if (FAST && RELAX) {
                    stencil_mem_relaxed<N, RK, 1, RP, true, 1>(C.ry, pa, pb, pc, g);
                } else if (FAST) {
                    stencil_mem<N, RK, true, 1, RP, true>(A.ry, C.ry, 0, pa, pb, pc, 1., g);
                } else {
                    const int my = w_ymat(R, A);
                    ong = gx >= 0 && my != WNOROW;
                    if (ong) stencil_mem<N, RK, false, 1, RP, true>(A.ry, C.ry, my, pa, pb, pc, 1., g);
                }
                if (ong) {
                    double SG[N][N];
                    ld_cell<N, RP>(srow(R), eo, SG);
#pragma unroll
                    for (int a = 0; a < N; a++)
#pragma unroll
                        for (int b = 0; b < N; b++)
                            g[a][b] = (FAST && RELAX) ? __dmul_rn(SG[a][b], g[a][b]) : __fma_rn(SG[a][b], g[a][b], __dmul_rn(g[a][b], 0.));
                }
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int b = 0; b < N; b++) gy[NGY - 1][a][b] = g[a][b];
            }
            WPHASE();
            if (emit) {
                const double* x0 = xrow(iy);
                // beta != 0 / curvilinear volume: the epilogue reads y / vol with lane-strided loads; pull the lines into L1
                // now so that the latency is gone by then (hot loops use beta == 0, vol == nullptr)
                if (DOT && !WALK_WRING && outlane) {  // weights of the fused dot: same treatment
                    const size_t gp = (size_t)(iy * N) * LD + (size_t)gx * N;
#pragma unroll
                    for (int ky = 0; ky < N; ky++) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.w + gp + (size_t)ky * LD));
                }
                if (!PLAIN && (A.beta != 0. || A.vol != nullptr) && outlane) {
                    const size_t gp = (size_t)(iy * N) * LD + (size_t)gx * N;
#pragma unroll
                    for (int ky = 0; ky < N; ky++) {
                        if (A.beta != 0.) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.y + gp + (size_t)ky * LD));
                        if (A.vol != nullptr) asm volatile("prefetch.global.L1 [%0];" ::"l"(A.vol + gp + (size_t)ky * LD));
                    }
                }
                const int ym = FAST ? 0 : w_ymat(iy, A);
                const bool on = FAST || gx >= 0;
                // ---- GX(iy) for my cell, then the neighbours' by shuffle
                double gxv[N][N], gxm[N][N], gxp[N][N];
                {
#pragma unroll
                    for (int a = 0; a < N; a++)
#pragma unroll
                        for (int b = 0; b < N; b++) gxv[a][b] = 0.;
                    if (on) {
                        if (FAST && RELAX) stencil_mem_relaxed<N, RK, RP, 1, false, 1>(C.rx, x0 + el, x0 + eo, x0 + er, gxv);
                        else if (FAST) stencil_mem<N, RK, true, RP, 1, false>(A.rx, C.rx, 0, x0 + el, x0 + eo, x0 + er, 1., gxv);
                        else stencil_mem<N, RK, false, RP, 1, false>(A.rx, C.rx, gx, x0 + el, x0 + eo, x0 + er, 1., gxv);
                        double S0[N][N];
                        ld_cell<N, RP>(srow(iy), eo, S0);
#pragma unroll
                        for (int a = 0; a < N; a++)
#pragma unroll
                            for (int b = 0; b < N; b++)
                                gxv[a][b] = (FAST && RELAX) ? __dmul_rn(S0[a][b], gxv[a][b]) : __fma_rn(S0[a][b], gxv[a][b], __dmul_rn(gxv[a][b], 0.));
                    }
#pragma unroll
                    for (int a = 0; a < N; a++)
#pragma unroll
                        for (int b = 0; b < N; b++) {
                            gxm[a][b] = LK != 0 ? shfl_up_d(gxv[a][b]) : 0.;
                            gxp[a][b] = LK != 1 ? shfl_down_d(gxv[a][b]) : 0.;
                        }
                }
                WPHASE();
                // ---- the cell's n x n outputs
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int b = 0; b < N; b++) acc[a][b] = 0.;
                if (on && FAST && RELAX) {
                    // acc = -(Ly GY) - (Lx GX) + jfactor ((Jx + Jy) x): one chain per output for the fluxes, one for the jumps
                    stencil_lines_relaxed<N, LK, true, -1>(C.ly, gy[LK == 0 ? 0 : -1 - LLO], gy[-LLO], gy[LK == 1 ? -LLO : 1 - LLO], acc);
                    stencil_lines_relaxed<N, LK, false, -1>(C.lx, gxm, gxv, gxp, acc);
                    WPHASE();
                    if (A.jfactor != 0.) {
                        double accj[N][N];
#pragma unroll
                        for (int a = 0; a < N; a++)
#pragma unroll
                            for (int b = 0; b < N; b++) accj[a][b] = 0.;
                        stencil_mem_relaxed<N, 2, RP, 1, false, 1>(C.jx, x0 + el, x0 + eo, x0 + er, accj);
                        WPHASE();
                        stencil_mem_relaxed<N, 2, 1, RP, true, 1>(C.jy, xrow(iy - 1) + eo, x0 + eo, xrow(iy + 1) + eo, accj);
#pragma unroll
                        for (int a = 0; a < N; a++)
#pragma unroll
                            for (int b = 0; b < N; b++) acc[a][b] = __fma_rn(A.jfactor, accj[a][b], acc[a][b]);
                    }
                } else if (on) {
                    // Ly ty (alpha = 1, beta = 0): GY rows iy-1, iy, iy+1 = gy[-1-LLO], gy[-LLO], gy[1-LLO]
                    if (FAST) stencil_lines<N, LK, true, true>(A.ly, C.ly, 0, gy[LK == 0 ? 0 : -1 - LLO], gy[-LLO], gy[LK == 1 ? -LLO : 1 - LLO], 1., acc);
                        else stencil_lines<N, LK, false, true>(A.ly, C.ly, ym, gy[LK == 0 ? 0 : -1 - LLO], gy[-LLO], gy[LK == 1 ? -LLO : 1 - LLO], 1., acc);
                    // - Lx tx - t   (alpha = -1, beta = -1)
#pragma unroll
                    for (int a = 0; a < N; a++)
#pragma unroll
                        for (int b = 0; b < N; b++) acc[a][b] = __dmul_rn(acc[a][b], -1.);
                    if (FAST) stencil_lines<N, LK, true, false>(A.lx, C.lx, 0, gxm, gxv, gxp, -1., acc);
                        else stencil_lines<N, LK, false, false>(A.lx, C.lx, gx, gxm, gxv, gxp, -1., acc);
                    WPHASE();
                    if (A.jfactor != 0.) {
                        if (FAST) stencil_mem<N, 2, true, RP, 1, false>(A.jx, C.jx, 0, x0 + el, x0 + eo, x0 + er, A.jfactor, acc);
                        else stencil_mem<N, 2, false, RP, 1, false>(A.jx, C.jx, gx, x0 + el, x0 + eo, x0 + er, A.jfactor, acc);
                        WPHASE();
                        const double* xd = xrow(iy - 1) + eo;
                        const double* xu = xrow(iy + 1) + eo;
                        if (FAST) stencil_mem<N, 2, true, 1, RP, true>(A.jy, C.jy, 0, xd, x0 + eo, xu, A.jfactor, acc);
                        else stencil_mem<N, 2, false, 1, RP, true>(A.jy, C.jy, ym, xd, x0 + eo, xu, A.jfactor, acc);
                    }
                }
                WPHASE();
                }
            };
            if (fast) stencils(std::true_type{});
            else stencils(std::false_type{});
            if (emit) {
                const double* x0 = xrow(iy);
                // ---- epilogue  y = fma(alpha, t/vol, beta*y)  (+ the exact dot); staged for the TMA store
                if (ALLTMA || A.tma_store) {
                    if (lane == 0) bulk_wait_read<0>();  // the previous store has read the staging buffer
                    __syncwarp();
                }
                if (outlane) {
                    const size_t gb = (size_t)(iy * N) * LD + (size_t)gx * N;
                    if (!PLAIN && (A.vol != nullptr || A.beta != 0.)) {
                        double yo[N][N], vo[N][N];  // all loads first: their latencies overlap
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) {
                                const size_t g = gb + (size_t)ky * LD + kx;
                                yo[ky][kx] = A.beta == 0. ? 0. : A.y[g];
                                vo[ky][kx] = A.vol ? __ldg(A.vol + g) : 1.;
                            }
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) {
                                double t = acc[ky][kx];
                                if (A.vol) t = __ddiv_rn(t, vo[ky][kx]);
                                const double b = A.beta == 0. ? 0. : __dmul_rn(yo[ky][kx], A.beta);
                                acc[ky][kx] = __fma_rn(A.alpha, t, b);
                            }
                    } else {
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) acc[ky][kx] = __fma_rn(A.alpha, acc[ky][kx], 0.);
                    }
                    if (!PLAIN && A.helm) {  // pointwiseDot(1., chi, x, -helm_alpha, y): y *= -helm_alpha; y = fma(1*chi, x, y)
                        const double mha = -A.helm_alpha;
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) {
                                const double c = A.helm_chi ? __ldg(A.helm_chi + gb + (size_t)ky * LD + kx) : 1.;
                                acc[ky][kx] = __fma_rn(__dmul_rn(1., c), x0[ky * RP + eo + kx], __dmul_rn(acc[ky][kx], mha));
                            }
                    }
                    if (ALLTMA || A.tma_store) {
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) ob[ky * OP + (lane - HL) * N + kx] = acc[ky][kx];
                    } else {
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) A.y[gb + (size_t)ky * LD + kx] = acc[ky][kx];
                    }
                    if (DOT) {
                        double WV[N][N], XC[N][N];
                        if (WALK_WRING) {
                            ld_cell<N, RP>(wrow(iy), eo, WV);
                        } else {
#pragma unroll
                            for (int ky = 0; ky < N; ky++)
#pragma unroll
                                for (int kx = 0; kx < N; kx++) WV[ky][kx] = __ldg(A.w + gb + (size_t)ky * LD + kx);
                        }
                        ld_cell<N, RP>(x0, eo, XC);
                        double res[N][N];
                        bool spill = false;
#pragma unroll
                        for (int ky = 0; ky < N; ky++)
#pragma unroll
                            for (int kx = 0; kx < N; kx++) {
                                double pr = __dmul_rn(__dmul_rn(XC[ky][kx], WV[ky][kx]), acc[ky][kx]);
                                if (!isfinite(pr)) { bad = 1; pr = 0.; }
                                res[ky][kx] = fpe[NE == N ? kx : ky * N + kx].add_lazy(pr);
                                spill = spill || res[ky][kx] != 0.;
                            }
                        if (spill) {  // rare: residues the expansions cannot hold go to the shared accumulator (exact)
#pragma unroll 1
                            for (int k = 0; k < N * N; k++) sa::accumulate(dsm, res[k / N][k % N], 1);
                        }
                    }
                }
                if (ALLTMA || A.tma_store) {
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_y, ob, c0 * N, iy * N);
                        bulk_commit();
                    }
                }
            }
            if (FOLD && !emit) {  // the row stores of this step have read their slots before the producer may refill them
                if (lane == 0) bulk_wait_read<0>();
            }
            // ---- slide the GY window, release the dead rows, refill the rings
#pragma unroll
            for (int i = 0; i < NGY - 1; i++)
#pragma unroll
                for (int a = 0; a < N; a++)
#pragma unroll
                    for (int b = 0; b < N; b++) gy[i][a][b] = gy[i + 1][a][b];
            __syncwarp();
            xr = xc + (unsigned)(iy - fx);
            sr = sc + (unsigned)(iy - s0row + 1);
            if (emit) wr = wc + (unsigned)(iy - iy0 + 1);
            if (iy + 1 == iy1) {  // task done: everything it loaded is dead
                xc += (unsigned)nticks; sc += (unsigned)(iy1 - iy0 + NP); wc += (unsigned)(iy1 - iy0); tickc += (unsigned)nticks;
                xr = xc; sr = sc; wr = wc;
            }
            produce();
        }
    }
    if (!ALLTMA) cp_async_wait<0>();
    if ((ALLTMA || A.tma_store) && lane == 0) bulk_wait<0>();
    // this warp's rows are done: from here on only the exact-dot tail runs.  Allowing the dependent kernel in now (and not at the
    // top: its blocks would sit beside ours through the main loop, measured 9 % slower) overlaps its launch and prologue with it
    pdl_trigger();
    if (DOT) {
#pragma unroll
        for (int k = 1; k < NE; k++) fpe[0].merge(fpe[k], dsm);
#undef WPHASE
        fpe[0].flush_warp(dsm);
        fused_dot_finish(sa::block_finish<1>(dsm, bad, A.slot, 0), A.pcg, A.slot.result, A.p2p, A.epoch);
    }
}

// ------------------------------------------------------------------------------------------------ host (launch)
struct WalkPartition {
    int Nx = -1, Ny = -1, nwarps = -1, UL = -1, key = -1;
    int4* d_tasks = nullptr;
    int* d_tbegin = nullptr;
    int ntasks = 0;
};
// elliptic_walker.cu
int build_partition(WalkPartition& P, int Nx, int Ny, int UL, int HL, int nwarps, int fx_lo, int fx_hi, bool wrapx,
                    bool tma, int min_rows, bool dot, cudaStream_t st);

template <int N, int DIRK, bool DOT, bool PLAIN, bool ALLTMA, bool RELAX = false, bool FOLD = false>
static int wlaunch_go(const WalkArgs& A, const EllipticCoef<N, Offs<DIRK>::BPL>& C, const CUtensorMap& mx, const CUtensorMap& ms,
                      const CUtensorMap& mw, const CUtensorMap& my, const CUtensorMap& mp, const CUtensorMap& mpn, int grid,
                      cudaStream_t st) {
    using L = WL<N, DIRK, DOT, FOLD>;
    static bool configured = false;
    if (!configured) {
        DGB_CUDA(cudaFuncSetAttribute(elliptic2d_walker_kernel<N, DIRK, DOT, PLAIN, ALLTMA, RELAX, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::BYTES));
        configured = true;
    }
    if (A.pdl)
        DGB_CUDA(launch_pdl(elliptic2d_walker_kernel<N, DIRK, DOT, PLAIN, ALLTMA, RELAX, FOLD>, dim3(grid), dim3(L::THREADS), L::BYTES, st, A, C, mx, ms, mw, my, mp, mpn));
    else
        elliptic2d_walker_kernel<N, DIRK, DOT, PLAIN, ALLTMA, RELAX, FOLD><<<grid, L::THREADS, L::BYTES, st>>>(A, C, mx, ms, mw, my, mp, mpn);
    DGB_LAUNCHED();
    return 0;
}

// FOLD: x is z (the preconditioned residual), p_old / p_new the two direction buffers; all three carry the ghost rows of a
// slab like x does
template <int N, int DIRK, bool DOT, bool PLAIN, bool FOLD = false>
static int wlaunch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st, const FusedDot* fd,
                   const double* p_old = nullptr, double* p_new = nullptr) {
    constexpr int B = Offs<DIRK>::BPL;
    using L = WL<N, DIRK, DOT, FOLD>;
    static int no_tma = -1;
    if (no_tma < 0) { const char* e = getenv("DGB_NO_TMA"); no_tma = (e && atoi(e)) ? 1 : 0; }
    WalkArgs A;
    A.rx = view(p.rightx); A.ry = view(p.righty); A.lx = view(p.leftx); A.ly = view(p.lefty);
    A.jx = view(p.jumpx); A.jy = view(p.jumpy);
    A.Nx = p.Nx; A.Ny = p.slab ? p.slab_rows : p.Ny; A.wrapx = p.wrapx; A.wrapy = p.slab ? 0 : p.wrapy;
    A.slab = p.slab; A.ghost = p.slab ? p.slab_ghost : 0; A.yoff = p.slab ? p.slab_yoff : 0; A.Nyg = p.Ny; A.pery = p.wrapy;
    if (p.slab && p.slab_ghost < L::HL) {
        set_error("elliptic2d slab: %d ghost cell rows given, the stencil needs %d", p.slab_ghost, L::HL);
        return DGB_ERR_INVALID;
    }
    A.fx_lo = std::max({p.rightx.i_lo, p.leftx.i_lo, p.jumpx.i_lo});
    A.fx_hi = std::min({p.rightx.i_hi, p.leftx.i_hi, p.jumpx.i_hi});
    A.fy_lo = std::max({p.righty.i_lo, p.lefty.i_lo, p.jumpy.i_lo});
    A.fy_hi = std::min({p.righty.i_hi, p.lefty.i_hi, p.jumpy.i_hi});
    A.sigma = p.sigma; A.vol = p.vol; A.x = x; A.y = y; A.w = nullptr;
    A.alpha = alpha; A.beta = beta; A.jfactor = p.jfactor;
    A.helm = p.helm ? 1 : 0; A.helm_alpha = p.helm_alpha; A.helm_chi = p.helm_chi;
    A.pcg = nullptr; A.slot = sa::DotSlot{nullptr, nullptr, nullptr, nullptr};
    A.p2p = P2pView{}; A.p2p.enabled = 0; A.epoch = 0;
    A.pdl = 0;
    if (DOT) { A.w = fd->w; A.slot = fd->slot; A.pcg = fd->pcg; A.p2p = fd->p2p; A.epoch = fd->epoch; A.pdl = fd->pdl; }
    CUtensorMap mx, ms, mw, my, mp, mpn;
    memset(&mx, 0, sizeof(mx)); memset(&ms, 0, sizeof(ms)); memset(&mw, 0, sizeof(mw)); memset(&my, 0, sizeof(my));
    memset(&mp, 0, sizeof(mp)); memset(&mpn, 0, sizeof(mpn));
    const long long gh = (long long)A.ghost * N * p.Nx * N;  // doubles in the ghost rows below the slab
    const int ld = p.Nx * N;
    A.tma_load = !no_tma && make_map(&mx, x - gh, (A.Ny + 2 * A.ghost) * N, ld, N, L::RP) &&
                 make_map(&ms, p.sigma - gh, (A.Ny + 2 * A.ghost) * N, ld, N, L::RP) &&
                 (!DOT || !WALK_WRING || make_map(&mw, A.w, A.Ny * N, ld, N, L::RP));
    A.tma_store = !no_tma && make_map(&my, y, A.Ny * N, ld, N, L::OP);
    A.fold_pn = nullptr;
    if (FOLD) {
        A.fold_pn = p_new - gh;
        const bool ok = A.tma_load && A.tma_store && !A.wrapx && make_map(&mp, p_old - gh, (A.Ny + 2 * A.ghost) * N, ld, N, L::RP) &&
                        make_map(&mpn, p_new - gh, (A.Ny + 2 * A.ghost) * N, ld, N, L::RP);
        if (!ok) { set_error("elliptic2d walker kernel: the folded direction update needs TMA-describable operands"); return DGB_ERR_UNSUPPORTED; }
    }
    // one persistent CTA per SM; the work is cut into one cost-weighted piece per warp
    const int grid = sm_count(), nwarps = grid * L::WARPS;
    void*& slot = p.walk_part[DOT ? 1 : 0];
    if (!slot) slot = new WalkPartition();
    WalkPartition& P = *reinterpret_cast<WalkPartition*>(slot);
    int e = build_partition(P, p.Nx, A.Ny, L::UL, L::HL, nwarps, A.fx_lo, A.fx_hi, A.wrapx, A.tma_load, L::SX, DOT, st);
    if (e) return e;
    A.tasks = P.d_tasks; A.tbegin = P.d_tbegin;
    EllipticCoef<N, B> C;
    fill<N, B>(C.rx, p.rightx); fill<N, B>(C.ry, p.righty); fill<N, B>(C.lx, p.leftx); fill<N, B>(C.ly, p.lefty);
    fill<N, 3>(C.jx, p.jumpx); fill<N, 3>(C.jy, p.jumpy);
    if constexpr (FOLD) {
        return wlaunch_go<N, DIRK, DOT, PLAIN, true, false, true>(A, C, mx, ms, mw, my, mp, mpn, grid, st);
    } else {
        if (PLAIN && p.relaxed && A.tma_load && A.tma_store && !A.wrapx)
            return wlaunch_go<N, DIRK, DOT, PLAIN, true, PLAIN>(A, C, mx, ms, mw, my, mp, mpn, grid, st);
        if (A.tma_load && A.tma_store && !A.wrapx) return wlaunch_go<N, DIRK, DOT, PLAIN, true>(A, C, mx, ms, mw, my, mp, mpn, grid, st);
        return wlaunch_go<N, DIRK, DOT, PLAIN, false>(A, C, mx, ms, mw, my, mp, mpn, grid, st);
    }
}

}  // namespace dgb
