// Fused Elliptic2d apply (see elliptic.cu for the overview).
// Persistent kernel: min(#tiles, 2 x #SM) CTAs of 256 threads walk over tiles of TX x TY cells.  Per tile
//   phase 0  x (halo H cells) and sigma (halo 1 cell) arrive in shared memory by TMA (cp.async.bulk.tensor.2d,
//            out-of-bounds elements are zero-filled by the hardware = non-periodic boundary); tiles that touch a
//            periodic seam, or operands TMA cannot describe, use an LDGSTS (cp.async) loader instead;
//   phase 1  fluxes tx = sigma Rx x, ty = sigma Ry x on the tile plus the ring the adjoint derivative reaches;
//   phase 2  one thread per cell replays the reference's rounding sequence for its n x n outputs in registers;
//   phase 3  coalesced epilogue y = alpha t / vol + beta y, optionally fused with the exact dot(x, w, y).
//
// Rounding sequence per output element (inc/dg/elliptic.h:431-458 on top of
// inc/dg/backend/sparseblockmat_omp_kernels.h:36-50 and inc/dg/topology/multiply.h:18-32), with
// blk(M,d) = fma-chain over q of M's block in slot d:
//   gx = 0; for d: gx = fma(1, blk(Rx,d), gx)            tx = fma(sigma, gx, gx*0)      (identity chi tensor)
//   gy likewise with Ry                                   ty = fma(sigma, gy, gy*0)
//   t  = 0; for d: t = fma(1, blk(Ly,d)[ty], t);  t = t*(-1);  for d: t = fma(-1, blk(Lx,d)[tx], t)
//   for d: t = fma(jfactor, blk(Jx,d)[x], t);  for d: t = fma(jfactor, blk(Jy,d)[x], t)
//   y = fma(alpha, t/vol, beta*y)
#include "elliptic_dev.cuh"

namespace dgb {

#ifndef DGB_TY
#define DGB_TY 8
#endif
constexpr int TX = 32, TY = DGB_TY, FUSED_THREADS = TX * TY;  // one thread per cell of the tile
constexpr int FUSED_MIN_CTAS = 512 / FUSED_THREADS;           // 16 warps per SM

struct FusedArgs {
    MatView rx, ry, lx, ly, jx, jy;
    int Nx, Ny, wrapx, wrapy;
    int fx_lo, fx_hi, fy_lo, fy_hi;  // cells [lo, hi) that are interior rows of all three x- resp. y-matrices
    int ntx, ntiles, use_tma;
    // Elliptic3d in compute-in-2d mode: nplanes consecutive planes of plane_stride doubles in x, y, sigma (and the
    // Helmholtz chi); tile index = plane * ntiles + tile of the plane.  nplanes == 1: the 2-d operator
    int nplanes;
    size_t plane_stride;
    // slab mode (domain decomposition in y): the operands x and sigma carry `ghost` cell rows on either side, row 0
    // of the slab is global cell row `yoff` of `Nyg`; the periodic wrap in y is done by the halo exchange
    int slab, ghost, yoff, Nyg, pery;
    const double* sigma;
    const double* vol;
    const double* x;
    double* y;
    double alpha, beta, jfactor;
    int helm;                 // GeneralHelmholtz epilogue: y = chi x - helm_alpha y (helmholtz.h:74-80)
    double helm_alpha;
    const double* helm_chi;
    // optional fused dot(x, w, y) of the PCG step (pcg.h:165-166): products round(round(x*w)*y)
    const double* dot_w;
    sa::DotSlot slot;
    PcgState* pcg;
    P2pView p2p;
    unsigned long long epoch;
};

// out[k] = fma(a, sum_q C[d][k][q] * s[(off_d*N + q)*stride], out[k]) for the slots d of block-row `cell` of M
template <int N, int BPL>
__device__ __forceinline__ void apply_row(const MatView& M, const double (&C)[BPL][N][N], int cell, const double* s,
                                          int stride, double a, double (&out)[N]) {
    if (cell >= M.i_lo && cell < M.i_hi) {
#pragma unroll
        for (int d = 0; d < BPL; d++) {
            const double* p = s + (M.off[d] * N) * stride;
            double xv[N];
#pragma unroll
            for (int q = 0; q < N; q++) xv[q] = p[q * stride];
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) t = __fma_rn(C[d][k][q], xv[q], t);
                out[k] = __fma_rn(a, t, out[k]);
            }
        }
    } else {
#pragma unroll 1
        for (int d = 0; d < BPL; d++) {
            const int col = M.cols[cell * BPL + d];
            if (col < 0) continue;
            // boundary rows keep their own slot -> neighbour assignment (dx.h:85-97): take the offset from the
            // column index, undoing the periodic wrap
            int o = col - cell;
            if (o > 1) o -= M.num;
            else if (o < -1) o += M.num;
            const double* blk = M.data + (size_t)M.didx[cell * BPL + d] * N * N;
            const double* p = s + (o * N) * stride;
            double xv[N];
#pragma unroll
            for (int q = 0; q < N; q++) xv[q] = p[q * stride];
#pragma unroll
            for (int k = 0; k < N; k++) {
                double t = 0.;
#pragma unroll
                for (int q = 0; q < N; q++) t = __fma_rn(__ldg(blk + k * N + q), xv[q], t);
                out[k] = __fma_rn(a, t, out[k]);
            }
        }
    }
}

template <int N, int KIND, int STRIDE>
__device__ __forceinline__ void apply_fast(const double (&C)[Offs<KIND>::BPL][N][N], const double* s, double a, double (&out)[N]) {
#pragma unroll
    for (int d = 0; d < Offs<KIND>::BPL; d++) {
        double xv[N];
#pragma unroll
        for (int q = 0; q < N; q++) xv[q] = s[(Offs<KIND>::at(d) * N + q) * STRIDE];
#pragma unroll
        for (int k = 0; k < N; k++) {
            double t = 0.;
#pragma unroll
            for (int q = 0; q < N; q++) t = __fma_rn(C[d][k][q], xv[q], t);
            out[k] = __fma_rn(a, t, out[k]);
        }
    }
}

constexpr int NOROW = -(1 << 30);
// local cell row c (may lie in the halo) -> row used for ADDRESSING the operand, NOROW if no data exists
__device__ __forceinline__ int yaddr(int c, const FusedArgs& A) {
    if (!A.slab) { int g = gcell(c, A.Ny, A.wrapy); return g < 0 ? NOROW : g; }
    if (c < -A.ghost || c >= A.Ny + A.ghost) return NOROW;
    const int g = c + A.yoff;
    if (!A.pery && (g < 0 || g >= A.Nyg)) return NOROW;
    return c;
}
// local cell row c -> block-row index of the (global) y-matrices, NOROW if the row does not exist
__device__ __forceinline__ int ymat(int c, const FusedArgs& A) {
    if (!A.slab) { int g = gcell(c, A.Ny, A.wrapy); return g < 0 ? NOROW : g; }
    int g = c + A.yoff;
    if (g < 0 || g >= A.Nyg) {
        if (!A.pery) return NOROW;
        g = g < 0 ? g + A.Nyg : g - A.Nyg;
    }
    return g;
}

// LDGSTS loader of a (rows x cols) tile whose first cell is (cy, cx): one smem row at a time, columns by thread
template <int N, int ROWS, int COLS, int PITCH>
__device__ __forceinline__ void load_tile_ldgsts(double* dst, const double* src, int cy, int cx, const FusedArgs& A, int tid) {
    const int LDG = A.Nx * N;
    for (int c = tid % 128; c < COLS; c += 128) {
        const int gx = gcell(cx + c / N, A.Nx, A.wrapx);
        const int gcol = gx * N + c % N;
        for (int r = tid / 128; r < ROWS; r += FUSED_THREADS / 128) {
            const int gy = yaddr(cy + r / N, A);
            const bool ok = gx >= 0 && gy != NOROW;
            cp_async8(dst + r * PITCH + c, ok ? src + ((long long)(gy * N + r % N) * LDG + gcol) : src, ok);
        }
    }
}

template <int N, int B>
struct Tile {
    static constexpr int H = (B == 2) ? 1 : 2;                          // halo of the x tile in cells
    static constexpr int XR = (TY + 2 * H) * N, XC = (TX + 2 * H) * N;  // x tile
    static constexpr int SR = (TY + 2) * N, SC = (TX + 2) * N;          // sigma tile (one-cell ring)
    // TMA needs a 16-byte aligned start in the contiguous dimension: the first tile column (cx0 - halo)*N has the
    // parity of halo*N for every tile (TX*N is even), so odd cases load one extra column on each side
    static constexpr int XSH = (H * N) & 1, SSH = N & 1;
    static constexpr int XP = XC + 2 * XSH, SP = SC + 2 * SSH;          // row pitches of the x / sigma tiles
    static constexpr int TXR = TY * N, TXC = (TX + 2) * N;              // tx: tile rows, one-cell ring in x
    static constexpr int TYR = (TY + 2) * N, TYC = TX * N;              // ty: one-cell ring in y, tile columns
    static constexpr int OR = TY * N, OC = TX * N;                      // output staging tile (aliases tx)
    static constexpr size_t pad(size_t doubles) { return (doubles + 15) / 16 * 16; }  // keep every buffer 128-B aligned
    static constexpr size_t XS = 0, SS = pad(XR * XP), TXS = SS + pad(SR * SP), TYS = TXS + pad(TXR * TXC),
                            END = TYS + pad(TYR * TYC);
    static constexpr size_t BYTES = END * sizeof(double) + 16;  // + mbarrier
};

// phases 1-3 for one staged tile.  FAST = every cell of the tile and of its ring is an interior row of all six
// matrices and the tile lies completely inside the domain: no per-item checks, stencil offsets are immediates.
template <int N, int DIRK, bool DOT, bool FAST>
__device__ __forceinline__ void compute_tile(const FusedArgs& A, const EllipticCoef<N, Offs<DIRK>::BPL>& C, double* xs, double* ss,
                                             double* txs, double* tys, int cx0, int cy0, int tid, sa::Fpe& fpe, int& bad,
                                             long long* dsm, size_t poff) {
    constexpr int B = Offs<DIRK>::BPL;
    constexpr int RK = DIRK, LK = DIRK == 0 ? 1 : (DIRK == 1 ? 0 : 2);  // stencil kinds of the right / left derivatives
    using TL = Tile<N, B>;
    constexpr int H = TL::H, XC = TL::XP, SC = TL::SP, TXC = TL::TXC, TYC = TL::TYC, OC = TL::OC;
    const int LDG = A.Nx * N;
    // flux cells the adjoint derivative reaches (relative to the tile): [LO, TX-1+HI]
    constexpr int LO = Offs<LK>::first, HI = Offs<LK>::first + Offs<LK>::BPL - 1;

    // ---- phase 1a: tx = sigma * (Rx x); item = (row r, cell c), c = 0 is the cell left of the tile
    {
        int r = tid / (TX + 2), c = tid - r * (TX + 2);
        constexpr int DR = FUSED_THREADS / (TX + 2), DC = FUSED_THREADS - DR * (TX + 2);
        for (; r < TL::TXR; r += DR, c += DC) {
            if (c >= TX + 2) { c -= TX + 2; if (++r >= TL::TXR) break; }
            double g[N];
#pragma unroll
            for (int k = 0; k < N; k++) g[k] = 0.;
            if (c - 1 >= LO && c - 1 <= TX - 1 + HI) {
                const double* sx = xs + (H * N + r) * XC + (H - 1 + c) * N;
                bool on = true;
                if (FAST) apply_fast<N, RK, 1>(C.rx, sx, 1., g);
                else {
                    const int gx = gcell(cx0 - 1 + c, A.Nx, A.wrapx);
                    on = gx >= 0 && cy0 + r / N < A.Ny;
                    if (on) apply_row<N, B>(A.rx, C.rx, gx, sx, 1, 1., g);  // rows of the slab itself always exist
                }
                if (on) {
                    const double* sg = ss + (N + r) * SC + c * N;
#pragma unroll
                    for (int k = 0; k < N; k++) g[k] = __fma_rn(sg[k], g[k], __dmul_rn(g[k], 0.));
                }
            }
#pragma unroll
            for (int k = 0; k < N; k++) txs[r * TXC + c * N + k] = g[k];
        }
    }
    // ---- phase 1b: ty = sigma * (Ry x); item = (cell r, column c), r = 0 is the cell below the tile
    {
        int r = tid / TYC, c = tid - r * TYC;
        constexpr int DR = FUSED_THREADS / TYC, DC = FUSED_THREADS - DR * TYC;
        for (; r < TY + 2; r += DR, c += DC) {
            if (c >= TYC) { c -= TYC; if (++r >= TY + 2) break; }
            double g[N];
#pragma unroll
            for (int k = 0; k < N; k++) g[k] = 0.;
            if (r - 1 >= LO && r - 1 <= TY - 1 + HI) {
                const double* sx = xs + ((H - 1 + r) * N) * XC + H * N + c;
                bool on = true;
                if (FAST) apply_fast<N, RK, XC>(C.ry, sx, 1., g);
                else {
                    const int gy = ymat(cy0 - 1 + r, A);
                    on = gy != NOROW && cx0 + c / N < A.Nx && cy0 - 1 + r < A.Ny + (A.slab ? A.ghost : 1);
                    if (on) apply_row<N, B>(A.ry, C.ry, gy, sx, XC, 1., g);
                }
                if (on) {
                    const double* sg = ss + (r * N) * SC + N + c;
#pragma unroll
                    for (int k = 0; k < N; k++) g[k] = __fma_rn(sg[k * SC], g[k], __dmul_rn(g[k], 0.));
                }
            }
#pragma unroll
            for (int k = 0; k < N; k++) tys[(r * N + k) * TYC + c] = g[k];
        }
    }
    __syncthreads();

    // ---- phase 2: one thread per cell
    const int cx = tid % TX, cy = tid / TX;
    const int ix = cx0 + cx, iy = cy0 + cy;
    double acc[N][N];  // [ky][kx]
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
        for (int b = 0; b < N; b++) acc[a][b] = 0.;
    if (FAST || (ix < A.Nx && iy < A.Ny)) {
        // Ly ty (alpha = 1, beta = 0)
#pragma unroll
        for (int kx = 0; kx < N; kx++) {
            double col[N];
#pragma unroll
            for (int k = 0; k < N; k++) col[k] = 0.;
            const double* sp = tys + ((cy + 1) * N) * TYC + cx * N + kx;
            if (FAST) apply_fast<N, LK, TYC>(C.ly, sp, 1., col);
            else apply_row<N, B>(A.ly, C.ly, iy + A.yoff, sp, TYC, 1., col);
#pragma unroll
            for (int k = 0; k < N; k++) acc[k][kx] = col[k];
        }
        // - Lx tx - t   (alpha = -1, beta = -1)
#pragma unroll
        for (int ky = 0; ky < N; ky++) {
            double row[N];
#pragma unroll
            for (int k = 0; k < N; k++) row[k] = __dmul_rn(acc[ky][k], -1.);
            const double* sp = txs + (cy * N + ky) * TXC + (cx + 1) * N;
            if (FAST) apply_fast<N, LK, 1>(C.lx, sp, -1., row);
            else apply_row<N, B>(A.lx, C.lx, ix, sp, 1, -1., row);
#pragma unroll
            for (int k = 0; k < N; k++) acc[ky][k] = row[k];
        }
        if (A.jfactor != 0.) {
#pragma unroll
            for (int ky = 0; ky < N; ky++) {
                double row[N];
#pragma unroll
                for (int k = 0; k < N; k++) row[k] = acc[ky][k];
                const double* sp = xs + ((cy + H) * N + ky) * XC + (cx + H) * N;
                if (FAST) apply_fast<N, 2, 1>(C.jx, sp, A.jfactor, row);
                else apply_row<N, 3>(A.jx, C.jx, ix, sp, 1, A.jfactor, row);
#pragma unroll
                for (int k = 0; k < N; k++) acc[ky][k] = row[k];
            }
#pragma unroll
            for (int kx = 0; kx < N; kx++) {
                double col[N];
#pragma unroll
                for (int k = 0; k < N; k++) col[k] = acc[k][kx];
                const double* sp = xs + ((cy + H) * N) * XC + (cx + H) * N + kx;
                if (FAST) apply_fast<N, 2, XC>(C.jy, sp, A.jfactor, col);
                else apply_row<N, 3>(A.jy, C.jy, iy + A.yoff, sp, XC, A.jfactor, col);
#pragma unroll
                for (int k = 0; k < N; k++) acc[k][kx] = col[k];
            }
        }
    }
    __syncthreads();  // every read of txs is done: reuse it as the output staging tile (OR x OC)
    double* outs = txs;
#pragma unroll
    for (int ky = 0; ky < N; ky++)
#pragma unroll
        for (int kx = 0; kx < N; kx++) outs[(cy * N + ky) * OC + cx * N + kx] = acc[ky][kx];
    __syncthreads();

    // ---- phase 3: coalesced epilogue  y = fma(alpha, t/vol, beta*y).  Warp w owns rows w*N .. w*N+N-1 of the tile,
    //      its lanes the columns lane + 32 j: every access of a warp is one contiguous 256-byte segment.
    const int warp = tid >> 5, lane = tid & 31;
    const size_t gbase = (size_t)(cy0 * N + warp * N) * LDG + cx0 * N + lane;
    const double* so = outs + (warp * N) * OC + lane;
    const double* sxr = xs + (H * N + warp * N) * XC + H * N + lane;
    double yin[N][N], vin[N][N], win[N][N];
    bool okv[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            const size_t g = gbase + (size_t)i * LDG + 32 * j;
            okv[i][j] = FAST || ((cy0 + warp) < A.Ny && (cx0 + (lane + 32 * j) / N) < A.Nx);
            yin[i][j] = 0.; vin[i][j] = 1.; win[i][j] = 0.;
            if (okv[i][j]) {
                if (A.beta != 0.) yin[i][j] = A.y[poff + g];
                if (A.vol) vin[i][j] = __ldg(A.vol + g);
                if (DOT) win[i][j] = __ldg(A.dot_w + g);
            }
        }
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
            if (!okv[i][j]) continue;
            double t = so[i * OC + 32 * j];
            if (A.vol) t = __ddiv_rn(t, vin[i][j]);
            const double b = A.beta == 0. ? 0. : __dmul_rn(yin[i][j], A.beta);
            double v = __fma_rn(A.alpha, t, b);
            if (A.helm) {  // pointwiseDot(1., chi, x, -helm_alpha, y): y *= -helm_alpha; y = fma(1*chi, x, y)
                const double c = A.helm_chi ? __ldg(A.helm_chi + poff + gbase + (size_t)i * LDG + 32 * j) : 1.;
                v = __fma_rn(__dmul_rn(1., c), sxr[i * XC + 32 * j], __dmul_rn(v, -A.helm_alpha));
            }
            A.y[poff + gbase + (size_t)i * LDG + 32 * j] = v;
            if (DOT) {
                double pr = __dmul_rn(__dmul_rn(sxr[i * XC + 32 * j], win[i][j]), v);
                if (!isfinite(pr)) { bad = 1; pr = 0.; }
                fpe.add(pr, dsm);
            }
        }
    __syncthreads();  // the tile buffers are free for the next TMA / LDGSTS round
}

template <int N, int DIRK, bool DOT>
__global__ void __launch_bounds__(FUSED_THREADS, FUSED_MIN_CTAS)
elliptic2d_fused_kernel(const __grid_constant__ FusedArgs A, const __grid_constant__ EllipticCoef<N, Offs<DIRK>::BPL> C,
                        const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_s) {
    constexpr int B = Offs<DIRK>::BPL;
    using TL = Tile<N, B>;
    constexpr int H = TL::H, XC = TL::XP, SC = TL::SP;  // XC, SC: row pitches
    constexpr int NW = FUSED_THREADS / 32;
    __shared__ long long dsm[DOT ? sa::BINS : 1];  // one accumulator per block
    extern __shared__ __align__(128) double smem[];
    double* xs = smem + TL::XS + TL::XSH;  // element (r, c) of the x tile is xs[r * XC + c]
    double* ss = smem + TL::SS + TL::SSH;
    double* txs = smem + TL::TXS;
    double* tys = smem + TL::TYS;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + TL::END);
    const int tid = threadIdx.x;
    sa::Fpe fpe;
    int bad = 0;
    if (DOT) {
        sa::block_init<1>(dsm);
        fpe.clear();
    }
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    pdl_wait();  // programmatic dependent launch (common.cuh): nothing of the predecessor is touched above this line
    if (DOT) {
        if (A.pcg->done) return;  // solver already converged: the remaining launches of the batch are no-ops
    }
    unsigned phase = 0;
    // rows/columns of cells that are interior rows of every matrix (host-computed intersection)
    for (int gtile = blockIdx.x; gtile < A.ntiles * A.nplanes; gtile += gridDim.x) {
        int plane = 0, tile = gtile;
        if (A.nplanes > 1) { plane = gtile / A.ntiles; tile = gtile - plane * A.ntiles; }
        const size_t poff = (size_t)plane * A.plane_stride;
        const int tyi = tile / A.ntx, txi = tile - tyi * A.ntx;
        const int cx0 = txi * TX, cy0 = tyi * TY;
        // ---- phase 0  (stacked planes: the rows beyond a plane belong to its neighbours, so tiles at the y boundary
        //      always take the boundary-aware loader)
        const bool seam = (A.wrapx && (cx0 - H < 0 || cx0 + TX + H > A.Nx)) ||
                          (!A.slab && (A.wrapy || A.nplanes > 1) && (cy0 - H < 0 || cy0 + TY + H > A.Ny));
        if (A.use_tma && !seam) {
            if (tid == 0) {
                mbar_expect_tx(bar, (unsigned)((TL::XR * XC + TL::SR * SC) * sizeof(double)));
                // in slab mode the maps start at the first ghost row
                tma_load_2d(smem + TL::XS, &map_x, bar, (cx0 - H) * N - TL::XSH, (cy0 - H + A.ghost + plane * A.Ny) * N);
                tma_load_2d(smem + TL::SS, &map_s, bar, (cx0 - 1) * N - TL::SSH, (cy0 - 1 + A.ghost + plane * A.Ny) * N);
                mbar_wait(bar, phase);  // one thread polls, the CTA sleeps on the barrier below
            }
            phase ^= 1;
            __syncthreads();
        } else {
            load_tile_ldgsts<N, TL::XR, TL::XC, XC>(xs, A.x + poff, cy0 - H, cx0 - H, A, tid);
            load_tile_ldgsts<N, TL::SR, TL::SC, SC>(ss, A.sigma + poff, cy0 - 1, cx0 - 1, A, tid);
            cp_async_wait_all();
            __syncthreads();
        }
        const bool fast = cx0 - 1 >= A.fx_lo && cx0 + TX + 1 <= A.fx_hi && cy0 + A.yoff - 1 >= A.fy_lo &&
                          cy0 + A.yoff + TY + 1 <= A.fy_hi && cy0 + TY <= A.Ny;
        if (fast) compute_tile<N, DIRK, DOT, true>(A, C, xs, ss, txs, tys, cx0, cy0, tid, fpe, bad, dsm, poff);
        else compute_tile<N, DIRK, DOT, false>(A, C, xs, ss, txs, tys, cx0, cy0, tid, fpe, bad, dsm, poff);
    }
    pdl_trigger();  // the tiles are done: the dependent kernel's launch and prologue may overlap the exact-dot tail
    if (DOT) {
        fpe.flush_warp(dsm);
        fused_dot_finish(sa::block_finish<1>(dsm, bad, A.slot, 0), A.pcg, A.slot.result, A.p2p, A.epoch);
    }
}

template <int N, int DIRK, bool DOT>
static int launch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                  const FusedDot* fd, int nplanes = 1) {
    constexpr int B = Offs<DIRK>::BPL;
    using TL = Tile<N, B>;
    static bool configured = false;
    static int no_tma = -1;
    if (!configured) {
        DGB_CUDA(cudaFuncSetAttribute(elliptic2d_fused_kernel<N, DIRK, DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL::BYTES));
        configured = true;
    }
    if (no_tma < 0) { const char* e = getenv("DGB_NO_TMA"); no_tma = (e && atoi(e)) ? 1 : 0; }
    FusedArgs A;
    A.rx = view(p.rightx); A.ry = view(p.righty); A.lx = view(p.leftx); A.ly = view(p.lefty);
    A.jx = view(p.jumpx); A.jy = view(p.jumpy);
    A.Nx = p.Nx; A.Ny = p.slab ? p.slab_rows : p.Ny; A.wrapx = p.wrapx; A.wrapy = p.slab ? 0 : p.wrapy;
    A.slab = p.slab; A.ghost = p.slab ? p.slab_ghost : 0; A.yoff = p.slab ? p.slab_yoff : 0; A.Nyg = p.Ny; A.pery = p.wrapy;
    if (p.slab && p.slab_ghost < TL::H) {
        set_error("elliptic2d slab: %d ghost cell rows given, the stencil needs %d", p.slab_ghost, TL::H);
        return DGB_ERR_INVALID;
    }
    A.fx_lo = std::max({p.rightx.i_lo, p.leftx.i_lo, p.jumpx.i_lo});
    A.fx_hi = std::min({p.rightx.i_hi, p.leftx.i_hi, p.jumpx.i_hi});
    A.fy_lo = std::max({p.righty.i_lo, p.lefty.i_lo, p.jumpy.i_lo});
    A.fy_hi = std::min({p.righty.i_hi, p.lefty.i_hi, p.jumpy.i_hi});
    A.ntx = (p.Nx + TX - 1) / TX;
    A.ntiles = A.ntx * ((A.Ny + TY - 1) / TY);
    A.nplanes = nplanes; A.plane_stride = (size_t)p.size;
    if (nplanes != 1 && (p.slab || DOT || nplanes < 1)) { set_error("elliptic2d fused kernel: planes need a plain 2-d plan"); return DGB_ERR_UNSUPPORTED; }
    A.sigma = p.sigma; A.vol = p.vol; A.x = x; A.y = y;
    A.alpha = alpha; A.beta = beta; A.jfactor = p.jfactor;
    A.helm = p.helm ? 1 : 0; A.helm_alpha = p.helm_alpha; A.helm_chi = p.helm_chi;
    A.dot_w = nullptr; A.pcg = nullptr; A.slot = sa::DotSlot{nullptr, nullptr, nullptr, nullptr};
    A.p2p = P2pView{}; A.p2p.enabled = 0; A.epoch = 0;
    if (DOT) { A.dot_w = fd->w; A.slot = fd->slot; A.pcg = fd->pcg; A.p2p = fd->p2p; A.epoch = fd->epoch; }
    CUtensorMap mx, ms;
    memset(&mx, 0, sizeof(mx));
    memset(&ms, 0, sizeof(ms));
    const long long gh = (long long)A.ghost * N * p.Nx * N;  // doubles in the ghost rows below the slab
    A.use_tma = !no_tma && make_map(&mx, x - gh, (A.Ny * nplanes + 2 * A.ghost) * N, p.Nx * N, TL::XR, TL::XP) &&
                make_map(&ms, p.sigma - gh, (A.Ny * nplanes + 2 * A.ghost) * N, p.Nx * N, TL::SR, TL::SP);
    EllipticCoef<N, B> C;
    fill<N, B>(C.rx, p.rightx); fill<N, B>(C.ry, p.righty); fill<N, B>(C.lx, p.leftx); fill<N, B>(C.ly, p.lefty);
    fill<N, 3>(C.jx, p.jumpx); fill<N, 3>(C.jy, p.jumpy);
    int per_sm = std::max(1, std::min(FUSED_MIN_CTAS, (int)(220 * 1024 / (TL::BYTES + 1024))));
    int grid = (int)std::min<long long>((long long)A.ntiles * nplanes, (long long)per_sm * sm_count());
    if (DOT && fd->pdl)
        DGB_CUDA(launch_pdl(elliptic2d_fused_kernel<N, DIRK, DOT>, dim3(grid), dim3(FUSED_THREADS), TL::BYTES, st, A, C, mx, ms));
    else
        elliptic2d_fused_kernel<N, DIRK, DOT><<<grid, FUSED_THREADS, TL::BYTES, st>>>(A, C, mx, ms);
    DGB_LAUNCHED();
    return 0;
}

template <bool DOT>
static int dispatch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st,
                    const FusedDot* fd, int nplanes = 1) {
    switch (p.n * 10 + p.dirk) {
        case 20: return launch<2, 0, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 21: return launch<2, 1, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 22: return launch<2, 2, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 30: return launch<3, 0, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 31: return launch<3, 1, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 32: return launch<3, 2, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 40: return launch<4, 0, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 41: return launch<4, 1, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
        case 42: return launch<4, 2, DOT>(p, alpha, x, beta, y, st, fd, nplanes);
    }
    set_error("elliptic2d fused kernel: unsupported n=%d direction kind=%d", p.n, p.dirk);
    return DGB_ERR_UNSUPPORTED;
}
int elliptic2d_fused_launch(Elliptic2dPlan& p, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    if (elliptic2d_walker_supported(p)) return elliptic2d_walker_launch(p, alpha, x, beta, y, st, nullptr);
    return dispatch<false>(p, alpha, x, beta, y, st, nullptr);
}
// the 2-d operator on nplanes stacked planes in ONE launch of the tile kernel (p.sigma points to the 3-d sigma)
int elliptic2d_fused_launch_planes(Elliptic2dPlan& p, int nplanes, double alpha, const double* x, double beta, double* y, cudaStream_t st) {
    return dispatch<false>(p, alpha, x, beta, y, st, nullptr, nplanes);
}
// y = A x fused with dot(x, w, y) and the PCG alpha update (pcg.h:165-166)
int elliptic2d_fused_launch_dot(Elliptic2dPlan& p, const double* x, double* y, cudaStream_t st, const FusedDot& fd) {
    if (elliptic2d_walker_supported(p, true)) return elliptic2d_walker_launch(p, 1., x, 0., y, st, &fd);
    return dispatch<true>(p, 1., x, 0., y, st, &fd);
}

}  // namespace dgb
