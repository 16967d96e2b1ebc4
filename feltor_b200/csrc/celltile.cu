// Field-line interpolation matrices in a CELL-TILED layout (north star item 4: "gather kernel with a dedicated layout").
//
// What the matrices of dg::geo::Fieldaligned look like (inc/geometries/fieldaligned.h:631-657: projection * interpolation):
// the n^2 rows of one TARGET cell all hold the SAME columns -- the nodes of the few SOURCE cells the field lines of that cell
// end in (4 cells for 87 % of the cells of config 4, 2..5 overall: 36.5 entries per row) -- i.e. per target cell the matrix is a
// small DENSE n^2 x U block.  A row-wise (CSR / sliced-ELL) kernel loads one gathered operand per fused multiply-add and is
// bound by the L1 gather path (ncu: 88 % L1, 7 % of the HBM roofline).  Here:
//   * the plan stores, per target cell, one record per column in CSR order: the n^2 coefficients of the cell's rows and the
//     index of the column's operand in the tile's staging buffer; a tile's records are contiguous (one TMA bulk copy);
//   * a CTA owns a strip of 32 target cells and PL consecutive planes; it stages the union of their source cells for the PL
//     planes in shared memory (cp.async, layout [point][plane]: a thread's two planes are one 16-byte word);
//   * a thread owns one target cell and TWO planes: per column it loads one 16-byte operand pair and the record (128-bit
//     shared-memory loads, broadcast to the eight threads of the cell) and issues 2 n^2 FMAs -- 6 loads per 18 FMAs at n = 3
//     instead of 18 per 18; the results leave through a shared-memory transpose so that bphi / g move in full lines;
//   * the matrix is read once per PL planes (it stays in L2), f roughly 2.3 times from L2 (overlapping source regions).
// Summation order per row = column order of the CSR row (ascending), so the results are bitwise those of dgb_csr_spmv_planes / dgb_gather_* and of the reference's OpenMP kernel.
#include "async_copy.cuh"
#include <vector>
#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace dgb {

constexpr int CT_CELLS = 16;     // target cells per tile (a strip along x)
constexpr int CT_THREADS = 128;  // 16 cells x 8 plane pairs; two CTAs per SM overlap staging and arithmetic
constexpr int CT_MAX_SRC = 1024; // source cells a tile may need (staging-slot table in shared memory)

struct CellTilePlan {
    int n = 0, Nx = 0, Ny = 0, nn = 0, num_rows = 0, ntiles = 0, tiles_x = 0;
    int max_src = 0;             // largest number of source cells of a tile
    int max_entries = 0;         // largest number of column entries of a tile (sum over its target cells)
    int PL = 16;                 // planes per CTA (16 if the staging buffer fits, else 8)
    long long nnz = 0;
    int* tile_src_off = nullptr; // [ntiles + 1] -> src_cell
    int* src_cell = nullptr;     // source cell ids (cy * Nx + cx) of a tile, ascending
    int* cell_off = nullptr;     // [ncells + 1] first column entry of a target cell
    double* coef = nullptr;      // [entries][PADN]: n^2 coefficients (row of the cell fastest) + in slot n^2 the operand's staging index
};

__device__ __forceinline__ void ct_cp_async8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void ct_cp_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }

__device__ __forceinline__ unsigned long long ct_time() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// experiment hook (DGB_CT_TRACE=1): phase timestamps of every CTA
#define CT_MARK(k) do { if (trace && threadIdx.x == 0) trace[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 8 + (k)] = ct_time(); } while (0)
struct CtView {
    const int* tile_src_off;
    const int* src_cell;
    const int* cell_off;
    const double* coef;
};

// one matrix applied to the staged planes: acc[k][0..1] += sum_u coef[u][k] * f[loc[u]][2 pp .. 2 pp + 1]
// (the CSR kernels multiply alpha = 1 into every value first, sparsematrix_omp.h:39 -- an exact identity, not repeated here)
// a column entry in shared memory: NN coefficients (row of the cell fastest) and, in the padding slot NN, the 16-bit index
// of the column's operand in the staging buffer (stored as the low bits of that double)
template <int NN, int PADN, int PITCH>
__device__ __forceinline__ void ct_load_column(const double2* cf, const double* fsm, int pp, double (&c)[PADN], double2& fv) {
#pragma unroll
    for (int k = 0; k < PADN / 2; k++) {
        const double2 t = cf[k];
        c[2 * k] = t.x;
        c[2 * k + 1] = t.y;
    }
    const int lc = (int)(__double_as_longlong(c[NN]) & 0xffff);
    fv = *reinterpret_cast<const double2*>(fsm + (size_t)lc * PITCH + 2 * pp);
}
// b .. e: the entries of this thread's target cell inside the tile's staged block `cfs`
template <int NN, int PADN, int PITCH>
__device__ __forceinline__ void ct_apply(const double* cfs, int b, int e, bool active, const double* fsm, int pp, double (&acc)[NN][2]) {
#pragma unroll
    for (int k = 0; k < NN; k++) acc[k][0] = acc[k][1] = 0.;
    if (!active) return;
    const double2* cf = reinterpret_cast<const double2*>(cfs + (size_t)b * PADN);
    int u = b;
    // two columns per trip: their loads (all shared memory: operands conflict-free, coefficients broadcast to the eight
    // threads of the cell) are independent and issue before the first FMA
    for (; u + 1 < e; u += 2, cf += PADN) {
        double c0[PADN], c1[PADN];
        double2 f0, f1;
        ct_load_column<NN, PADN, PITCH>(cf, fsm, pp, c0, f0);
        ct_load_column<NN, PADN, PITCH>(cf + PADN / 2, fsm, pp, c1, f1);
#pragma unroll
        for (int k = 0; k < NN; k++) {
            acc[k][0] = __fma_rn(c0[k], f0.x, acc[k][0]);
            acc[k][1] = __fma_rn(c0[k], f0.y, acc[k][1]);
        }
#pragma unroll
        for (int k = 0; k < NN; k++) {
            acc[k][0] = __fma_rn(c1[k], f1.x, acc[k][0]);
            acc[k][1] = __fma_rn(c1[k], f1.y, acc[k][1]);
        }
    }
    if (u < e) {
        double c0[PADN];
        double2 f0;
        ct_load_column<NN, PADN, PITCH>(cf, fsm, pp, c0, f0);
#pragma unroll
        for (int k = 0; k < NN; k++) {
            acc[k][0] = __fma_rn(c0[k], f0.x, acc[k][0]);
            acc[k][1] = __fma_rn(c0[k], f0.y, acc[k][1]);
        }
    }
}

// stage the source cells of `tile` for planes p0 .. p0 + PL - 1 of x shifted by `shift` (periodic in the plane index).
// A warp takes one source cell at a time: its NN x PL elements, node index fastest (the N nodes of a node row are adjacent
// in memory).  No division in the loop: the offsets of the slots and of the planes come from small shared tables.
template <int NN, int N, int PL, int PITCH, int PADNV>
__device__ __forceinline__ void ct_stage(const CtView& V, int tile, const double* x, int Nx, size_t plane_size, int nplanes, int p0, int shift,
                                         double* fsm, int* sbase, long long* ploff, int ent0, int ent1, double* cfs, unsigned long long* bar) {
    const int sb = V.tile_src_off[tile], S = V.tile_src_off[tile + 1] - sb;
    const int rowlen = Nx * N;
    for (int s = threadIdx.x; s < S; s += CT_THREADS) {
        const int sc = V.src_cell[sb + s];
        const int cy = sc / Nx, cx = sc - cy * Nx;
        sbase[s] = cy * N * rowlen + cx * N;
    }
    if (threadIdx.x < PL) {
        int pl = min(p0 + (int)threadIdx.x, nplanes - 1) + shift;
        pl %= nplanes;
        if (pl < 0) pl += nplanes;
        ploff[threadIdx.x] = (long long)pl * (long long)plane_size;
    }
    __syncthreads();
    // the dense blocks of the tile's target cells are one contiguous range of the plan: ONE bulk copy (TMA engine)
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)(ent1 - ent0) * PADNV * 8u;
        if (bytes) {
            mbar_expect_tx(bar, bytes);
            bulk_load_1d(cfs, V.coef + (size_t)ent0 * PADNV, bytes, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int s = warp; s < S; s += CT_THREADS / 32) {
        const double* base = x + sbase[s];
        double* dst = fsm + (size_t)s * NN * PITCH;
#pragma unroll
        for (int e = lane; e < NN * PL; e += 32) {
            const int j = e / NN, q = e - j * NN;  // NN is a compile-time constant
            ct_cp_async8(dst + q * PITCH + j, base + ploff[j] + (q / N) * rowlen + q % N);
        }
    }
}

// MODE 0: y[pl] = alpha M x[(pl + shift) mod nplanes] + beta y[pl]
// MODE 1: DS::centered for periodic z (ds.h:481-485, 776-786): g = alpha bphi (I+ f[k+1] - I- f[k-1]) / 2 / dphi + beta g
template <int N, int PL, int MODE>
__global__ void __launch_bounds__(CT_THREADS)
celltile_kernel(CtView P, CtView M, int max_src, int max_entries, int Nx, int Ny, int tiles_x, int nplanes, int shift, double alpha, const double* __restrict__ x,
                const double* __restrict__ bphi, double delta, double beta, double* __restrict__ y, unsigned long long* trace) {
    constexpr int NN = N * N, PADN = (NN + 2) & ~1, PITCH = PL + 2;  // pitch of a staged point: PL planes + 2 (bank spread, 16-byte rows)
    extern __shared__ __align__(128) double fsm[];   // [max_src * NN][PITCH] operands | [max_entries][PADN] coefficients | loc
    __shared__ int sbase[CT_MAX_SRC];
    __shared__ long long ploff[PL];
    __shared__ unsigned long long bar;
    double* cfs = fsm + (size_t)max_src * NN * PITCH;
    const int tile = blockIdx.x, p0 = blockIdx.y * PL;
    const int t = threadIdx.x >> 3, pp = threadIdx.x & 7;  // target cell slot, plane pair (PL = 16); PL = 8: pairs 4..7 idle
    const int cyt = tile / tiles_x, cx0 = (tile - cyt * tiles_x) * CT_CELLS, cx = cx0 + t;
    const bool active = cx < Nx && 2 * pp < PL;
    const int c0 = cyt * Nx + cx0, c1 = cyt * Nx + min(cx0 + CT_CELLS, Nx), cell = cyt * Nx + min(cx, Nx - 1);
    const size_t plane_size = (size_t)Nx * Ny * NN;
    const int rowlen = Nx * N;
    CT_MARK(0);
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    double a[NN][2], b[NN][2];
    {
        const int e0 = P.cell_off[c0], e1 = P.cell_off[c1];
        ct_stage<NN, N, PL, PITCH, PADN>(P, tile, x, Nx, plane_size, nplanes, p0, MODE == 1 ? 1 : shift, fsm, sbase, ploff, e0, e1, cfs, &bar);
        ct_cp_async_wait();
        mbar_wait(&bar, 0);
        __syncthreads();
        CT_MARK(1);
        ct_apply<NN, PADN, PITCH>(cfs, P.cell_off[cell] - e0, P.cell_off[cell + 1] - e0, active, fsm, pp, a);
    }
    if (MODE == 1) {
        __syncthreads();
        CT_MARK(2);
        const int e0 = M.cell_off[c0], e1 = M.cell_off[c1];
        ct_stage<NN, N, PL, PITCH, PADN>(M, tile, x, Nx, plane_size, nplanes, p0, -1, fsm, sbase, ploff, e0, e1, cfs, &bar);
        ct_cp_async_wait();
        mbar_wait(&bar, 1);
        __syncthreads();
        CT_MARK(3);
        ct_apply<NN, PADN, PITCH>(cfs, M.cell_off[cell] - e0, M.cell_off[cell + 1] - e0, active, fsm, pp, b);
    }
    // ---- epilogue through shared memory: a thread holds n^2 values of one cell for two planes, global memory wants runs along
    // x.  The tile's results go to out[plane][node row][x], then consecutive threads read bphi / write y consecutively
    // (each node row of the tile is CT_CELLS * N contiguous doubles).
    __syncthreads();
    CT_MARK(4);
    constexpr int ROW = CT_CELLS * N;
    double* out = fsm;
    if (active) {
#pragma unroll
        for (int h = 0; h < 2; h++)
#pragma unroll
            for (int k = 0; k < NN; k++)
                out[((2 * pp + h) * N + k / N) * ROW + t * N + k % N] = MODE == 1 ? __dsub_rn(a[k][h], b[k][h]) : a[k][h];
    }
    __syncthreads();
    const int ncx = min(CT_CELLS, Nx - cx0) * N;  // doubles per node row of this tile
    constexpr int TRIPS = (PL * N * ROW + CT_THREADS - 1) / CT_THREADS;
    // all global loads of the trips first (bphi, and y when beta != 0): their latencies overlap instead of adding up
    double bv[TRIPS], yv[TRIPS];
    size_t gidx[TRIPS];
    bool ok[TRIPS];
#pragma unroll
    for (int tr = 0; tr < TRIPS; tr++) {
        const int i = threadIdx.x + tr * CT_THREADS;
        const int xx = i % ROW, r = i / ROW, ky = r % N, j = r / N;
        const int pl = p0 + j;
        ok[tr] = i < PL * N * ROW && xx < ncx && pl < nplanes;
        gidx[tr] = (size_t)min(pl, nplanes - 1) * plane_size + (size_t)(cyt * N + ky) * rowlen + cx0 * N + min(xx, ncx - 1);
        bv[tr] = (MODE == 1 && ok[tr]) ? bphi[gidx[tr]] : 0.;
        yv[tr] = (beta != 0. && ok[tr]) ? y[gidx[tr]] : 0.;
    }
#pragma unroll
    for (int tr = 0; tr < TRIPS; tr++) {
        if (!ok[tr]) continue;
        const double v0 = out[threadIdx.x + tr * CT_THREADS];
        if (MODE == 1) {
            const double v = __ddiv_rn(__ddiv_rn(__dmul_rn(__dmul_rn(alpha, bv[tr]), v0), 2.), delta);
            y[gidx[tr]] = beta == 0. ? v : __dadd_rn(v, __dmul_rn(beta, yv[tr]));
        } else {
            // the sum was formed with alpha = 1; the reference multiplies alpha into every value first (sparsematrix_omp.h:39,47),
            // which commutes bitwise for alpha = +-1 only -- enforced on the host
            const double tsum = alpha == 1. ? v0 : -v0;
            y[gidx[tr]] = beta == 0. ? tsum : __fma_rn(beta, yv[tr], tsum);
        }
    }
    __syncthreads();
    CT_MARK(5);
}

static CtView view(const CellTilePlan* P) { return CtView{P->tile_src_off, P->src_cell, P->cell_off, P->coef}; }

template <class T>
static int upload(T** dst, const std::vector<T>& v, cudaStream_t st) {
    DGB_CUDA(cudaMalloc(dst, sizeof(T) * std::max<size_t>(v.size(), 1)));
    if (!v.empty()) DGB_CUDA(cudaMemcpyAsync(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice, st));
    return 0;
}

static size_t ct_smem(int max_src, int max_entries, int nn, int PL) {
    const int padn = (nn + 2) & ~1;
    int n = 2;
    while (n * n < nn) n++;
    const size_t stage = (size_t)std::max(max_src, 1) * nn * (PL + 2) * sizeof(double) + (size_t)std::max(max_entries, 1) * padn * sizeof(double) + 16;
    const size_t epilogue = (size_t)PL * n * CT_CELLS * n * sizeof(double);  // the output tile re-uses the staging area
    return std::max(stage, epilogue);
}

template <int N, int MODE>
static int ct_launch(const CellTilePlan* P, const CellTilePlan* M, int PL, int nplanes, int shift, double alpha, const double* x,
                     const double* bphi, double delta, double beta, double* y, cudaStream_t st) {
    const int max_src = std::max(P->max_src, M ? M->max_src : 0), max_entries = std::max(P->max_entries, M ? M->max_entries : 0);
    const size_t smem = ct_smem(max_src, max_entries, P->nn, PL);
    dim3 grid(P->ntiles, (nplanes + PL - 1) / PL);
    const CtView vp = view(P), vm = M ? view(M) : view(P);
    unsigned long long* trace = nullptr;
    const size_t nct = (size_t)grid.x * grid.y;
    static int want_trace = -1;
    if (want_trace < 0) { const char* e = getenv("DGB_CT_TRACE"); want_trace = (e && atoi(e)) ? 1 : 0; }
    if (want_trace) { DGB_CUDA(cudaMalloc(&trace, nct * 8 * sizeof(unsigned long long))); DGB_CUDA(cudaMemset(trace, 0, nct * 64)); }
    if (PL == 16) {
        DGB_CUDA(cudaFuncSetAttribute(celltile_kernel<N, 16, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        celltile_kernel<N, 16, MODE><<<grid, CT_THREADS, smem, st>>>(vp, vm, max_src, max_entries, P->Nx, P->Ny, P->tiles_x, nplanes, shift, alpha, x, bphi, delta, beta, y, trace);
    } else {
        DGB_CUDA(cudaFuncSetAttribute(celltile_kernel<N, 8, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        celltile_kernel<N, 8, MODE><<<grid, CT_THREADS, smem, st>>>(vp, vm, max_src, max_entries, P->Nx, P->Ny, P->tiles_x, nplanes, shift, alpha, x, bphi, delta, beta, y, trace);
    }
    DGB_LAUNCHED();
    if (trace) {
        std::vector<unsigned long long> h(nct * 8);
        DGB_CUDA(cudaMemcpy(h.data(), trace, nct * 64, cudaMemcpyDeviceToHost));
        cudaFree(trace);
        unsigned long long t0 = ~0ull, t1 = 0;
        double ph[5] = {0, 0, 0, 0, 0};
        for (size_t c = 0; c < nct; c++) {
            t0 = std::min(t0, h[c * 8]); t1 = std::max(t1, h[c * 8 + 5]);
            for (int k = 0; k < 5; k++) ph[k] += (double)(h[c * 8 + k + 1] - h[c * 8 + k]);
        }
        fprintf(stderr, "[celltile trace] %zu CTAs, kernel span %.1f us; mean per CTA (us): stage+ %.2f compute+ %.2f stage- %.2f compute- %.2f epilogue %.2f\n",
                nct, (t1 - t0) * 1e-3, ph[0] / nct * 1e-3, ph[1] / nct * 1e-3, ph[2] / nct * 1e-3, ph[3] / nct * 1e-3, ph[4] / nct * 1e-3);
    }
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_celltile_plan_create(dgb_celltile_plan** out, int n, int Nx, int Ny, const int* pos_dev, const int* idx_dev, const double* val_dev,
                             dgb_stream_t s) {
    if (n < 2 || n > 4 || Nx < 1 || Ny < 1) { set_error("dgb_celltile_plan_create: n = %d (2..4), Nx = %d, Ny = %d", n, Nx, Ny); return DGB_ERR_UNSUPPORTED; }
    cudaStream_t st = as_stream(s);
    const int nn = n * n, num_rows = nn * Nx * Ny, rowlen = Nx * n, padn = (nn + 2) & ~1;
    std::vector<int> pos(num_rows + 1);
    DGB_CUDA(cudaMemcpyAsync(pos.data(), pos_dev, sizeof(int) * (num_rows + 1), cudaMemcpyDeviceToHost, st));
    DGB_CUDA(cudaStreamSynchronize(st));
    const long long nnz = pos[num_rows];
    std::vector<int> idx(nnz);
    std::vector<double> val(nnz);
    if (nnz) {
        DGB_CUDA(cudaMemcpyAsync(idx.data(), idx_dev, sizeof(int) * nnz, cudaMemcpyDeviceToHost, st));
        DGB_CUDA(cudaMemcpyAsync(val.data(), val_dev, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st));
        DGB_CUDA(cudaStreamSynchronize(st));
    }
    auto row_of = [&](int cy, int cx, int k) { return (cy * n + k / n) * rowlen + cx * n + k % n; };
    // 1. every target cell: one column list shared by its n^2 rows (else this is not a field-aligned interpolation matrix)
    const int ncells = Nx * Ny;
    std::vector<int> cell_off(ncells + 1, 0);
    for (int c = 0; c < ncells; c++) {
        const int cy = c / Nx, cx = c % Nx, r0 = row_of(cy, cx, 0), len = pos[r0 + 1] - pos[r0];
        for (int k = 1; k < nn; k++) {
            const int r = row_of(cy, cx, k);
            if (pos[r + 1] - pos[r] != len || (len && memcmp(&idx[pos[r]], &idx[pos[r0]], sizeof(int) * len) != 0)) {
                set_error("dgb_celltile_plan_create: the rows of target cell (%d, %d) do not share one column list", cx, cy);
                return DGB_ERR_UNSUPPORTED;
            }
        }
        for (int u = 0; u < len; u++)
            if (idx[pos[r0] + u] < 0 || idx[pos[r0] + u] >= num_rows) { set_error("dgb_celltile_plan_create: column index out of range"); return DGB_ERR_INVALID; }
        cell_off[c + 1] = cell_off[c] + len;
    }
    // 2. tiles = strips of CT_CELLS target cells along x; their source cells (ascending) get staging slots
    CellTilePlan* P = new CellTilePlan();
    P->n = n; P->Nx = Nx; P->Ny = Ny; P->nn = nn; P->num_rows = num_rows; P->nnz = nnz;
    P->tiles_x = (Nx + CT_CELLS - 1) / CT_CELLS;
    P->ntiles = P->tiles_x * Ny;
    std::vector<int> tile_src_off(P->ntiles + 1, 0), src_cell;
    std::vector<double> coef((size_t)cell_off[ncells] * padn, 0.);
    std::vector<int> slot_of(ncells, -1), touched;
    for (int tile = 0; tile < P->ntiles; tile++) {
        const int cy = tile / P->tiles_x, cx0 = (tile % P->tiles_x) * CT_CELLS, cx1 = std::min(Nx, cx0 + CT_CELLS);
        touched.clear();
        for (int cx = cx0; cx < cx1; cx++) {
            const int r0 = row_of(cy, cx, 0);
            for (int e = pos[r0]; e < pos[r0 + 1]; e++) {
                const int col = idx[e], sc = (col / rowlen / n) * Nx + (col % rowlen) / n;
                if (slot_of[sc] < 0) { slot_of[sc] = 0; touched.push_back(sc); }
            }
        }
        std::sort(touched.begin(), touched.end());
        for (size_t k = 0; k < touched.size(); k++) slot_of[touched[k]] = (int)k;
        P->max_src = std::max(P->max_src, (int)touched.size());
        P->max_entries = std::max(P->max_entries, cell_off[cy * Nx + cx1] - cell_off[cy * Nx + cx0]);
        if (touched.size() * nn > 65535 || touched.size() > (size_t)CT_MAX_SRC) { delete P; set_error("dgb_celltile_plan_create: a tile needs %zu source cells", touched.size()); return DGB_ERR_UNSUPPORTED; }
        for (int cx = cx0; cx < cx1; cx++) {
            const int c = cy * Nx + cx, r0 = row_of(cy, cx, 0);
            for (int u = 0; u < pos[r0 + 1] - pos[r0]; u++) {
                const int col = idx[pos[r0] + u], py = col / rowlen, px = col % rowlen;
                const int sc = (py / n) * Nx + px / n, node = (py % n) * n + px % n;
                for (int k = 0; k < nn; k++) coef[(size_t)(cell_off[c] + u) * padn + k] = val[pos[row_of(cy, cx, k)] + u];
                const long long bits = slot_of[sc] * nn + node;  // index of the column's operand in the staging buffer
                memcpy(&coef[(size_t)(cell_off[c] + u) * padn + nn], &bits, sizeof(double));
            }
        }
        src_cell.insert(src_cell.end(), touched.begin(), touched.end());
        tile_src_off[tile + 1] = (int)src_cell.size();
        for (int sc : touched) slot_of[sc] = -1;
    }
    // 3. planes per CTA from the shared-memory budget (two CTAs per SM at PL = 16 when the source regions are compact)
    P->PL = ct_smem(P->max_src, P->max_entries, nn, 16) <= 200 * 1024 ? 16 : 8;
    if (ct_smem(P->max_src, P->max_entries, nn, P->PL) > 200 * 1024) { delete P; set_error("dgb_celltile_plan_create: source regions too large (%d cells per tile)", P->max_src); return DGB_ERR_UNSUPPORTED; }
    int e = 0;
    if (!e) e = upload(&P->tile_src_off, tile_src_off, st);
    if (!e) e = upload(&P->src_cell, src_cell, st);
    if (!e) e = upload(&P->cell_off, cell_off, st);
    if (!e) e = upload(&P->coef, coef, st);
    if (!e && cudaStreamSynchronize(st) != cudaSuccess) e = (int)cudaErrorUnknown;  // the vectors are pageable host memory
    if (e) { dgb_celltile_plan_destroy(reinterpret_cast<dgb_celltile_plan*>(P)); return e; }
    *out = reinterpret_cast<dgb_celltile_plan*>(P);
    return 0;
}
int dgb_celltile_plan_destroy(dgb_celltile_plan* h) {
    CellTilePlan* P = reinterpret_cast<CellTilePlan*>(h);
    if (!P) return 0;
    cudaFree(P->tile_src_off); cudaFree(P->src_cell); cudaFree(P->cell_off); cudaFree(P->coef);
    delete P;
    return 0;
}
int dgb_celltile_plan_info(const dgb_celltile_plan* h, int* ntiles, int* max_source_cells, int* planes_per_cta, long long* nnz) {
    const CellTilePlan* P = reinterpret_cast<const CellTilePlan*>(h);
    if (!P) { set_error("dgb_celltile_plan_info: NULL plan"); return DGB_ERR_INVALID; }
    if (ntiles) *ntiles = P->ntiles;
    if (max_source_cells) *max_source_cells = P->max_src;
    if (planes_per_cta) *planes_per_cta = P->PL;
    if (nnz) *nnz = P->nnz;
    return 0;
}
int dgb_celltile_spmv_planes(const dgb_celltile_plan* h, double alpha, const double* x, double beta, double* y, int nplanes, int shift,
                             dgb_stream_t s) {
    const CellTilePlan* P = reinterpret_cast<const CellTilePlan*>(h);
    if (!P) { set_error("dgb_celltile_spmv_planes: NULL plan"); return DGB_ERR_INVALID; }
    if (nplanes <= 0) return 0;
    if (x == y) { set_error("dgb_celltile_spmv_planes: x must not alias y"); return DGB_ERR_INVALID; }
    if ((alpha != 1. && alpha != -1.) || beta == 1.) {
        set_error("dgb_celltile_spmv_planes: bit-exact for alpha = +-1 and beta != 1 only (use dgb_gather_spmv_planes otherwise)");
        return DGB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = as_stream(s);
    switch (P->n) {
        case 2: return ct_launch<2, 0>(P, nullptr, P->PL, nplanes, shift, alpha, x, nullptr, 0., beta, y, st);
        case 3: return ct_launch<3, 0>(P, nullptr, P->PL, nplanes, shift, alpha, x, nullptr, 0., beta, y, st);
        default: return ct_launch<4, 0>(P, nullptr, P->PL, nplanes, shift, alpha, x, nullptr, 0., beta, y, st);
    }
}
int dgb_celltile_ds_centered(const dgb_celltile_plan* plus, const dgb_celltile_plan* minus, int nplanes, double alpha, const double* f,
                             const double* bphi, double delta_phi, double beta, double* g, dgb_stream_t s) {
    const CellTilePlan* P = reinterpret_cast<const CellTilePlan*>(plus);
    const CellTilePlan* M = reinterpret_cast<const CellTilePlan*>(minus);
    if (!P || !M || P->n != M->n || P->Nx != M->Nx || P->Ny != M->Ny) { set_error("dgb_celltile_ds_centered: plans of different grids"); return DGB_ERR_INVALID; }
    if (nplanes <= 0) return 0;
    if (f == g) { set_error("dgb_celltile_ds_centered: f must not alias g"); return DGB_ERR_INVALID; }
    const int PL = std::min(P->PL, M->PL);
    cudaStream_t st = as_stream(s);
    switch (P->n) {
        case 2: return ct_launch<2, 1>(P, M, PL, nplanes, 0, alpha, f, bphi, delta_phi, beta, g, st);
        case 3: return ct_launch<3, 1>(P, M, PL, nplanes, 0, alpha, f, bphi, delta_phi, beta, g, st);
        default: return ct_launch<4, 1>(P, M, PL, nplanes, 0, alpha, f, bphi, delta_phi, beta, g, st);
    }
}
}
