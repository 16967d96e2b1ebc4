// Sparse matrix - matrix product A = B C on the device, bit-identical to the reference's HOST kernel
// dg::detail::spgemm_cpu_kernel (inc/dg/backend/sparsematrix_cpu.h:19-95) -- the reference has no device version; its customer on
// the path is dg::geo::Fieldaligned, which multiplies the fine-grid projection with the field-line interpolation three times in
// its constructor (inc/geometries/fieldaligned.h:645-658).  Measured on B200 at that size (n = 3, 96 x 96 cells, mx = my = 10:
// 74.6 M entries per operand, 672 M candidate products, profiles/spgemm_r02.json): 22 ms with device-resident operands, 221 ms
// through the host-array entry points (copies included), 7.2 s for the reference's serial host kernel; bitwise equal.
//
// Reference semantics, kept exactly:
//   * the columns of a row of A are the distinct columns reached through the row of B, SORTED ascending (inputs may be unsorted
//     and may hold duplicates; explicit zeros are kept);
//   * A_ij = sum over the candidates (pB ascending, then pC ascending) of B_val[pB] * C_val[pC], accumulated into a workspace
//     that starts at 0 as  w = fma(b, c, w)  (gcc -mfma contracts `workspace[j] += B_val[pB] * C_val[pC]`; pinned on the live
//     reference in tests/test_spgemm.py) -- so the ORDER of the candidates of one column matters and is preserved here.
// Design: one WARP owns a row.  It expands the row's candidates load-balanced (32 entries of B at a time, a warp scan of the
// lengths of their C rows, every lane takes one candidate of the flattened list), keeps the row's columns in a warp-private
// open-addressing hash table in shared memory with the workspace value beside the key, and applies candidates that meet in the
// same column in lane order (match_any groups, one round per rank) -- the candidate order of the serial loop.  The row is then
// compacted, sorted by column (bitonic, shared memory) and written.  Pass 1 only counts the distinct columns; the row offsets are
// scanned on the host (setup path).  Rows with more distinct columns than the table of the fast kernel holds (512) are redone by
// a one-warp-per-CTA variant with a 4096-column table; beyond that the call reports DGB_ERR_UNSUPPORTED.
#include "common.cuh"
#include <vector>

namespace dgb {

constexpr int SPG_EMPTY = -1;

template <int H, bool NUMERIC>
struct SpgShared {
    int key[H];
    double w[NUMERIC ? H : 1];
    unsigned long long list[NUMERIC ? H / 2 : 1];
    int bc0[32], bpre[33];
    double bv[32];
};

// insert / find column j; returns the slot or -1 when the table is full.  *fresh = 1 if this lane created the entry
template <int H>
__device__ __forceinline__ int spg_slot(int* key, int j, int* fresh) {
    unsigned h = ((unsigned)j * 2654435761u) & (H - 1);
    *fresh = 0;
    for (int probe = 0; probe < H; probe++) {
        const int old = atomicCAS(key + h, SPG_EMPTY, j);
        if (old == SPG_EMPTY) { *fresh = 1; return (int)h; }
        if (old == j) return (int)h;
        h = (h + 1) & (H - 1);
    }
    return -1;
}

template <int H, int WARPS, bool NUMERIC>
__global__ void __launch_bounds__(32 * WARPS)
spgemm_kernel(int num_rows, const int* __restrict__ rows, const int* __restrict__ Bpos, const int* __restrict__ Bidx, const double* __restrict__ Bval,
              const int* __restrict__ Cpos, const int* __restrict__ Cidx, const double* __restrict__ Cval, int* __restrict__ counts,
              int* __restrict__ overflow_rows, int* __restrict__ overflow_count, const int* __restrict__ Apos, int* __restrict__ Aidx,
              double* __restrict__ Aval) {
    extern __shared__ __align__(16) unsigned char spg_raw[];
    SpgShared<H, NUMERIC>& S = reinterpret_cast<SpgShared<H, NUMERIC>*>(spg_raw)[threadIdx.x / 32];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    for (int t = blockIdx.x * WARPS + threadIdx.x / 32; t < num_rows; t += gridDim.x * WARPS) {
        const int row = rows ? rows[t] : t;
        for (int q = lane; q < H; q += 32) { S.key[q] = SPG_EMPTY; if (NUMERIC) S.w[q] = 0.; }
        __syncwarp();
        int distinct = 0;
        bool full = false;
        const int b0 = Bpos[row], b1 = Bpos[row + 1];
        for (int base = b0; base < b1 && !full; base += 32) {
            // 32 entries of B: their C rows, flattened
            const int pB = base + lane;
            int len = 0;
            if (pB < b1) {
                const int k = Bidx[pB];
                S.bc0[lane] = Cpos[k];
                len = Cpos[k + 1] - S.bc0[lane];
                if (NUMERIC) S.bv[lane] = Bval[pB];
            }
            int pre = len;  // inclusive warp scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += v; }
            S.bpre[lane + 1] = pre;
            if (lane == 0) S.bpre[0] = 0;
            __syncwarp();
            const int total = S.bpre[32];
            for (int c0 = 0; c0 < total; c0 += 32) {
                const int id = c0 + lane;
                const bool live = id < total;
                int slot = -1, fresh = 0;
                double b = 0., c = 0.;
                if (live) {
                    int lo = 0, hi = 31;  // last s with bpre[s] <= id
                    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S.bpre[mid] <= id) lo = mid; else hi = mid - 1; }
                    const int pC = S.bc0[lo] + (id - S.bpre[lo]);
                    slot = spg_slot<H>(S.key, Cidx[pC], &fresh);
                    if (NUMERIC) { b = S.bv[lo]; c = Cval[pC]; }
                }
                distinct += __popc(__ballot_sync(0xffffffffu, fresh != 0));
                if (__any_sync(0xffffffffu, (live && slot < 0)) || distinct > H / 2) { full = true; break; }
                if (NUMERIC) {
                    // candidates of one column in lane (= candidate) order
                    const unsigned peers = __match_any_sync(0xffffffffu, slot);
                    const int rank = live ? __popc(peers & lt) : 0;
                    int rounds = rank;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) rounds = max(rounds, __shfl_xor_sync(0xffffffffu, rounds, o));
                    for (int r = 0; r <= rounds; r++) {
                        if (live && rank == r) S.w[slot] = __fma_rn(b, c, S.w[slot]);
                        __syncwarp();
                    }
                }
            }
            __syncwarp();
        }
        if (full) {  // too many distinct columns for this table: hand the row to the next variant
            if (lane == 0) {
                if (overflow_rows) overflow_rows[atomicAdd(overflow_count, 1)] = row;
                else atomicAdd(overflow_count, 1);
                if (!NUMERIC) counts[row] = 0;
            }
            __syncwarp();
            continue;
        }
        if (!NUMERIC) {
            if (lane == 0) counts[row] = distinct;
            __syncwarp();
            continue;
        }
        // compact (column, slot) pairs, sort by column, write the row
        int m = 0;
        for (int q0 = 0; q0 < H; q0 += 32) {
            const int kq = S.key[q0 + lane];
            const unsigned has = __ballot_sync(0xffffffffu, kq != SPG_EMPTY);
            if (kq != SPG_EMPTY) S.list[m + __popc(has & lt)] = ((unsigned long long)(unsigned)kq << 32) | (unsigned)(q0 + lane);
            m += __popc(has);
        }
        int P2 = 1;
        while (P2 < m) P2 <<= 1;
        for (int q = m + lane; q < P2; q += 32) S.list[q] = ~0ull;
        __syncwarp();
        for (int size = 2; size <= P2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int q = lane; q < P2 / 2; q += 32) {
                    const int i = 2 * q - (q & (stride - 1)), j = i + stride;
                    const bool up = (i & size) == 0;
                    const unsigned long long a = S.list[i], bb = S.list[j];
                    if ((a > bb) == up) { S.list[i] = bb; S.list[j] = a; }
                }
                __syncwarp();
            }
        const int a0 = Apos[row];
        for (int q = lane; q < m; q += 32) {
            const unsigned long long e = S.list[q];
            Aidx[a0 + q] = (int)(e >> 32);
            Aval[a0 + q] = S.w[(int)(e & 0xffffffffu)];
        }
        __syncwarp();
    }
}

constexpr int SPG_H_FAST = 1024, SPG_W_FAST = 4, SPG_H_BIG = 8192, SPG_W_BIG = 1;

template <int H, int WARPS, bool NUMERIC>
static int spg_launch(int num_rows, const int* rows, const int* Bpos, const int* Bidx, const double* Bval, const int* Cpos, const int* Cidx,
                      const double* Cval, int* counts, int* overflow_rows, int* overflow_count, const int* Apos, int* Aidx, double* Aval,
                      cudaStream_t st) {
    if (num_rows == 0) return 0;
    const size_t bytes = sizeof(SpgShared<H, NUMERIC>) * WARPS;
    auto kern = spgemm_kernel<H, WARPS, NUMERIC>;
    DGB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const long long want = ((long long)num_rows + WARPS - 1) / WARPS, cap = (long long)sm_count() * 8;
    kern<<<(unsigned)std::min(want, cap), 32 * WARPS, bytes, st>>>(num_rows, rows, Bpos, Bidx, Bval, Cpos, Cidx, Cval, counts, overflow_rows,
                                                                      overflow_count, Apos, Aidx, Aval);
    DGB_LAUNCHED();
    return 0;
}

// scratch of one call: freed on every return path (DGB_CUDA returns early on errors)
template <class T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t count) { return cudaMalloc(&p, (count ? count : 1) * sizeof(T)); }
};

struct Spgemm {
    int rows = 0;
    long long nnz = 0;
    int* pos = nullptr;        // device row offsets of A
    int* big_rows = nullptr;   // rows that need the large table
    int big = 0;
    // device copies of the operands when the caller's arrays live on the host
    int *Bpos = nullptr, *Bidx = nullptr, *Cpos = nullptr, *Cidx = nullptr;
    double *Bval = nullptr, *Cval = nullptr;
    bool owns = false;
};

static void spg_free(Spgemm* s) {
    if (!s) return;
    cudaFree(s->pos); cudaFree(s->big_rows);
    if (s->owns) { cudaFree(s->Bpos); cudaFree(s->Bidx); cudaFree(s->Bval); cudaFree(s->Cpos); cudaFree(s->Cidx); cudaFree(s->Cval); }
    delete s;
}

// pass 1: distinct columns per row -> row offsets (scan on the host)
static int spg_symbolic(Spgemm* s, int B_rows, const int* Bpos, const int* Bidx, const int* Cpos, const int* Cidx, cudaStream_t st) {
    s->rows = B_rows;
    DevBuf<int> counts_buf, ocount_buf;
    DGB_CUDA(counts_buf.alloc((size_t)B_rows + 1));
    DGB_CUDA(cudaMalloc(&s->big_rows, ((size_t)B_rows + 1) * sizeof(int)));
    DGB_CUDA(ocount_buf.alloc(2));
    int *counts = counts_buf.p, *ocount = ocount_buf.p;
    DGB_CUDA(cudaMemsetAsync(ocount, 0, 2 * sizeof(int), st));
    int e = spg_launch<SPG_H_FAST, SPG_W_FAST, false>(B_rows, nullptr, Bpos, Bidx, nullptr, Cpos, Cidx, nullptr, counts, s->big_rows, ocount,
                                                      nullptr, nullptr, nullptr, st);
    int h_over[2] = {0, 0};
    if (!e) {
        DGB_CUDA(cudaMemcpyAsync(h_over, ocount, sizeof(int), cudaMemcpyDeviceToHost, st));
        DGB_CUDA(cudaStreamSynchronize(st));
        s->big = h_over[0];
        if (s->big) {
            e = spg_launch<SPG_H_BIG, SPG_W_BIG, false>(s->big, s->big_rows, Bpos, Bidx, nullptr, Cpos, Cidx, nullptr, counts, nullptr, ocount + 1,
                                                        nullptr, nullptr, nullptr, st);
            if (!e) {
                DGB_CUDA(cudaMemcpyAsync(h_over + 1, ocount + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
                DGB_CUDA(cudaStreamSynchronize(st));
                if (h_over[1]) { set_error("dgb_csr_spgemm: %d rows of the product have more than %d distinct columns", h_over[1], SPG_H_BIG / 2); e = DGB_ERR_UNSUPPORTED; }
            }
        }
    }
    if (!e) {
        std::vector<int> h((size_t)B_rows + 1);
        DGB_CUDA(cudaMemcpyAsync(h.data(), counts, (size_t)B_rows * sizeof(int), cudaMemcpyDeviceToHost, st));
        DGB_CUDA(cudaStreamSynchronize(st));
        long long run = 0;
        for (int i = 0; i < B_rows; i++) { const int c = h[i]; h[i] = (int)run; run += c; }
        if (run > 0x7fffffffll) { set_error("dgb_csr_spgemm: the product has %lld entries, too many for int offsets", run); e = DGB_ERR_UNSUPPORTED; }
        else {
            h[B_rows] = (int)run;
            s->nnz = run;
            DGB_CUDA(cudaMalloc(&s->pos, ((size_t)B_rows + 1) * sizeof(int)));
            DGB_CUDA(cudaMemcpyAsync(s->pos, h.data(), ((size_t)B_rows + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
            DGB_CUDA(cudaStreamSynchronize(st));
        }
    }
    return e;
}

static int spg_numeric(Spgemm* s, const int* Bpos, const int* Bidx, const double* Bval, const int* Cpos, const int* Cidx, const double* Cval,
                       int* Aidx, double* Aval, cudaStream_t st) {
    DevBuf<int> ocount_buf;
    DGB_CUDA(ocount_buf.alloc(1));
    int* ocount = ocount_buf.p;
    DGB_CUDA(cudaMemsetAsync(ocount, 0, sizeof(int), st));
    // the fast kernel skips the rows it cannot hold (they overflow again); the large variant fills them in
    int e = spg_launch<SPG_H_FAST, SPG_W_FAST, true>(s->rows, nullptr, Bpos, Bidx, Bval, Cpos, Cidx, Cval, nullptr, nullptr, ocount, s->pos, Aidx, Aval, st);
    if (!e && s->big)
        e = spg_launch<SPG_H_BIG, SPG_W_BIG, true>(s->big, s->big_rows, Bpos, Bidx, Bval, Cpos, Cidx, Cval, nullptr, nullptr, ocount, s->pos, Aidx, Aval, st);
    if (!e) DGB_CUDA(cudaStreamSynchronize(st));  // the scratch counter must outlive the kernels that bump it
    return e;
}

template <class T>
static int to_device(T** dst, const T* src, size_t count, cudaStream_t st) {
    DGB_CUDA(cudaMalloc(dst, (count ? count : 1) * sizeof(T)));
    if (count) DGB_CUDA(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_csr_spgemm_symbolic(dgb_spgemm** out, int B_rows, int B_cols, int C_cols, const int* B_pos, const int* B_idx, const int* C_pos,
                            const int* C_idx, long long* nnz, dgb_stream_t s) {
    if (!out || B_rows < 0 || B_cols < 0 || C_cols < 0 || !B_pos || !C_pos || !nnz) { set_error("dgb_csr_spgemm_symbolic: invalid argument"); return DGB_ERR_INVALID; }
    Spgemm* p = new Spgemm();
    int e = spg_symbolic(p, B_rows, B_pos, B_idx, C_pos, C_idx, as_stream(s));
    if (e) { spg_free(p); return e; }
    *nnz = p->nnz;
    *out = reinterpret_cast<dgb_spgemm*>(p);
    return 0;
}
int dgb_csr_spgemm_numeric(dgb_spgemm* h, const int* B_pos, const int* B_idx, const double* B_val, const int* C_pos, const int* C_idx,
                           const double* C_val, int* A_pos, int* A_idx, double* A_val, dgb_stream_t s) {
    Spgemm* p = reinterpret_cast<Spgemm*>(h);
    if (!p || !A_pos || (p->nnz && (!A_idx || !A_val))) { set_error("dgb_csr_spgemm_numeric: invalid argument"); return DGB_ERR_INVALID; }
    cudaStream_t st = as_stream(s);
    DGB_CUDA(cudaMemcpyAsync(A_pos, p->pos, ((size_t)p->rows + 1) * sizeof(int), cudaMemcpyDeviceToDevice, st));
    return spg_numeric(p, B_pos, B_idx, B_val, C_pos, C_idx, C_val, A_idx, A_val, st);
}
int dgb_csr_spgemm_destroy(dgb_spgemm* h) { spg_free(reinterpret_cast<Spgemm*>(h)); return 0; }

// the same for HOST arrays (the reference's operator* works on thrust::host_vector): begin uploads the operands and returns the
// number of entries of the product, finish downloads it into caller-allocated arrays and releases everything
int dgb_csr_spgemm_host_begin(dgb_spgemm** out, int B_rows, int B_cols, int C_cols, const int* B_pos, const int* B_idx, const double* B_val,
                              const int* C_pos, const int* C_idx, const double* C_val, long long* nnz) {
    if (!out || B_rows < 0 || B_cols < 0 || C_cols < 0 || !B_pos || !C_pos || !nnz) { set_error("dgb_csr_spgemm_host_begin: invalid argument"); return DGB_ERR_INVALID; }
    Spgemm* p = new Spgemm();
    p->owns = true;
    const size_t nb = (size_t)B_pos[B_rows], nc = (size_t)C_pos[B_cols];
    int e = 0;
    if (!e) e = to_device(&p->Bpos, B_pos, (size_t)B_rows + 1, nullptr);
    if (!e) e = to_device(&p->Bidx, B_idx, nb, nullptr);
    if (!e) e = to_device(&p->Bval, B_val, nb, nullptr);
    if (!e) e = to_device(&p->Cpos, C_pos, (size_t)B_cols + 1, nullptr);
    if (!e) e = to_device(&p->Cidx, C_idx, nc, nullptr);
    if (!e) e = to_device(&p->Cval, C_val, nc, nullptr);
    if (!e) e = spg_symbolic(p, B_rows, p->Bpos, p->Bidx, p->Cpos, p->Cidx, nullptr);
    if (e) { spg_free(p); return e; }
    *nnz = p->nnz;
    *out = reinterpret_cast<dgb_spgemm*>(p);
    return 0;
}
int dgb_csr_spgemm_host_finish(dgb_spgemm* h, int* A_pos, int* A_idx, double* A_val) {
    Spgemm* p = reinterpret_cast<Spgemm*>(h);
    if (!p || !p->owns || !A_pos || (p->nnz && (!A_idx || !A_val))) { set_error("dgb_csr_spgemm_host_finish: invalid argument"); spg_free(p); return DGB_ERR_INVALID; }
    int* d_idx = nullptr;
    double* d_val = nullptr;
    int e = 0;
    if (cudaMalloc(&d_idx, (size_t)(p->nnz ? p->nnz : 1) * sizeof(int)) != cudaSuccess || cudaMalloc(&d_val, (size_t)(p->nnz ? p->nnz : 1) * sizeof(double)) != cudaSuccess) {
        set_error("dgb_csr_spgemm_host_finish: out of device memory"); e = DGB_ERR_INVALID;
    }
    if (!e) e = spg_numeric(p, p->Bpos, p->Bidx, p->Bval, p->Cpos, p->Cidx, p->Cval, d_idx, d_val, nullptr);
    if (!e) {
        cudaError_t c = cudaMemcpy(A_pos, p->pos, ((size_t)p->rows + 1) * sizeof(int), cudaMemcpyDeviceToHost);
        if (c == cudaSuccess && p->nnz) c = cudaMemcpy(A_idx, d_idx, (size_t)p->nnz * sizeof(int), cudaMemcpyDeviceToHost);
        if (c == cudaSuccess && p->nnz) c = cudaMemcpy(A_val, d_val, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToHost);
        if (c != cudaSuccess) { set_error("dgb_csr_spgemm_host_finish: %s", cudaGetErrorString(c)); e = (int)c; }
    }
    cudaFree(d_idx); cudaFree(d_val);
    spg_free(p);
    return e;
}
}
