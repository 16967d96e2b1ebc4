// Bit-reproducible dot products: dgb_exdot2 / dgb_exdot3 and friends.
// Replaces exdot_gpu (inc/dg/backend/exblas/exdot_cuda.cuh:319-357: fixed 64x512 threads, three launches, a
// blocking D2H of 39 words and a host-side Round) by ONE persistent kernel sized to the SM count whose last block
// normalises, rounds and publishes {acc[39], value, status} in device memory.
#include "superacc.cuh"
#include "async_copy.cuh"
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace dgb {

struct DotWs {
    sa::DotSlot slot;      // global accumulators / status / ticket / (result unused here)
    dgb_dot_result* result;  // default device result record
    dgb_dot_result* host_result;  // pinned
    int max_blocks;
    int nslots;
};

constexpr int DOT_THREADS = 256;
constexpr int DOT_WARPS = DOT_THREADS / 32;

// products follow the reference exactly: 2 operands round(x*y); 3 operands round(round(x*w)*y)
// (exdot_cuda.cuh:54,147-148); non-finite products raise status and are not accumulated.
template <int NOPS>
__device__ __forceinline__ double dot_product(double a, double b, double c, int& bad) {
    double p = NOPS == 3 ? __dmul_rn(__dmul_rn(a, b), c) : __dmul_rn(a, c);
    if (!isfinite(p)) { bad = 1; p = 0.; }
    return p;
}

// One pass, software pipelined: the U 128-bit loads per operand of the NEXT trip are in flight while the current trip
// goes through the floating-point expansion, so with BPS resident CTAs every SM keeps
// BPS * 256 * NOPS * U * 16 bytes outstanding all the time (the kernel is a pure HBM stream: 16 / 24 B per element).
template <int NOPS, int U, int BPS, int NE>  // NE independent expansions per thread (NE divides 2 U)
__global__ void __launch_bounds__(DOT_THREADS, BPS)
exdot_kernel(const double* __restrict__ x, double xs, const double* __restrict__ w, double wsc,
             const double* __restrict__ y, double ys, size_t n, sa::DotSlot slot) {
    __shared__ long long smem[sa::BINS];  // one accumulator per block
    sa::block_init<1>(smem);
    long long* my = smem;
    sa::Fpe fpe[NE];  // independent expansions: their add cascades interleave in the FP64 pipe
#pragma unroll
    for (int u = 0; u < NE; u++) fpe[u].clear();
    int bad = 0;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nvec = n / 2;
    const double2 z2 = make_double2(0., 0.);
    double2 a[U], b[U], c[U];
    auto load = [&](size_t base, double2 (&A)[U], double2 (&B)[U], double2 (&Cc)[U]) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t idx = base + (size_t)u * T;
            const bool ok = idx < nvec;
            A[u] = ok ? (x ? ld2(x + 2 * idx) : make_double2(xs, xs)) : z2;
            if (NOPS == 3) B[u] = ok ? (w ? ld2(w + 2 * idx) : make_double2(wsc, wsc)) : z2;
            Cc[u] = ok ? (y ? ld2(y + 2 * idx) : make_double2(ys, ys)) : z2;
        }
    };
    size_t i = tid;
    load(i, a, b, c);
    while (i < nvec) {
        const size_t inext = i + (size_t)U * T;
        double2 an[U], bn[U], cn[U];
        load(inext, an, bn, cn);  // everything beyond nvec loads as zero (adds +0 exactly)
        double res[2 * U];
        bool spill = false;
#pragma unroll
        for (int u = 0; u < U; u++) {
            res[2 * u] = fpe[(2 * u) % NE].add_lazy(dot_product<NOPS>(a[u].x, b[u].x, c[u].x, bad));
            res[2 * u + 1] = fpe[(2 * u + 1) % NE].add_lazy(dot_product<NOPS>(a[u].y, b[u].y, c[u].y, bad));
            spill = spill || res[2 * u] != 0.0 || res[2 * u + 1] != 0.0;
        }
        if (spill) {  // rare: a residue the expansions cannot hold goes to the shared accumulator (exact)
#pragma unroll
            for (int u = 0; u < 2 * U; u++) sa::accumulate(my, res[u], 1);  // unrolled: res[] stays in registers
        }
#pragma unroll
        for (int u = 0; u < U; u++) { a[u] = an[u]; b[u] = bn[u]; c[u] = cn[u]; }
        i = inext;
    }
    if ((n & 1) && tid == 0)
        fpe[0].add(dot_product<NOPS>(x ? x[n - 1] : xs, NOPS == 3 ? (w ? w[n - 1] : wsc) : 0., y ? y[n - 1] : ys, bad), my);
#pragma unroll
    for (int u = 1; u < NE; u++) fpe[0].merge(fpe[u], my);
    fpe[0].flush_warp(my);
    sa::block_finish<1>(smem, bad, slot);
}

// ---------------------------------------------------------------------------------------------------------------------
// Streaming variant for large vectors: the operands reach the SM through the TMA engine.
// The register-prefetch kernel above keeps one trip of loads per thread in flight (~49 KB per SM), which at the loaded
// HBM latency of B200 (> 1 us) caps it near 4.6 TB/s.  Here one producer lane per CTA keeps a ring of STAGES chunks per
// operand in flight with 1-d bulk copies (cp.async.bulk, SASS UBLKCP; completion counted in bytes on an mbarrier), i.e.
// up to 192 KB per SM, independent of the register budget; 16 consumer warps take the chunks out of shared memory with
// conflict-free 128-bit loads and run them through four independent floating-point expansions each.  Chunks are dealt
// round robin to one persistent CTA per SM; the exact arithmetic, the accumulator and the finish are those of the kernel
// above, so the result words are identical (integer accumulation is order independent).
constexpr int TDOT_CONSUMERS = 512, TDOT_THREADS = TDOT_CONSUMERS + 32, TDOT_CHUNK = 2048;  // doubles per operand and stage
template <int NOPS>
struct TDot {
    static constexpr int STAGES = NOPS == 3 ? 4 : 6;
    static constexpr unsigned STAGE_BYTES = NOPS * TDOT_CHUNK * 8;
    static constexpr size_t SMEM = (size_t)STAGES * STAGE_BYTES + 2 * STAGES * 8 + 16;
};
template <int NOPS, bool PLAIN = false>
__global__ void __launch_bounds__(TDOT_THREADS, 1)
exdot_tma_kernel(const double* __restrict__ x, const double* __restrict__ w, const double* __restrict__ y, size_t n, sa::DotSlot slot) {
    using P = TDot<NOPS>;
    constexpr int STAGES = P::STAGES;
    __shared__ long long acc_sm[sa::BINS];
    extern __shared__ __align__(128) unsigned char ring_raw[];
    double* ring = reinterpret_cast<double*>(ring_raw);
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring_raw + (size_t)STAGES * P::STAGE_BYTES);
    unsigned long long* empty = full + STAGES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < sa::BINS; i += blockDim.x) acc_sm[i] = 0;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full + s, 1); mbar_init(empty + s, TDOT_CONSUMERS / 32); }
    }
    __syncthreads();
    const size_t nchunks = n / TDOT_CHUNK;
    // chunks of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int mine = nchunks > blockIdx.x ? (int)((nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    int bad = 0;
    sa::Fpe fpe[4];
#pragma unroll
    for (int u = 0; u < 4; u++) fpe[u].clear();
    if (warp == TDOT_CONSUMERS / 32) {
        // ---- producer: one lane issues every copy of this CTA
        if (lane == 0) {
            int s = 0;
            unsigned ph = 1;  // parity of the PREVIOUS use of slot s (first round: nothing to wait for)
            for (int it = 0; it < mine; it++) {
                if (it >= STAGES) mbar_wait(empty + s, ph);
                const size_t base = ((size_t)blockIdx.x + (size_t)it * gridDim.x) * (size_t)TDOT_CHUNK;
                double* dst = ring + (size_t)s * NOPS * TDOT_CHUNK;
                mbar_expect_tx(full + s, P::STAGE_BYTES);
                bulk_load_1d(dst, x + base, TDOT_CHUNK * 8, full + s);
                if (NOPS == 3) bulk_load_1d(dst + TDOT_CHUNK, w + base, TDOT_CHUNK * 8, full + s);
                bulk_load_1d(dst + (NOPS - 1) * TDOT_CHUNK, y + base, TDOT_CHUNK * 8, full + s);
                if (++s == STAGES) { s = 0; ph ^= 1u; }
            }
        }
    } else {
        // ---- consumers
        int s = 0;
        unsigned ph = 0;
        for (int it = 0; it < mine; it++) {
            mbar_wait(full + s, ph);
            const double* src = ring + (size_t)s * NOPS * TDOT_CHUNK;
            double2 a[2], b[2], c[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int e = 2 * (tid + u * TDOT_CONSUMERS);
                a[u] = *reinterpret_cast<const double2*>(src + e);
                if (NOPS == 3) b[u] = *reinterpret_cast<const double2*>(src + TDOT_CHUNK + e);
                c[u] = *reinterpret_cast<const double2*>(src + (NOPS - 1) * TDOT_CHUNK + e);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);  // the operands are in registers: the slot may be refilled
            if (++s == STAGES) { s = 0; ph ^= 1u; }
            if (PLAIN) {  // timing experiment only (DGB_DOT_DEBUG_PLAIN): the memory pipeline without the exact arithmetic
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    fpe[2 * u].a[0] += dot_product<NOPS>(a[u].x, b[u].x, c[u].x, bad);
                    fpe[2 * u + 1].a[0] += dot_product<NOPS>(a[u].y, b[u].y, c[u].y, bad);
                }
                continue;
            }
            double res[4];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                res[2 * u] = fpe[2 * u].add_lazy(dot_product<NOPS>(a[u].x, b[u].x, c[u].x, bad));
                res[2 * u + 1] = fpe[2 * u + 1].add_lazy(dot_product<NOPS>(a[u].y, b[u].y, c[u].y, bad));
            }
            if (res[0] != 0.0 || res[1] != 0.0 || res[2] != 0.0 || res[3] != 0.0) {  // rare: exact spill to the shared accumulator
#pragma unroll
                for (int u = 0; u < 4; u++) sa::accumulate(acc_sm, res[u], 1);
            }
        }
        // elements behind the last full chunk: block 0, straight from global memory
        if (blockIdx.x == 0) {
            for (size_t i = nchunks * TDOT_CHUNK + tid; i < n; i += TDOT_CONSUMERS)
                fpe[0].add(dot_product<NOPS>(x[i], NOPS == 3 ? w[i] : 0., y[i], bad), acc_sm);
        }
#pragma unroll
        for (int u = 1; u < 4; u++) fpe[0].merge(fpe[u], acc_sm);
    }
    fpe[0].flush_warp(acc_sm);  // the producer warp carries empty expansions: adds nothing
    sa::block_finish<1>(acc_sm, bad, slot);
}

// operands that are not 16-byte aligned: scalar loads, same arithmetic
template <int NOPS>
__global__ void __launch_bounds__(DOT_THREADS)
exdot_scalar_kernel(const double* __restrict__ x, double xs, const double* __restrict__ w, double wsc,
                    const double* __restrict__ y, double ys, size_t n, sa::DotSlot slot) {
    __shared__ long long smem[sa::BINS];  // one accumulator per block
    sa::block_init<1>(smem);
    long long* my = smem;
    sa::Fpe fpe;
    fpe.clear();
    int bad = 0;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T)
        fpe.add(dot_product<NOPS>(x ? x[i] : xs, NOPS == 3 ? (w ? w[i] : wsc) : 0., y ? y[i] : ys, bad), my);
    fpe.flush_warp(my);
    sa::block_finish<1>(smem, bad, slot);
}

// combine `nparts` normalised accumulators (multi-GPU / recursive vectors)
__global__ void __launch_bounds__(64) superacc_combine_kernel(const long long* parts, int nparts, const int* status,
                                                              dgb_dot_result* result) {
    __shared__ long long acc[sa::BINS + 1];
    if (threadIdx.x < sa::BINS) acc[threadIdx.x] = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nparts; b0 += 64) {  // 65 normalised words (< 2^56 each) cannot overflow int64
        int b1 = min(b0 + 64, nparts);
        if (threadIdx.x < sa::BINS) {
            long long sum = acc[threadIdx.x];
            for (int b = b0; b < b1; b++) sum += parts[(size_t)b * sa::BINS + threadIdx.x];
            acc[threadIdx.x] = sum;
        }
        __syncthreads();
        if (threadIdx.x == 0) acc[sa::BINS] = sa::normalize(acc, 1);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        int bad = 0;
        if (status)
            for (int b = 0; b < nparts; b++) bad |= status[b];
        for (int i = 0; i < sa::BINS; i++) result->acc[i] = acc[i];
        result->value = sa::round_normalized(acc, (int)acc[sa::BINS]);
        result->status = bad;
        result->pad = 0;
    }
}

// tuning knob for experiments: DGB_DOT_VARIANT = 0 (U=2, 3 CTAs/SM, 2 FPE)  1 (U=4, 2 CTAs/SM, 4 FPE)  2 (U=2, 2 CTAs/SM, 4 FPE)  3 (U=2, 3 CTAs/SM, 4 FPE)
static int dot_variant() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DGB_DOT_VARIANT"); v = e ? atoi(e) : 0; if (v < 0 || v > 3) v = 0; }
    return v;
}
template <int NOPS, int U, int BPS, int NE>
static void exdot_go(int max_blocks, size_t n, const double* x, double xs, const double* w, double wsc, const double* y, double ys,
                     const sa::DotSlot& slot, cudaStream_t st) {
    const size_t per_block = (size_t)DOT_THREADS * 2 * U;
    size_t want = (n + per_block - 1) / per_block;
    if (want == 0) want = 1;
    static int waves = 0;  // CTAs beyond one resident wave are balanced by the hardware block scheduler
    if (waves == 0) { const char* e = getenv("DGB_DOT_WAVES"); waves = e ? atoi(e) : 1; if (waves < 1) waves = 1; }
    size_t cap = (size_t)sm_count() * BPS * waves;
    if (cap > (size_t)max_blocks) cap = max_blocks;
    exdot_kernel<NOPS, U, BPS, NE><<<(unsigned)(want < cap ? want : cap), DOT_THREADS, 0, st>>>(x, xs, w, wsc, y, ys, n, slot);
}
template <int NOPS>
static void exdot_pick(int max_blocks, size_t n, const double* x, double xs, const double* w, double wsc, const double* y, double ys,
                       const sa::DotSlot& slot, cudaStream_t st) {
    switch (dot_variant()) {
        case 1: exdot_go<NOPS, 4, 2, 4>(max_blocks, n, x, xs, w, wsc, y, ys, slot, st); break;
        case 2: exdot_go<NOPS, 2, 2, 4>(max_blocks, n, x, xs, w, wsc, y, ys, slot, st); break;
        case 3: exdot_go<NOPS, 2, 3, 4>(max_blocks, n, x, xs, w, wsc, y, ys, slot, st); break;
        default: exdot_go<NOPS, 2, 3, 2>(max_blocks, n, x, xs, w, wsc, y, ys, slot, st); break;
    }
}

int exdot_launch(DotWs* ws, int nops, size_t n, const double* x, double xs, const double* w, double wsc,
                 const double* y, double ys, dgb_dot_result* result, dgb_stream_t s) {
    if (!ws) { set_error("dgb_exdot: workspace is NULL"); return DGB_ERR_INVALID; }
    sa::DotSlot slot = ws->slot;
    slot.result = result ? result : ws->result;
    bool vec = (!x || aligned16(x)) && (!w || aligned16(w)) && (!y || aligned16(y));
    cudaStream_t st = as_stream(s);
    static int tma_min = -1;  // vectors at least this long take the TMA-pipelined kernel (DGB_DOT_TMA_MIN; 0 disables it)
    if (tma_min < 0) { const char* e = getenv("DGB_DOT_TMA_MIN"); tma_min = e ? atoi(e) : 1 << 20; }
    const bool all_vectors = x && y && (nops == 2 || w);
    if (vec && all_vectors && tma_min > 0 && n >= (size_t)tma_min) {
        const size_t nchunks = n / TDOT_CHUNK;
        const unsigned grid = (unsigned)(nchunks < (size_t)sm_count() ? (nchunks ? nchunks : 1) : (size_t)sm_count());
        static bool configured = false;
        if (!configured) {
            DGB_CUDA(cudaFuncSetAttribute(exdot_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TDot<2>::SMEM));
            DGB_CUDA(cudaFuncSetAttribute(exdot_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TDot<3>::SMEM));
            configured = true;
        }
        static int plain = -1;
        if (plain < 0) {
            const char* e = getenv("DGB_DOT_DEBUG_PLAIN");
            plain = (e && atoi(e)) ? 1 : 0;
            if (plain) DGB_CUDA(cudaFuncSetAttribute(exdot_tma_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TDot<2>::SMEM));
        }
        if (plain && nops == 2) exdot_tma_kernel<2, true><<<grid, TDOT_THREADS, TDot<2>::SMEM, st>>>(x, nullptr, y, n, slot);
        else if (nops == 3) exdot_tma_kernel<3><<<grid, TDOT_THREADS, TDot<3>::SMEM, st>>>(x, w, y, n, slot);
        else exdot_tma_kernel<2><<<grid, TDOT_THREADS, TDot<2>::SMEM, st>>>(x, nullptr, y, n, slot);
    } else if (vec) {
        if (nops == 3) exdot_pick<3>(ws->max_blocks, n, x, xs, w, wsc, y, ys, slot, st);
        else exdot_pick<2>(ws->max_blocks, n, x, xs, nullptr, 0., y, ys, slot, st);
    } else {
        size_t want = (n + DOT_THREADS - 1) / DOT_THREADS;
        if (want == 0) want = 1;
        size_t cap = (size_t)sm_count() * 4;
        const unsigned grid = (unsigned)(want < cap ? want : cap);
        if (nops == 3) exdot_scalar_kernel<3><<<grid, DOT_THREADS, 0, st>>>(x, xs, w, wsc, y, ys, n, slot);
        else exdot_scalar_kernel<2><<<grid, DOT_THREADS, 0, st>>>(x, xs, nullptr, 0., y, ys, n, slot);
    }
    DGB_LAUNCHED();
    return 0;
}

// host restatement of accumulate.h:267-349 for the convenience helpers (operates on host memory only)
static int normalize_host(int64_t* acc) {
    int64_t carry_in = acc[0] >> 56;
    acc[0] -= (int64_t)((uint64_t)carry_in << 56);
    for (int i = 1; i < 39; ++i) {
        int64_t v = acc[i] + carry_in;
        int64_t carry_out = v >> 56;
        acc[i] = v - (int64_t)((uint64_t)carry_out << 56);
        carry_in = carry_out;
    }
    acc[38] += (int64_t)((uint64_t)carry_in << 56);
    return carry_in < 0;
}
static double round_host(const int64_t* in) {
    int64_t acc[39];
    memcpy(acc, in, sizeof(acc));
    int negative = normalize_host(acc);
    const int64_t MASK = (1ll << 56) - 1;
    int i;
    for (i = 38; i >= 0 && acc[i] == 0; --i) {}
    if (negative)
        for (; i >= 0 && (acc[i] & MASK) == MASK; --i) {}
    if (i < 0) return 0.0;
    int64_t hiword = negative ? MASK - acc[i] : acc[i];
    double rounded = (double)hiword;
    double hi = std::ldexp(rounded, (i - 20) * 56);
    if (i == 0) return negative ? -hi : hi;
    hiword -= std::llrint(rounded);
    double mid = std::ldexp((double)hiword, (i - 20) * 56);
    int64_t sticky = 0;
    for (int j = 0; j != i - 1; ++j) sticky |= negative ? ((1ll << 56) - acc[j]) : acc[j];
    int64_t loword = negative ? ((1ll << 56) - acc[i - 1]) : acc[i - 1];
    loword |= !!sticky;
    double lo = std::ldexp((double)loword, (i - 1 - 20) * 56);
    if (mid != 0) {
        union { double d; int64_t l; } u;
        u.d = mid + lo;
        u.l |= (lo != 0.0);
        lo = u.d;
    }
    hi = hi + lo;
    return negative ? -hi : hi;
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_dot_ws_create(dgb_dot_ws** out) {
    DotWs* ws = new DotWs();
    ws->max_blocks = 1 << 16;
    ws->nslots = 4;  // fused kernels may carry up to 4 simultaneous dots
    DGB_CUDA(cudaMalloc(&ws->slot.gacc, ws->nslots * sa::GACC_WORDS * sizeof(long long)));
    DGB_CUDA(cudaMemset(ws->slot.gacc, 0, ws->nslots * sa::GACC_WORDS * sizeof(long long)));
    DGB_CUDA(cudaMalloc(&ws->slot.gstatus, ws->nslots * sizeof(int)));
    DGB_CUDA(cudaMemset(ws->slot.gstatus, 0, ws->nslots * sizeof(int)));
    DGB_CUDA(cudaMalloc(&ws->slot.ticket, ws->nslots * sizeof(unsigned int)));
    DGB_CUDA(cudaMemset(ws->slot.ticket, 0, ws->nslots * sizeof(unsigned int)));
    DGB_CUDA(cudaMalloc(&ws->result, ws->nslots * sizeof(dgb_dot_result)));
    DGB_CUDA(cudaMemset(ws->result, 0, ws->nslots * sizeof(dgb_dot_result)));
    DGB_CUDA(cudaMallocHost(&ws->host_result, ws->nslots * sizeof(dgb_dot_result)));
    ws->slot.result = ws->result;
    *out = reinterpret_cast<dgb_dot_ws*>(ws);
    return 0;
}
int dgb_dot_ws_destroy(dgb_dot_ws* p) {
    DotWs* ws = reinterpret_cast<DotWs*>(p);
    if (!ws) return 0;
    cudaFree(ws->slot.gacc);
    cudaFree(ws->slot.gstatus);
    cudaFree(ws->slot.ticket);
    cudaFree(ws->result);
    cudaFreeHost(ws->host_result);
    delete ws;
    return 0;
}
int dgb_exdot2(dgb_dot_ws* ws, size_t n, const double* x, double xs, const double* y, double ys,
               dgb_dot_result* result, dgb_stream_t s) {
    return exdot_launch(reinterpret_cast<DotWs*>(ws), 2, n, x, xs, nullptr, 0., y, ys, result, s);
}
int dgb_exdot3(dgb_dot_ws* ws, size_t n, const double* x, double xs, const double* w, double wsc, const double* y,
               double ys, dgb_dot_result* result, dgb_stream_t s) {
    return exdot_launch(reinterpret_cast<DotWs*>(ws), 3, n, x, xs, w, wsc, y, ys, result, s);
}
static int dot_sync(DotWs* ws, int64_t* acc_host, double* value, int* status, dgb_stream_t s) {
    DGB_CUDA(cudaMemcpyAsync(ws->host_result, ws->result, sizeof(dgb_dot_result), cudaMemcpyDeviceToHost, as_stream(s)));
    DGB_CUDA(cudaStreamSynchronize(as_stream(s)));
    if (acc_host) memcpy(acc_host, ws->host_result->acc, sizeof(int64_t) * DGB_BIN_COUNT);
    if (value) *value = ws->host_result->value;
    if (status) *status = ws->host_result->status;
    if (ws->host_result->status != 0) {
        set_error("dot product failed since one of the inputs contains NaN or Inf");
        return DGB_ERR_NOTFINITE;
    }
    return 0;
}
int dgb_dot2(dgb_dot_ws* p, size_t n, const double* x, const double* y, int64_t* acc_host, double* value, int* status,
             dgb_stream_t s) {
    DotWs* ws = reinterpret_cast<DotWs*>(p);
    int e = exdot_launch(ws, 2, n, x, 0., nullptr, 0., y, 0., nullptr, s);
    if (e) return e;
    return dot_sync(ws, acc_host, value, status, s);
}
int dgb_dot3(dgb_dot_ws* p, size_t n, const double* x, const double* w, const double* y, int64_t* acc_host,
             double* value, int* status, dgb_stream_t s) {
    DotWs* ws = reinterpret_cast<DotWs*>(p);
    int e = exdot_launch(ws, 3, n, x, 0., w, 0., y, 0., nullptr, s);
    if (e) return e;
    return dot_sync(ws, acc_host, value, status, s);
}
int dgb_superacc_normalize_host(int64_t* acc, int* negative) {
    const int neg = normalize_host(acc);
    if (negative) *negative = neg;
    return 0;
}
double dgb_superacc_round_host(const int64_t* acc) { return round_host(acc); }
int dgb_superacc_combine(const int64_t* parts, int nparts, const int32_t* status_parts, dgb_dot_result* result,
                         dgb_stream_t s) {
    if (nparts < 1) { set_error("dgb_superacc_combine: nparts < 1"); return DGB_ERR_INVALID; }
    superacc_combine_kernel<<<1, 64, 0, as_stream(s)>>>(reinterpret_cast<const long long*>(parts), nparts,
                                                        status_parts, result);
    DGB_LAUNCHED();
    return 0;
}
}
