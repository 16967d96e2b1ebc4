// Device-side long accumulator ("superaccumulator") arithmetic for the bit-reproducible dot.
//
// Number format (must stay compatible with exblas::cpu::Normalize/Round, inc/dg/backend/exblas/config.h:86-92):
// 39 signed 64-bit words, word i has weight 2^(56*(i-20)); a NORMALISED accumulator has every word in
// [0, 2^56) except the top one, which carries the sign (accumulate.h:267-285).
//
// The design here is our own (not a port of exdot_cuda.cuh):
//  * each thread keeps a small floating-point expansion (FPE) of NF doubles in registers; adding a value is a
//    cascade of error-free TwoSums, so the expansion holds its inputs' sum EXACTLY; a non-zero residue (rare for
//    data of bounded dynamic range) is spilled to a shared-memory accumulator;
//  * the shared accumulators (one per warp) are updated with native 64-bit shared atomics; signed wrap-around of
//    a word is detected from the value the atomic returns and compensated in the next word, so the update is
//    value-preserving under any interleaving;
//  * the per-block result is normalised and added word by word into one global accumulator with the same wrap-safe
//    atomics (integer addition is associative => any grid size and arrival order gives the same normalised words);
//    the last block to finish (atomic ticket) normalises, rounds (accumulate.h:297-349, replicated operation by
//    operation) and writes the result record.
#pragma once
#include "common.cuh"

namespace dgb {
namespace sa {

constexpr int BINS = 39;
constexpr int DIGITS = 56;
constexpr int KRX = 8;
constexpr int F_WORDS = 20;
constexpr int NF = 3;  // FPE size in registers
constexpr int SPREAD = 16;  // global accumulators per dot: CTAs spread their partials to cut same-address atomic contention
constexpr int GACC_WORDS = SPREAD * BINS;  // int64 words of global scratch per slot

// error-free transformation: a + b = r + s exactly (Knuth TwoSum, 6 flops, no branch)
__device__ __forceinline__ double two_sum(double a, double b, double& s) {
    double r = __dadd_rn(a, b);
    double z = __dsub_rn(r, a);
    s = __dadd_rn(__dsub_rn(a, __dsub_rn(r, z)), __dsub_rn(b, z));
    return r;
}

// acc[i] += x with detection of signed 64-bit wrap-around; shared or global memory.
// Returns +1 if the true sum exceeded 2^63-1 (stored value is 2^64 too small), -1 if below -2^63, else 0.
__device__ __forceinline__ int atomic_add_wrap(long long* w, long long x, long long& stored) {
    unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(w), (unsigned long long)x);
    long long o = (long long)old;
    long long nw = (long long)(old + (unsigned long long)x);
    stored = nw;
    // signed overflow iff operands have the same sign and the result's sign differs
    bool of = ((o ^ nw) & (x ^ nw)) < 0;
    if (!of) return 0;
    return o > 0 ? 1 : -1;
}

// add the integer x to word i of a (possibly shared or global, concurrently updated) accumulator with element
// stride `stride`.  Never loses a bit: when the 64-bit word wraps, the lost +-2^64 equals +-2^8 units of the next
// word (2^64 = 2^8 * 2^56) and is added there; the wrapped word itself stays a valid two's complement digit, so the
// represented VALUE sum_i acc[i] 2^(56 (i-20)) is preserved under any interleaving of the atomics.
__device__ inline void add_word(long long* acc, int i, long long x, int stride) {
    while (i < BINS) {
        long long stored;
        int wrap = atomic_add_wrap(&acc[i * stride], x, stored);
        if (wrap == 0) return;
        x = (long long)wrap * (1ll << KRX);
        ++i;
    }
}

// add the double x exactly to the accumulator (decomposition into 56-bit digits, cf. accumulate.h:217-236)
static __device__ __noinline__ void accumulate(long long* acc, double x, int stride) {
    if (x == 0.0) return;
    int e = ((int)((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ff) - 0x3ff;
    int exp_word = e / DIGITS;  // truncation toward zero, as the reference
    int iup = exp_word + F_WORDS;
    double xscaled = ldexp(x, -DIGITS * exp_word);
    for (int i = iup; i >= 0 && xscaled != 0.0; --i) {
        double xr = rint(xscaled);
        long long xi = __double2ll_rn(xscaled);
        add_word(acc, i, xi, stride);
        xscaled = __dsub_rn(xscaled, xr);
        xscaled = __dmul_rn(xscaled, 72057594037927936.0);  // 2^56
    }
}

// sequential carry propagation (accumulate.h:267-285); returns 1 if negative
__device__ inline int normalize(long long* acc, int stride) {
    long long carry_in = acc[0] >> DIGITS;
    acc[0] -= (long long)((unsigned long long)carry_in << DIGITS);
    int i;
    for (i = 1; i < BINS; ++i) {
        long long v = acc[i * stride] + carry_in;
        long long carry_out = v >> DIGITS;
        acc[i * stride] = v - (long long)((unsigned long long)carry_out << DIGITS);
        carry_in = carry_out;
    }
    acc[(BINS - 1) * stride] += (long long)((unsigned long long)carry_in << DIGITS);
    return carry_in < 0;
}

// accumulate.h:297-349 replicated step by step on an already normalised accumulator (sign given)
__device__ inline double round_normalized(const long long* acc, int negative) {
    const long long MASK = (1ll << DIGITS) - 1;
    int i;
    for (i = BINS - 1; i >= 0 && acc[i] == 0; --i) {}
    if (negative) {
        for (; i >= 0 && (acc[i] & MASK) == MASK; --i) {}
    }
    if (i < 0) return 0.0;
    long long hiword = negative ? MASK - acc[i] : acc[i];
    double rounded = __ll2double_rn(hiword);
    double hi = ldexp(rounded, (i - F_WORDS) * DIGITS);
    if (i == 0) return negative ? -hi : hi;
    hiword -= __double2ll_rn(rounded);
    double mid = ldexp(__ll2double_rn(hiword), (i - F_WORDS) * DIGITS);
    long long sticky = 0;
    for (int j = 0; j != i - 1; ++j) sticky |= negative ? ((1ll << DIGITS) - acc[j]) : acc[j];
    long long loword = negative ? ((1ll << DIGITS) - acc[i - 1]) : acc[i - 1];
    loword |= (sticky != 0);
    double lo = ldexp(__ll2double_rn(loword), (i - 1 - F_WORDS) * DIGITS);
    if (mid != 0.0) {
        // OddRoundSumNonnegative, mylibm.hpp:118-134
        long long l = __double_as_longlong(__dadd_rn(mid, lo));
        l |= (long long)(lo != 0.0);
        lo = __longlong_as_double(l);
    }
    hi = __dadd_rn(hi, lo);
    return negative ? -hi : hi;
}

// ---------------------------------------------------------------------------------------------------------
// Per-thread floating point expansion of NFP doubles.  NFP = 3 absorbs ~2^100 of dynamic range without touching the
// shared accumulator (standalone dots, which are memory bound anyway); NFP = 2 costs 12 instead of 18 FP64 adds per
// element and still never spills for data whose products span less than ~2^40 -- the choice for dots fused into
// FP64-pipe-bound kernels.  Either way the result is exact: what the expansion cannot hold goes to the accumulator.
template <int NFP>
struct FpeT {
    double a[NFP];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NFP; i++) a[i] = 0.0;
    }
    // add x; a residue the expansion cannot hold is spilled (exactly) to `acc` (shared, stride 1).  The expansion
    // itself stays valid: after the cascade a[] + residue equals the old a[] + x exactly.
    __device__ __forceinline__ void add(double x, long long* acc) {
#pragma unroll
        for (int i = 0; i < NFP; i++) {
            double s;
            a[i] = two_sum(a[i], x, s);
            x = s;
        }
        if (x != 0.0) accumulate(acc, x, 1);
    }
    // the same without the spill: returns the residue (almost always 0.0).  Lets a caller run several independent
    // expansions through the FP64 pipe interleaved and test all residues with ONE branch (the cascade of one add is a
    // chain of 6 NFP dependent operations; a branch after every element would serialise the chains).
    __device__ __forceinline__ double add_lazy(double x) {
#pragma unroll
        for (int i = 0; i < NFP; i++) {
            double s;
            a[i] = two_sum(a[i], x, s);
            x = s;
        }
        return x;
    }
    // fold another expansion into this one (exact)
    __device__ __forceinline__ void merge(FpeT& o, long long* acc) {
#pragma unroll
        for (int i = 0; i < NFP; i++) { add(o.a[i], acc); o.a[i] = 0.0; }
    }
    __device__ __forceinline__ void flush(long long* acc) {
#pragma unroll
        for (int i = 0; i < NFP; i++) {
            accumulate(acc, a[i], 1);
            a[i] = 0.0;
        }
    }
    // exact warp-level reduction by shuffles, then lane 0 flushes: 3 instead of 96 accumulator updates per warp.
    // All 32 lanes must call.  (A lane that has handed its expansion to a partner adds nothing afterwards, so every
    // value -- including residues spilled on the way -- reaches the accumulator exactly once.)
    __device__ __forceinline__ void flush_warp(long long* acc) {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double v[NFP];
#pragma unroll
            for (int i = 0; i < NFP; i++) v[i] = __shfl_down_sync(0xffffffffu, a[i], off);
            if (lane < off) {
#pragma unroll
                for (int i = 0; i < NFP; i++)
                    if (v[i] != 0.0) add(v[i], acc);  // the tail components are usually zero: skip their add cascades
            }
        }
        if (lane == 0) flush(acc);
        else clear();
        __syncwarp();
    }
};
using Fpe = FpeT<NF>;

// flush_warp for two expansions at once: the two shuffle trees are independent, so their add cascades interleave in the
// FP64 pipe and the serial tail of a kernel that carries two dots is about as long as for one
template <int NFP>
__device__ __forceinline__ void flush_warp2(FpeT<NFP>& p, long long* pacc, FpeT<NFP>& q, long long* qacc) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        double v[NFP], w[NFP], rv[NFP], rw[NFP];
#pragma unroll
        for (int i = 0; i < NFP; i++) {
            v[i] = __shfl_down_sync(0xffffffffu, p.a[i], off);
            w[i] = __shfl_down_sync(0xffffffffu, q.a[i], off);
        }
        if (lane < off) {
            bool spill = false;
#pragma unroll
            for (int i = 0; i < NFP; i++) {
                rv[i] = p.add_lazy(v[i]);
                rw[i] = q.add_lazy(w[i]);
                spill = spill || rv[i] != 0.0 || rw[i] != 0.0;
            }
            if (spill) {
#pragma unroll 1
                for (int i = 0; i < NFP; i++) { accumulate(pacc, rv[i], 1); accumulate(qacc, rw[i], 1); }
            }
        }
    }
    if (lane == 0) { p.flush(pacc); q.flush(qacc); }
    else { p.clear(); q.clear(); }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------------
// Block- and grid-level reduction used by every kernel that carries a fused dot.
// Shared memory: NWARPS accumulators of BINS words (stride 1, accumulator w at smem + w*BINS).
// Global scratch ("slot"): partials[gridDim.x * BINS], status flags, a ticket; result record.
struct DotSlot {
    long long* gacc;          // [nslots][SPREAD][BINS] global accumulators, zero between launches
    int* gstatus;             // [nslots] OR of the per-block status flags, zero between launches
    unsigned int* ticket;     // [nslots] zero-initialised, reset by the finishing block
    dgb_dot_result* result;   // device
};

template <int NWARPS>
__device__ inline void block_init(long long* smem) {
    for (int i = threadIdx.x; i < NWARPS * BINS; i += blockDim.x) smem[i] = 0;
    __syncthreads();
}

// combines the per-warp accumulators of this block, adds the block partial into the global accumulator of the slot
// (wrap-safe 64-bit atomics: integer addition is associative => any grid size / arrival order gives the same VALUE)
// and lets the last block to arrive normalise, round and publish {acc, value, status}.
// `status` = 1 if this thread saw a non-finite product.  All threads of the block must call.  Returns true in ALL
// threads of the finishing block after the result is written (so callers can chain scalar post-processing).
// the scalar tail of a reduction, kept out of line so that its registers (39-word loops, ldexp, ...) do not count
// against the streaming loop of the calling kernel
static __device__ __noinline__ void normalize_noinline(long long* acc) { normalize(acc, 1); }
static __device__ __noinline__ void publish_result(long long* smem, const DotSlot& slot, int slot_idx) {
    int negative = normalize(smem, 1);  // after prenormalize() below the carry chain is short but must still run in order
    dgb_dot_result* r = slot.result + slot_idx;
    r->value = round_normalized(smem, negative);
    r->status = __ldcg(slot.gstatus + slot_idx);
    r->pad = 0;
    slot.gstatus[slot_idx] = 0;
    slot.ticket[slot_idx] = 0;
    __threadfence();
}

template <int NWARPS>
__device__ inline bool block_finish(long long* smem, int status, const DotSlot& slot, int slot_idx = 0) {
    __shared__ int s_last;
    int any_bad = __syncthreads_or(status);
    long long sum = 0;
    if (NWARPS == 1) {
        // one accumulator per block (the usual case: expansions are reduced by shuffles first, see flush_warp)
        if (threadIdx.x < BINS) sum = smem[threadIdx.x];
    } else {
        // 1. normalise each warp accumulator (one thread each), then sum word-wise (NWARPS <= 32 -> no overflow)
        if (threadIdx.x < NWARPS) normalize_noinline(smem + threadIdx.x * BINS);
        __syncthreads();
        if (threadIdx.x < BINS) {
#pragma unroll
            for (int w = 0; w < NWARPS; w++) sum += smem[w * BINS + threadIdx.x];
        }
    }
    __syncthreads();
    long long* gbase = slot.gacc + (size_t)slot_idx * GACC_WORDS;
    long long* gacc = gbase + (blockIdx.x % SPREAD) * BINS;
    // 2. word-parallel atomic accumulation into one of the SPREAD global accumulators of the slot
    if (threadIdx.x < BINS && sum != 0) add_word(gacc, threadIdx.x, sum, 1);
    if (threadIdx.x == 0 && any_bad) atomicOr(slot.gstatus + slot_idx, 1);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(slot.ticket + slot_idx, 1u);
        s_last = (t == gridDim.x * gridDim.y * gridDim.z - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    // 3. last block: fetch, reset, normalise, round
    __threadfence();
    // the copies hold arbitrary int64 digits: add low 56 bits and signed high parts separately (no overflow), the
    // high parts carry into the next word
    __shared__ long long s_hi[BINS];
    if (threadIdx.x < BINS) {
        long long lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < SPREAD; k++) {
            const long long v = __ldcg(gbase + k * BINS + threadIdx.x);
            gbase[k * BINS + threadIdx.x] = 0;  // ready for the next launch on the same stream
            const long long h = v >> DIGITS;
            hi += h;
            lo += v - (long long)((unsigned long long)h << DIGITS);
        }
        if (threadIdx.x == BINS - 1) { lo += (long long)((unsigned long long)hi << DIGITS); hi = 0; }  // top word keeps its sign
        smem[threadIdx.x] = lo;
        s_hi[threadIdx.x] = hi;
    }
    __syncthreads();
    if (threadIdx.x > 0 && threadIdx.x < BINS) smem[threadIdx.x] += s_hi[threadIdx.x - 1];
    __syncthreads();
    if (threadIdx.x == 0) publish_result(smem, slot, slot_idx);
    __syncthreads();
    if (threadIdx.x < BINS) slot.result[slot_idx].acc[threadIdx.x] = smem[threadIdx.x];  // the normalised words
    __threadfence();
    __syncthreads();
    return true;
}

// K dots of one kernel finished together: accumulators smem[k * BINS ..], slots slot0 .. slot0 + K - 1, ONE fence /
// ticket sequence (the ticket of slot0).  K <= 4, blockDim.x >= 64 K.  Thread group k = threads [64 k, 64 k + 39).
template <int K>
__device__ inline bool block_finish_multi(long long* smem, int status, const DotSlot& slot, int slot0) {
    __shared__ int s_last;
    __shared__ long long s_hi[K * BINS];
    const int any_bad = __syncthreads_or(status);
    const int grp = threadIdx.x >> 6, t = threadIdx.x & 63;
    if (grp < K && t < BINS) {
        const long long v = smem[grp * BINS + t];
        if (v != 0) add_word(slot.gacc + (size_t)(slot0 + grp) * GACC_WORDS + (blockIdx.x % SPREAD) * BINS, t, v, 1);
    }
    if (threadIdx.x < K && any_bad) atomicOr(slot.gstatus + slot0 + threadIdx.x, 1);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tk = atomicAdd(slot.ticket + slot0, 1u);
        s_last = (tk == gridDim.x * gridDim.y * gridDim.z - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (grp < K && t < BINS) {
        long long* gbase = slot.gacc + (size_t)(slot0 + grp) * GACC_WORDS;
        long long lo = 0, hi = 0;
#pragma unroll
        for (int k = 0; k < SPREAD; k++) {
            const long long v = __ldcg(gbase + k * BINS + t);
            gbase[k * BINS + t] = 0;
            const long long h = v >> DIGITS;
            hi += h;
            lo += v - (long long)((unsigned long long)h << DIGITS);
        }
        if (t == BINS - 1) { lo += (long long)((unsigned long long)hi << DIGITS); hi = 0; }
        smem[grp * BINS + t] = lo;
        s_hi[grp * BINS + t] = hi;
    }
    __syncthreads();
    if (grp < K && t > 0 && t < BINS) smem[grp * BINS + t] += s_hi[grp * BINS + t - 1];
    __syncthreads();
    if (grp < K && t == 0) publish_result(smem + grp * BINS, slot, slot0 + grp);
    __syncthreads();
    if (grp < K && t < BINS) slot.result[slot0 + grp].acc[t] = smem[grp * BINS + t];
    __threadfence();
    __syncthreads();
    return true;
}

}  // namespace sa
}  // namespace dgb
