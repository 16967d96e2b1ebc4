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
// Per-thread floating point expansion
struct Fpe {
    double a[NF];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < NF; i++) a[i] = 0.0;
    }
    // add x; a residue the expansion cannot hold is spilled (exactly) to `acc` (shared, stride 1).  The expansion
    // itself stays valid: after the cascade a[] + residue equals the old a[] + x exactly.
    __device__ __forceinline__ void add(double x, long long* acc) {
#pragma unroll
        for (int i = 0; i < NF; i++) {
            double s;
            a[i] = two_sum(a[i], x, s);
            x = s;
        }
        if (x != 0.0) accumulate(acc, x, 1);
    }
    __device__ __forceinline__ void flush(long long* acc) {
#pragma unroll
        for (int i = 0; i < NF; i++) {
            accumulate(acc, a[i], 1);
            a[i] = 0.0;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// Block- and grid-level reduction used by every kernel that carries a fused dot.
// Shared memory: NWARPS accumulators of BINS words (stride 1, accumulator w at smem + w*BINS).
// Global scratch ("slot"): partials[gridDim.x * BINS], status flags, a ticket; result record.
struct DotSlot {
    long long* gacc;          // [nslots][BINS] global accumulators, zero between launches
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
template <int NWARPS>
__device__ inline bool block_finish(long long* smem, int status, const DotSlot& slot, int slot_idx = 0) {
    __shared__ int s_last;
    int any_bad = __syncthreads_or(status);
    // 1. normalise each warp accumulator (one thread each), then sum word-wise (NWARPS <= 32 -> no overflow)
    if (threadIdx.x < NWARPS) normalize(smem + threadIdx.x * BINS, 1);
    __syncthreads();
    long long sum = 0;
    if (threadIdx.x < BINS) {
#pragma unroll
        for (int w = 0; w < NWARPS; w++) sum += smem[w * BINS + threadIdx.x];
    }
    __syncthreads();
    long long* gacc = slot.gacc + (size_t)slot_idx * BINS;
    // 2. word-parallel atomic accumulation into the global accumulator
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
    if (threadIdx.x < BINS) {
        smem[threadIdx.x] = __ldcg(gacc + threadIdx.x);
        gacc[threadIdx.x] = 0;  // ready for the next launch on the same stream
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int negative = normalize(smem, 1);
        dgb_dot_result* r = slot.result + slot_idx;
        for (int i = 0; i < BINS; i++) r->acc[i] = smem[i];
        r->value = round_normalized(smem, negative);
        r->status = __ldcg(slot.gstatus + slot_idx);
        r->pad = 0;
        slot.gstatus[slot_idx] = 0;
        slot.ticket[slot_idx] = 0;
        __threadfence();
    }
    __syncthreads();
    return true;
}

}  // namespace sa
}  // namespace dgb
