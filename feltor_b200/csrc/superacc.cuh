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
#ifndef DGB_TRACE  // tools/scratch/tail_bench.cu defines it to time the phases of the reduction tail
#define DGB_TRACE(k)
#endif

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
    // ldexp(x, -56 exp_word) as two exact multiplications by 2^(-28 exp_word) (|28 exp_word| <= 560: a normal double, and
    // the intermediate stays normal for every finite x including subnormals, so both products are exact)
    const double half_scale = __hiloint2double((1023 - (DIGITS / 2) * exp_word) << 20, 0);
    double xscaled = __dmul_rn(__dmul_rn(x, half_scale), half_scale);
    for (int i = iup; i >= 0 && xscaled != 0.0; --i) {
        double xr = rint(xscaled);
        long long xi = __double2ll_rn(xscaled);
        add_word(acc, i, xi, stride);
        xscaled = __dsub_rn(xscaled, xr);
        xscaled = __dmul_rn(xscaled, 72057594037927936.0);  // 2^56
    }
}

// The same for the reduction tails: the (at most three) digits of x are computed first, then their atomics are issued
// back to back -- in accumulate() every atomic waits for the previous one's return value (~250 cycles each in shared
// memory), which is what a flush of three components by one lane used to cost.  A double has 53 significant bits: two
// 56-bit digits cover it when the scaled value is >= 1, three when it is a fraction (negative exponents truncate toward
// zero); anything left after three digits (cannot happen for finite input) takes the general loop.
static __device__ __noinline__ void accumulate3(long long* acc, double x) {
    if (x == 0.0) return;
    const int e = ((int)((unsigned long long)__double_as_longlong(x) >> 52) & 0x7ff) - 0x3ff;
    const int exp_word = e / DIGITS;
    const int iup = exp_word + F_WORDS;
    const double half_scale = __hiloint2double((1023 - (DIGITS / 2) * exp_word) << 20, 0);
    double xscaled = __dmul_rn(__dmul_rn(x, half_scale), half_scale);
    long long d[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        d[k] = 0;
        if (iup - k >= 0 && xscaled != 0.0) {
            const double xr = rint(xscaled);
            d[k] = __double2ll_rn(xscaled);
            xscaled = __dmul_rn(__dsub_rn(xscaled, xr), 72057594037927936.0);
        }
    }
    unsigned long long old[3];
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (d[k] != 0) old[k] = atomicAdd(reinterpret_cast<unsigned long long*>(acc + (iup - k)), (unsigned long long)d[k]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (d[k] == 0) continue;
        const long long o = (long long)old[k], nw = (long long)(old[k] + (unsigned long long)d[k]);
        if (((o ^ nw) & (d[k] ^ nw)) < 0) add_word(acc, iup - k + 1, (o > 0 ? 1ll : -1ll) * (1ll << KRX), 1);  // wrapped word: see add_word
    }
    for (int i = iup - 3; i >= 0 && xscaled != 0.0; --i) {
        const double xr = rint(xscaled);
        add_word(acc, i, __double2ll_rn(xscaled), 1);
        xscaled = __dmul_rn(__dsub_rn(xscaled, xr), 72057594037927936.0);
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

// The same carry propagation by ONE WARP: lane l holds words 2l and 2l+1, splits them locally and hands the carry of its
// upper word to lane l+1 by shuffle; repeated until no word is out of range.  The canonical form (every word but the top
// one in [0, 2^56)) is unique, so the result equals the sequential loop's.  Typically 2-3 rounds instead of a 39-step
// dependent chain through shared memory; a borrow rippling through zero words costs one round per two words.
// All 32 lanes of the warp must call; acc is shared memory (stride 1); returns 1 in all lanes if the value is negative.
__device__ inline int normalize_warp(long long* acc) {
    const int lane = threadIdx.x & 31;
    const int i0 = 2 * lane, i1 = 2 * lane + 1;
    long long w0 = i0 < BINS ? acc[i0] : 0, w1 = i1 < BINS ? acc[i1] : 0;
    const bool split0 = i0 < BINS - 1, split1 = i1 < BINS - 1;  // the top word keeps its carry (sign)
    for (;;) {
        long long c = 0;
        if (split0) { c = w0 >> DIGITS; w0 -= (long long)((unsigned long long)c << DIGITS); w1 += c; }
        c = 0;
        if (split1) { c = w1 >> DIGITS; w1 -= (long long)((unsigned long long)c << DIGITS); }
        long long cin = __shfl_up_sync(0xffffffffu, c, 1);
        if (lane == 0) cin = 0;
        w0 += cin;
        // a received carry can push w0 out of range again (and with it w1 in the next round)
        const bool again = split0 && (w0 >> DIGITS) != 0;
        if (!__any_sync(0xffffffffu, again)) break;
    }
    if (i0 < BINS) acc[i0] = w0;
    if (i1 < BINS) acc[i1] = w1;
    const long long top = __shfl_sync(0xffffffffu, w0, (BINS - 1) / 2);  // word 38 = lane 19, w0
    __syncwarp();
    return top < 0;
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

// The same by one warp: the leading word and the sticky bit are found with ballots instead of two dependent walks through
// shared memory (39 + 37 loads in a row); the arithmetic on the three words involved is the function above, word for word.
// All 32 lanes must call with the same arguments; every lane returns the value.
__device__ inline double round_normalized_warp(const long long* acc, int negative) {
    const long long MASK = (1ll << DIGITS) - 1;
    const int lane = threadIdx.x & 31;
    const long long wlo = acc[lane], whi = lane + 32 < BINS ? acc[lane + 32] : 0;
    // bit w of `nz`: word w is non-zero; of `nf`: its low 56 bits are not all ones; of `st`: it contributes to the sticky bit
    const unsigned long long nz = (unsigned long long)__ballot_sync(0xffffffffu, wlo != 0) |
                                  ((unsigned long long)(__ballot_sync(0xffffffffu, whi != 0) & 0x7fu) << 32);
    if (nz == 0) return 0.0;
    int i = 63 - __clzll((long long)nz);
    if (negative) {
        const unsigned long long nf = (unsigned long long)__ballot_sync(0xffffffffu, (wlo & MASK) != MASK) |
                                      ((unsigned long long)(__ballot_sync(0xffffffffu, lane + 32 < BINS && (whi & MASK) != MASK) & 0x7fu) << 32);
        const unsigned long long below = nf & (i == 63 ? ~0ull : ((1ull << (i + 1)) - 1ull));
        if (below == 0) return 0.0;
        i = 63 - __clzll((long long)below);
    }
    long long hiword = negative ? MASK - acc[i] : acc[i];
    double rounded = __ll2double_rn(hiword);
    double hi = ldexp(rounded, (i - F_WORDS) * DIGITS);
    if (i == 0) return negative ? -hi : hi;
    hiword -= __double2ll_rn(rounded);
    double mid = ldexp(__ll2double_rn(hiword), (i - F_WORDS) * DIGITS);
    // sticky: OR of the terms of words 0 .. i-2 (a negative value contributes 2^56 - word, which is never zero)
    const unsigned long long st = (unsigned long long)__ballot_sync(0xffffffffu, negative ? true : wlo != 0) |
                                  ((unsigned long long)(__ballot_sync(0xffffffffu, negative ? true : whi != 0) & 0x7fu) << 32);
    const bool sticky = i >= 2 && (st & ((1ull << (i - 1)) - 1ull)) != 0;
    long long loword = negative ? ((1ll << DIGITS) - acc[i - 1]) : acc[i - 1];
    loword |= (long long)sticky;
    double lo = ldexp(__ll2double_rn(loword), (i - 1 - F_WORDS) * DIGITS);
    if (mid != 0.0) {
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
    // fold the expansion v[] into this one (exact) through a SHALLOW network of error-free TwoSums: component-wise sums
    // first (independent), then the errors move down one level each.  Every TwoSum is exact, so a[] plus the residues
    // that fall off the end equals the old a[] + v[] exactly; the residues (non-zero only when the operands span more
    // than NFP doubles of dynamic range) go to the accumulator `acc`.  Critical path: NFP TwoSums instead of NFP^2 for
    // the cascade -- FP64 adds have ~20 cycles of dependent-issue latency on B200 and the reduction tails run with one
    // warp per scheduler, so depth is what they cost.
    __device__ __forceinline__ void absorb(const double (&v)[NFP], long long* acc) {
        double s[NFP], e[NFP];
#pragma unroll
        for (int i = 0; i < NFP; i++) s[i] = two_sum(a[i], v[i], e[i]);
        double spill[NFP];
        spill[NFP - 1] = e[NFP - 1];
        // level k: the errors of level k-1 join the next lower component
#pragma unroll
        for (int lvl = 1; lvl < NFP; lvl++) {
            double ne[NFP];
#pragma unroll
            for (int i = NFP - 1; i >= lvl; i--) s[i] = two_sum(s[i], e[i - 1], ne[i]);
            spill[NFP - 1 - lvl] = ne[NFP - 1];
#pragma unroll
            for (int i = lvl; i < NFP - 1; i++) e[i] = ne[i];
        }
#pragma unroll
        for (int i = 0; i < NFP; i++) a[i] = s[i];
        bool any = false;
#pragma unroll
        for (int i = 0; i < NFP; i++) any = any || spill[i] != 0.0;
        if (any) {
#pragma unroll 1
            for (int i = 0; i < NFP; i++) accumulate(acc, spill[i], 1);
        }
    }
    // fold another expansion into this one (exact)
    __device__ __forceinline__ void merge(FpeT& o, long long* acc) {
        absorb(o.a, acc);
#pragma unroll
        for (int i = 0; i < NFP; i++) o.a[i] = 0.0;
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
            if (lane < off) absorb(v, acc);
        }
        // lane 0 holds the warp's sum: its NFP components go to the accumulator from NFP lanes at once
        double comp = 0.0;
#pragma unroll
        for (int i = 0; i < NFP; i++) {
            const double vi = __shfl_sync(0xffffffffu, a[i], 0);
            if (lane == i) comp = vi;
        }
        if (lane < NFP) accumulate3(acc, comp);
        clear();
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
        double v[NFP], w[NFP];
#pragma unroll
        for (int i = 0; i < NFP; i++) {
            v[i] = __shfl_down_sync(0xffffffffu, p.a[i], off);
            w[i] = __shfl_down_sync(0xffffffffu, q.a[i], off);
        }
        if (lane < off) { p.absorb(v, pacc); q.absorb(w, qacc); }  // two independent networks: they interleave
    }
    double comp = 0.0;
#pragma unroll
    for (int i = 0; i < NFP; i++) {
        const double vp = __shfl_sync(0xffffffffu, p.a[i], 0), vq = __shfl_sync(0xffffffffu, q.a[i], 0);
        if (lane == i) comp = vp;
        if (lane == NFP + i) comp = vq;
    }
    if (lane < 2 * NFP) accumulate3(lane < NFP ? pacc : qacc, comp);
    p.clear();
    q.clear();
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
// one warp: normalise the summed accumulator in `smem` (warp-parallel carry propagation), round, publish value / status
// and re-arm the slot for the next launch.  The normalised words stay in smem for the caller to copy out.
static __device__ __noinline__ void publish_result(long long* smem, const DotSlot& slot, int slot_idx) {
    const int negative = normalize_warp(smem);
    const double value = round_normalized_warp(smem, negative);
    if ((threadIdx.x & 31) == 0) {
        dgb_dot_result* r = slot.result + slot_idx;
        r->value = value;
        r->status = __ldcg(slot.gstatus + slot_idx);
        r->pad = 0;
        slot.gstatus[slot_idx] = 0;
        slot.ticket[slot_idx] = 0;
    }
    __syncwarp();
}

// the ticket: one acquire-release read-modify-write at gpu scope (release: this block's accumulator atomics, ordered before
// it by the preceding bar.sync, become visible before the count; acquire: the last block sees every other block's)
__device__ __forceinline__ unsigned int ticket_take(unsigned int* ticket) {
    unsigned int t;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    return t;
}

// Ordering of the grid-level hand-over (both finish functions): the word atomics of a block are ordered before its
// ticket by  [atomics] -> bar.sync -> acq_rel ticket RMW by thread 0 (release is cumulative over the barrier);  the last
// block reads the words after  ticket RMW (acquire) -> bar.sync -> loads.  One thread orders instead of a __threadfence by
// all of them (a fence by every thread waits for every thread's outstanding vector stores).
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
    DGB_TRACE(0);
    if (threadIdx.x < BINS && sum != 0) add_word(gacc, threadIdx.x, sum, 1);
    if (threadIdx.x == 0 && any_bad) atomicOr(slot.gstatus + slot_idx, 1);
    __syncthreads();
    DGB_TRACE(1);
    if (threadIdx.x == 0) s_last = ticket_take(slot.ticket + slot_idx) == gridDim.x * gridDim.y * gridDim.z - 1;
    __syncthreads();
    DGB_TRACE(2);
    if (!s_last) return false;
    // 3. last block: fetch, reset, normalise, round
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
    DGB_TRACE(3);
    if (threadIdx.x < 32) publish_result(smem, slot, slot_idx);
    __syncthreads();
    DGB_TRACE(4);
    if (threadIdx.x < BINS) slot.result[slot_idx].acc[threadIdx.x] = smem[threadIdx.x];  // the normalised words
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
    __syncthreads();
    if (threadIdx.x == 0) s_last = ticket_take(slot.ticket + slot0) == gridDim.x * gridDim.y * gridDim.z - 1;
    __syncthreads();
    if (!s_last) return false;
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
    if (grp < K && t < 32) publish_result(smem + grp * BINS, slot, slot0 + grp);  // the first warp of each group
    __syncthreads();
    if (grp < K && t < BINS) slot.result[slot0 + grp].acc[t] = smem[grp * BINS + t];
    __syncthreads();
    return true;
}

}  // namespace sa
}  // namespace dgb
