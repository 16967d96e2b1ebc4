// Internal: communicator for the y-decomposed (slab) solver: NCCL for bulk data, peer memory (CUDA IPC over
// NVLink/NVSwitch) for the 39-word exact-dot exchange.
#pragma once
#include "common.cuh"
namespace dgb {
struct Comm;
int comm_rank(const Comm* c);
int comm_size(const Comm* c);
// in-place sum over ranks of `count` int64 words on `st`
int comm_allreduce_i64(Comm* c, long long* buf, size_t count, cudaStream_t st);
// exchange `ghost_rows` rows of `row_len` doubles with the lower / upper neighbour of a ring (periodic) or chain:
// interior = first interior row of a buffer laid out [ghost_rows | nrows | ghost_rows]
int comm_halo_rows(Comm* c, double* interior, size_t row_len, size_t nrows, size_t ghost_rows, int periodic, cudaStream_t st);

// ---- peer-memory exchange of superaccumulator records ------------------------------------------------------
// Every rank owns an exchange buffer that all peers map (cudaIpcOpenMemHandle).  A global dot is ONE small kernel:
// store my normalised record into slot [channel][epoch parity][my rank] of every peer, each word together with the
// epoch; poll the slots of all ranks in MY buffer until they carry this epoch; add the records word by word
// (integers: associative, identical on every rank).  No collective library call, no host involvement; latency = one NVLink round trip.
constexpr int P2P_MAX_RANKS = 8, P2P_CHANNELS = 8;
// "LL" layout: every 8-byte payload word travels in ONE 16-byte store together with its epoch, so the receiver needs no
// fence -- it polls each element until the epoch matches (a 16-byte aligned vector store is observed atomically).
constexpr int P2P_REC_ELEMS = 64;  // 39 accumulator words + status, padded: 64 x 16 B = 1 KiB per (channel, parity, rank)
constexpr size_t P2P_BYTES = (size_t)P2P_CHANNELS * 2 * P2P_MAX_RANKS * P2P_REC_ELEMS * 16;
struct P2pView {
    long long* peer[P2P_MAX_RANKS];  // peer[r] = exchange buffer of rank r as mapped into this process
    int rank, size, enabled;
};
// the view and the next epoch of `count` consecutive channels starting at `first` (epochs advance by one per call)
P2pView comm_p2p_view(Comm* c);
// COLLECTIVE: map the cudaMalloc'ed buffer `base` of every rank into this process; peers[r] = rank r's buffer (own rank:
// base itself).  *mapped = 1 and peers filled if every rank mapped every buffer, *mapped = 0 (peers cleared) if the
// peer-memory path is disabled or any rank failed -- all ranks take the same decision: local failures (IPC handle, NCCL,
// CUDA) travel as a flag through the closing allreduce, which every rank reaches.  Returns a DGB_ERR_* code only for
// failures of that agreement itself.  Mapped buffers stay open until comm_p2p_unmap / the communicator is destroyed.
int comm_p2p_map(Comm* c, void* base, void** peers, int* mapped);
void comm_p2p_unmap(Comm* c, void** peers);
// neighbour barrier of the slab ring/chain on channel 7: returns when `lower` and `upper` (rank ids, -1 = none) have
// passed the same point of their streams, i.e. their preceding kernels (with their peer stores) are complete
int comm_p2p_neighbour_barrier(Comm* c, int lower, int upper, cudaStream_t st);
unsigned long long comm_p2p_next_epoch(Comm* c, int first, int count);

#ifdef __CUDACC__
__device__ __forceinline__ ulonglong2* p2p_rec(long long* base, int channel, unsigned long long epoch, int src) {
    return reinterpret_cast<ulonglong2*>(base) + (((size_t)channel * 2 + (epoch & 1ull)) * P2P_MAX_RANKS + src) * P2P_REC_ELEMS;
}
__device__ __forceinline__ void p2p_store(ulonglong2* p, unsigned long long data, unsigned long long epoch) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(data), "l"(epoch) : "memory");
}
__device__ __forceinline__ unsigned long long p2p_wait(const ulonglong2* p, unsigned long long epoch) {
    unsigned long long d, e;
    do {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(d), "=l"(e) : "l"(p) : "memory");
    } while (e != epoch);
    return d;
}
// All threads of a block of >= 64 threads call.  recs[first .. first+count) hold this rank's normalised records
// (dgb_dot_result layout: 39 words, value, status|pad); on return they hold the word-wise sums over ranks (not
// normalised) with status = number of ranks that met NaN/Inf.  The epochs of the channels first.. are `epoch`.
__device__ inline void p2p_allreduce_records(const P2pView& v, long long* recs, int first, int count, unsigned long long epoch) {
    const int t = threadIdx.x;
    if (t < 40) {
        for (int k = 0; k < count; k++) {
            long long* mine = recs + (size_t)(first + k) * 41;
            const long long w = t < 39 ? mine[t] : (long long)((reinterpret_cast<int*>(mine + 40)[0] != 0) || (reinterpret_cast<int*>(mine + 40)[1] != 0));
            for (int q = 0; q < v.size; q++) p2p_store(p2p_rec(v.peer[q], first + k, epoch, v.rank) + t, (unsigned long long)w, epoch);
        }
        for (int k = 0; k < count; k++) {
            long long* mine = recs + (size_t)(first + k) * 41;
            long long sum = 0;
            for (int r = 0; r < v.size; r++) sum += (long long)p2p_wait(p2p_rec(v.peer[v.rank], first + k, epoch, r) + t, epoch);
            if (t < 39) mine[t] = sum;
            else { reinterpret_cast<int*>(mine + 40)[0] = (int)sum; reinterpret_cast<int*>(mine + 40)[1] = 0; }
        }
    }
    __syncthreads();
}
#endif
}  // namespace dgb
