// Multi-GPU plumbing: one process per GPU, NCCL over NVLink/NVSwitch.  Replaces the MPI layer of the reference on
// this path: MPIKroneckerGather / MPIContiguousGather halo exchange (inc/dg/backend/mpi_gather_kron.h:143-291,
// mpi_gather.h:255-440: MPI_Isend/Irecv + cudaDeviceSynchronize per exchange) by one grouped ncclSend/ncclRecv
// enqueued on the compute stream (no host synchronisation), and exblas::reduce_mpi_cpu
// (inc/dg/backend/exblas/mpi_accumulate.h:42-125: D2H, MPI_Reduce on 39 longs, MPI_Bcast) by an in-place
// ncclAllReduce(int64, sum) of the normalised superaccumulator words on the device -- integer sums are associative,
// so the global dot is bit-reproducible for any number of GPUs.
// NCCL is bound at run time (dlopen) so that single-GPU users do not need it.
#include "comm.cuh"
#include "superacc.cuh"
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>
#include <vector>

namespace dgb {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt64 = 4, ncclFloat64 = 8 };  // nccl.h ncclDataType_t
enum { ncclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) {
#define BIND(name) *(void**)(&api.name) = dlsym(api.lib, "nccl" #name)
            BIND(GetUniqueId); BIND(CommInitRank); BIND(CommDestroy); BIND(AllReduce); BIND(AllGather); BIND(Send); BIND(Recv);
            BIND(GroupStart); BIND(GroupEnd); BIND(GetErrorString);
#undef BIND
        }
    }
    if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Send || !api.Recv || !api.GroupStart ||
        !api.GroupEnd)
        return nullptr;
    return &api;
}
#define DGB_NCCL(call)                                                                          \
    do {                                                                                        \
        ncclResult_t _r = (call);                                                               \
        if (_r != 0) {                                                                          \
            set_error("NCCL error %d (%s) in %s", _r, nccl()->GetErrorString ? nccl()->GetErrorString(_r) : "?", #call); \
            return DGB_ERR_INVALID;                                                             \
        }                                                                                       \
    } while (0)

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1;
    // peer-memory exchange (comm.cuh)
    bool p2p = false;
    long long* local = nullptr;
    long long* peer[P2P_MAX_RANKS] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    unsigned long long epoch[P2P_CHANNELS] = {0, 0, 0, 0, 0, 0, 0, 0};
};
int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_size(const Comm* c) { return c ? c->size : 1; }

P2pView comm_p2p_view(Comm* c) {
    P2pView v;
    for (int r = 0; r < P2P_MAX_RANKS; r++) v.peer[r] = c ? c->peer[r] : nullptr;
    v.rank = c ? c->rank : 0;
    v.size = c ? c->size : 1;
    v.enabled = c && c->p2p ? 1 : 0;
    return v;
}
unsigned long long comm_p2p_next_epoch(Comm* c, int first, int count) {
    // all channels of one call advance together; a channel that sat out some calls catches up to the maximum so that
    // every channel of the call carries the SAME epoch on every rank (ranks issue identical call sequences)
    unsigned long long e = 0;
    for (int k = first; k < first + count; k++) e = c->epoch[k] > e ? c->epoch[k] : e;
    e += 1;
    for (int k = first; k < first + count; k++) c->epoch[k] = e;
    return e;
}

// map every rank's exchange buffer into this process (CUDA IPC; the handles travel through an ncclAllGather)
static int comm_p2p_setup(Comm* c) {
    const char* off = getenv("DGB_NO_P2P");
    if ((off && atoi(off)) || c->size > P2P_MAX_RANKS || c->size < 2 || !nccl()->AllGather) return 0;
    DGB_CUDA(cudaMalloc(&c->local, P2P_BYTES));
    DGB_CUDA(cudaMemset(c->local, 0, P2P_BYTES));
    DGB_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine;
    cudaError_t ce = cudaIpcGetMemHandle(&mine, c->local);
    // every rank must take part in the gather even if its own handle could not be made: flag it with zeros
    unsigned char* dev = nullptr;
    const size_t HB = sizeof(cudaIpcMemHandle_t) + 8;
    DGB_CUDA(cudaMalloc(&dev, HB * c->size));
    std::vector<unsigned char> host(HB * c->size, 0);
    if (ce == cudaSuccess) { memcpy(host.data() + HB * c->rank, &mine, sizeof(mine)); host[HB * c->rank + sizeof(mine)] = 1; }
    else cudaGetLastError();
    DGB_CUDA(cudaMemcpy(dev + HB * c->rank, host.data() + HB * c->rank, HB, cudaMemcpyHostToDevice));
    DGB_NCCL(nccl()->AllGather(dev + HB * c->rank, dev, HB, 0 /* ncclInt8 */, c->comm, nullptr));
    DGB_CUDA(cudaDeviceSynchronize());
    DGB_CUDA(cudaMemcpy(host.data(), dev, HB * c->size, cudaMemcpyDeviceToHost));
    cudaFree(dev);
    bool ok = true;
    for (int r = 0; r < c->size; r++) ok = ok && host[HB * r + sizeof(mine)] == 1;
    for (int r = 0; r < c->size && ok; r++) {
        if (r == c->rank) { c->peer[r] = c->local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, host.data() + HB * r, sizeof(h));
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        c->peer[r] = reinterpret_cast<long long*>(p);
    }
    // agree on the outcome (a rank that failed to map a peer must not leave the others spinning on its flags)
    long long* flag = nullptr;
    DGB_CUDA(cudaMalloc(&flag, 8));
    long long hv = ok ? 0 : 1;
    DGB_CUDA(cudaMemcpy(flag, &hv, 8, cudaMemcpyHostToDevice));
    DGB_NCCL(nccl()->AllReduce(flag, flag, 1, ncclInt64, ncclSum, c->comm, nullptr));
    DGB_CUDA(cudaDeviceSynchronize());
    DGB_CUDA(cudaMemcpy(&hv, flag, 8, cudaMemcpyDeviceToHost));
    cudaFree(flag);
    c->p2p = hv == 0;
    if (getenv("DGB_COMM_VERBOSE"))
        fprintf(stderr, "dgb_comm rank %d/%d: peer-memory dot exchange %s\n", c->rank, c->size, c->p2p ? "enabled" : "unavailable (NCCL allreduce)");
    return 0;
}

int comm_p2p_map(Comm* c, void* basep, void** peers, int* mapped) {
    *mapped = 0;
    if (!c || !c->p2p) return 0;  // decided collectively at communicator creation: no rank communicates here
    for (int r = 0; r < c->size; r++) peers[r] = nullptr;
    const size_t HB = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<unsigned char> host(HB * c->size, 0);
    unsigned char* dev = nullptr;
    bool ok = cudaMalloc(&dev, HB * c->size + 8) == cudaSuccess;  // [handles | agreement flag]
    int nccl_err = 0;
    if (ok) {
        cudaIpcMemHandle_t mine;
        if (cudaIpcGetMemHandle(&mine, basep) == cudaSuccess) { memcpy(host.data() + HB * c->rank, &mine, sizeof(mine)); host[HB * c->rank + sizeof(mine)] = 1; }
        ok = cudaMemcpy(dev + HB * c->rank, host.data() + HB * c->rank, HB, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    // a rank whose scratch allocation failed cannot take part in the collectives at all: that is the one fatal case
    if (!dev) { cudaGetLastError(); set_error("comm_p2p_map: cudaMalloc of the handle exchange buffer failed"); return (int)cudaErrorMemoryAllocation; }
    if (nccl()->AllGather(dev + HB * c->rank, dev, HB, 0 /* ncclInt8 */, c->comm, nullptr) != 0) { ok = false; nccl_err = 1; }
    ok = (cudaDeviceSynchronize() == cudaSuccess) && ok;
    ok = (cudaMemcpy(host.data(), dev, HB * c->size, cudaMemcpyDeviceToHost) == cudaSuccess) && ok;
    for (int r = 0; r < c->size; r++) ok = ok && host[HB * r + sizeof(cudaIpcMemHandle_t)] == 1;
    for (int r = 0; r < c->size && ok; r++) {
        if (r == c->rank) { peers[r] = basep; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, host.data() + HB * r, sizeof(h));
        if (cudaIpcOpenMemHandle(&peers[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { peers[r] = nullptr; ok = false; }
    }
    cudaGetLastError();  // local failures are not sticky errors; they are reported through the flag
    long long hv = ok ? 0 : 1;
    long long* flag = reinterpret_cast<long long*>(dev + HB * c->size);
    int e = 0;
    if (cudaMemcpy(flag, &hv, 8, cudaMemcpyHostToDevice) != cudaSuccess) e = (int)cudaErrorUnknown;
    if (!e && !nccl_err && nccl()->AllReduce(flag, flag, 1, ncclInt64, ncclSum, c->comm, nullptr) != 0) e = DGB_ERR_INVALID;
    if (!e && cudaDeviceSynchronize() != cudaSuccess) e = (int)cudaErrorUnknown;
    if (!e && cudaMemcpy(&hv, flag, 8, cudaMemcpyDeviceToHost) != cudaSuccess) e = (int)cudaErrorUnknown;
    cudaFree(dev);
    if (e || nccl_err) {
        comm_p2p_unmap(c, peers);
        set_error("comm_p2p_map: the agreement collective failed (%s)", nccl_err ? "ncclAllGather" : "allreduce / CUDA");
        return e ? e : DGB_ERR_INVALID;
    }
    if (hv != 0) { comm_p2p_unmap(c, peers); return 0; }
    *mapped = 1;
    return 0;
}
void comm_p2p_unmap(Comm* c, void** peers) {
    if (!c) return;
    for (int r = 0; r < c->size; r++) {
        if (r != c->rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
        peers[r] = nullptr;
    }
}

__global__ void __launch_bounds__(32) p2p_barrier_kernel(P2pView v, int lower, int upper, unsigned long long epoch) {
    const int t = threadIdx.x;
    const int who = t == 0 ? lower : (t == 1 ? upper : -1);
    if (who < 0 || (t == 1 && upper == lower)) return;
    p2p_store(p2p_rec(v.peer[who], 7, epoch, v.rank), 1ull, epoch);
    p2p_wait(p2p_rec(v.peer[v.rank], 7, epoch, who), epoch);
}
int comm_p2p_neighbour_barrier(Comm* c, int lower, int upper, cudaStream_t st) {
    if (!c || !c->p2p) { set_error("peer-memory barrier without peer memory"); return DGB_ERR_INVALID; }
    const unsigned long long epoch = comm_p2p_next_epoch(c, 7, 1);
    p2p_barrier_kernel<<<1, 32, 0, st>>>(comm_p2p_view(c), lower, upper, epoch);
    DGB_LAUNCHED();
    return 0;
}

int comm_allreduce_i64(Comm* c, long long* buf, size_t count, cudaStream_t st) {
    if (!c || c->size == 1) return 0;
    DGB_NCCL(nccl()->AllReduce(buf, buf, count, ncclInt64, ncclSum, c->comm, st));
    return 0;
}

int comm_halo_rows(Comm* c, double* interior, size_t row_len, size_t nrows, size_t ghost_rows, int periodic, cudaStream_t st) {
    const int rank = comm_rank(c), size = comm_size(c);
    const size_t cnt = ghost_rows * row_len;
    if (cnt == 0) return 0;
    if (nrows < ghost_rows) { set_error("halo exchange: slab has fewer rows than the halo"); return DGB_ERR_INVALID; }
    double* lo_ghost = interior - cnt;
    double* up_ghost = interior + nrows * row_len;
    double* bottom = interior;
    double* top = interior + (nrows - ghost_rows) * row_len;
    int lower = rank - 1, upper = rank + 1;
    if (lower < 0) lower = periodic ? size - 1 : -1;
    if (upper >= size) upper = periodic ? 0 : -1;
    if (size == 1) {  // the ring closes on this device
        if (periodic) {
            DGB_CUDA(cudaMemcpyAsync(up_ghost, bottom, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
            DGB_CUDA(cudaMemcpyAsync(lo_ghost, top, cnt * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        return 0;
    }
    NcclApi* n = nccl();
    // posting order matters when lower == upper (two ranks): a peer's first send (its bottom rows) must meet our first
    // receive from it (our upper ghost)
    DGB_NCCL(n->GroupStart());
    if (lower >= 0) DGB_NCCL(n->Send(bottom, cnt, ncclFloat64, lower, c->comm, st));
    if (upper >= 0) DGB_NCCL(n->Send(top, cnt, ncclFloat64, upper, c->comm, st));
    if (upper >= 0) DGB_NCCL(n->Recv(up_ghost, cnt, ncclFloat64, upper, c->comm, st));
    if (lower >= 0) DGB_NCCL(n->Recv(lo_ghost, cnt, ncclFloat64, lower, c->comm, st));
    DGB_NCCL(n->GroupEnd());
    return 0;
}

// MPIGather::global_gather_init / _wait (inc/dg/backend/mpi_gather.h:454-705) for packed buffers: rank r receives recv_counts[p]
// doubles from every rank p (in rank order) and sends send_counts[p] doubles to it -- one grouped ncclSend / ncclRecv round, the
// part a rank sends to itself is a device copy.  Counts live on the host (they are fixed when the gather map is built).
int comm_gather_packed(Comm* c, const double* send, const int* send_counts, double* recv, const int* recv_counts, cudaStream_t st) {
    const int rank = comm_rank(c), size = comm_size(c);
    size_t so = 0, ro = 0;
    NcclApi* n = size > 1 ? nccl() : nullptr;
    if (n) DGB_NCCL(n->GroupStart());
    for (int p = 0; p < size; p++) {
        const size_t sc = (size_t)send_counts[p], rc = (size_t)recv_counts[p];
        if (p == rank) {
            if (sc != rc) { if (n) n->GroupEnd(); set_error("dgb_comm_gather: a rank's message to itself must have equal counts"); return DGB_ERR_INVALID; }
            if (sc) DGB_CUDA(cudaMemcpyAsync(recv + ro, send + so, sc * sizeof(double), cudaMemcpyDeviceToDevice, st));
        } else {
            if (sc) DGB_NCCL(n->Send(send + so, sc, ncclFloat64, p, c->comm, st));
            if (rc) DGB_NCCL(n->Recv(recv + ro, rc, ncclFloat64, p, c->comm, st));
        }
        so += sc; ro += rc;
    }
    if (n) DGB_NCCL(n->GroupEnd());
    return 0;
}

// normalise + round a summed superaccumulator record in place (status = number of ranks that met NaN/Inf)
__global__ void superacc_finalize_kernel(dgb_dot_result* r, int nrec) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nrec) return;
    long long acc[sa::BINS];
    for (int i = 0; i < sa::BINS; i++) acc[i] = r[k].acc[i];
    int neg = sa::normalize(acc, 1);
    for (int i = 0; i < sa::BINS; i++) r[k].acc[i] = acc[i];
    r[k].value = sa::round_normalized(acc, neg);
    r[k].status = r[k].status != 0 || r[k].pad != 0;
    r[k].pad = 0;
}

// peer-memory variant of the global dot: exchange + normalise + round in one launch (channels 4..)
__global__ void __launch_bounds__(64) p2p_dot_kernel(P2pView v, dgb_dot_result* r, int nrec, unsigned long long epoch) {
    // records use channels 4.. but are indexed from 0 in r: shift the pointer so that record k sits at index 4 + k
    p2p_allreduce_records(v, reinterpret_cast<long long*>(r) - 4 * 41, 4, nrec, epoch);
    const int k = threadIdx.x;
    if (k >= nrec) return;
    long long acc[sa::BINS];
    for (int i = 0; i < sa::BINS; i++) acc[i] = r[k].acc[i];
    int neg = sa::normalize(acc, 1);
    for (int i = 0; i < sa::BINS; i++) r[k].acc[i] = acc[i];
    r[k].value = sa::round_normalized(acc, neg);
    r[k].status = r[k].status != 0 || r[k].pad != 0;
    r[k].pad = 0;
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_comm_unique_id(char* id128) {
    NcclApi* n = nccl();
    if (!n) { set_error("dgb_comm: libnccl.so.2 could not be loaded"); return DGB_ERR_UNSUPPORTED; }
    ncclUniqueId id;
    DGB_NCCL(n->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return 0;
}
int dgb_comm_create(dgb_comm** out, const char* id128, int rank, int nranks) {
    if (nranks < 1 || rank < 0 || rank >= nranks) { set_error("dgb_comm_create: invalid rank %d of %d", rank, nranks); return DGB_ERR_INVALID; }
    Comm* c = new Comm();
    c->rank = rank; c->size = nranks;
    if (nranks > 1) {
        NcclApi* n = nccl();
        if (!n) { delete c; set_error("dgb_comm: libnccl.so.2 could not be loaded"); return DGB_ERR_UNSUPPORTED; }
        ncclUniqueId id;
        memcpy(id.internal, id128, 128);
        ncclResult_t r = n->CommInitRank(&c->comm, nranks, id, rank);
        if (r != 0) { delete c; set_error("ncclCommInitRank failed with %d", r); return DGB_ERR_INVALID; }
        int e = comm_p2p_setup(c);
        if (e) { delete c; return e; }
    }
    *out = reinterpret_cast<dgb_comm*>(c);
    return 0;
}
int dgb_comm_info(const dgb_comm* h, int* rank, int* size, int* peer_memory) {
    const Comm* c = reinterpret_cast<const Comm*>(h);
    if (!c) { set_error("dgb_comm_info: communicator is NULL"); return DGB_ERR_INVALID; }
    if (rank) *rank = c->rank;
    if (size) *size = c->size;
    if (peer_memory) *peer_memory = c->p2p ? 1 : 0;
    return 0;
}
int dgb_comm_destroy(dgb_comm* h) {
    Comm* c = reinterpret_cast<Comm*>(h);
    if (!c) return 0;
    for (int r = 0; r < c->size; r++)
        if (c->p2p && r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
    cudaFree(c->local);
    if (c->comm && nccl() && nccl()->CommDestroy) nccl()->CommDestroy(c->comm);
    delete c;
    return 0;
}
int dgb_comm_halo_rows(dgb_comm* h, double* interior, size_t row_len, size_t nrows, size_t ghost_rows, int periodic, dgb_stream_t s) {
    return comm_halo_rows(reinterpret_cast<Comm*>(h), interior, row_len, nrows, ghost_rows, periodic, as_stream(s));
}
int dgb_comm_gather(dgb_comm* h, const double* send, const int* send_counts, double* recv, const int* recv_counts, dgb_stream_t s) {
    if (!send_counts || !recv_counts) { set_error("dgb_comm_gather: NULL counts"); return DGB_ERR_INVALID; }
    return comm_gather_packed(reinterpret_cast<Comm*>(h), send, send_counts, recv, recv_counts, as_stream(s));
}
// global exact dot: every rank passes the record its local dgb_exdot2/3 produced; on return all ranks hold the
// normalised global accumulator, the correctly rounded value and the OR of the status flags
int dgb_comm_allreduce_dot(dgb_comm* h, dgb_dot_result* result_dev, int nrecords, dgb_stream_t s) {
    Comm* c = reinterpret_cast<Comm*>(h);
    if (nrecords < 1) return 0;
    if (c && c->p2p && nrecords <= P2P_CHANNELS - 4) {  // channels 0..3 belong to the PCG workspace, 4.. to this entry point
        const unsigned long long epoch = comm_p2p_next_epoch(c, 4, nrecords);
        p2p_dot_kernel<<<1, 64, 0, as_stream(s)>>>(comm_p2p_view(c), result_dev, nrecords, epoch);
        DGB_LAUNCHED();
        return 0;
    }
    int e = comm_allreduce_i64(c, reinterpret_cast<long long*>(result_dev), (size_t)nrecords * sizeof(dgb_dot_result) / 8, as_stream(s));
    if (e) return e;
    superacc_finalize_kernel<<<1, 32, 0, as_stream(s)>>>(result_dev, nrecords);
    DGB_LAUNCHED();
    return 0;
}
}
