"""Row-distributed CSR matrix - vector product on N GPUs: dg::MPIDistMat + dg::MPIGather + dg::make_mpi_matrix
(inc/dg/backend/mpi_matrix.h:394-523, mpi_gather.h:454-705, topology/mpi_projection.h:50-124) on the C ABI.

Every rank owns a block of rows (with GLOBAL column indices) and a block of the vector.  As in the reference a row that touches
any element of another rank moves to the OUTER matrix as a whole ("we need to grab the entire row to ensure reproducibility",
mpi_projection.h:56) and all its operands -- the rank's own ones included -- arrive through the gather buffer; the other rows
form the INNER matrix on local column indices.  symv follows mpi_matrix.h:492-521:

    1. pack the elements the others asked for (dgb_gather_indexed) and start the exchange (dgb_comm_gather, own stream)
    2. y = inner x                                       (dgb_csr_spmv; outer rows of y become 0)
    3. wait for the exchange
    4. y[scatter[i]] += (outer row i) . buffer           (dgb_csr_spmv_scatter_add: product and scatter fused)

Each row is summed in the order of its entries by one thread, so the result equals the single-GPU dgb_csr_spmv of the global
matrix bit for bit for any number of ranks and any distribution of the columns (tests/test_gpu_dist.py, tools/dist_check.py).

DistCsrPlan is host-only (numpy): tests/test_dist_cpu.py runs it under gloo."""
import ctypes as C
import numpy as np


def contiguous_owner(col_part):
    """global2localIdx of a contiguous column distribution [(offset, count), ...] (one entry per rank)"""
    offs = np.array([o for o, _ in col_part], dtype=np.int64)
    ends = np.array([o + c for o, c in col_part], dtype=np.int64)

    def global2local(gidx):
        gidx = np.asarray(gidx, dtype=np.int64)
        pid = np.searchsorted(ends, gidx, side="right")
        if gidx.size and (gidx.min() < 0 or pid.max() >= len(col_part)):
            raise ValueError("dg::Error: column index outside the global vector")
        return pid.astype(np.int64), gidx - offs[pid]
    return global2local


class DistCsrPlan:
    """dg::make_mpi_matrix (mpi_projection.h:50-124) for this rank's rows; numpy only.

    pos, idx, val: CSR of the local rows with global column indices; global2local(idx) -> (owner rank, local index).
    After construction `requests[p]` lists the local indices this rank needs from rank p (in buffer order); feed what the other
    ranks need from this one to set_sends()."""

    def __init__(self, rank, size, pos, idx, val, global2local, local_cols):
        pos = np.asarray(pos, dtype=np.int64)
        idx, val = np.asarray(idx), np.asarray(val, dtype=np.float64)
        nrows = pos.size - 1
        counts = np.diff(pos)
        pid, lidx = global2local(idx)
        row_of_entry = np.repeat(np.arange(nrows), counts)
        outer_row = np.zeros(nrows, dtype=bool)
        outer_row[row_of_entry[pid != rank]] = True              # 1st pass, mpi_projection.h:66-73
        entry_outer = outer_row[row_of_entry]
        self.rank, self.size, self.num_rows, self.local_cols = rank, size, nrows, int(local_cols)
        # inner matrix: all rows, the communicating ones empty (:104)
        self.inner_pos = np.concatenate([[0], np.cumsum(np.where(outer_row, 0, counts))]).astype(np.int32)
        self.inner_idx = lidx[~entry_outer].astype(np.int32)
        self.inner_val = val[~entry_outer]
        if self.inner_idx.size and self.inner_idx.max() >= local_cols:
            raise ValueError("dg::Error: local column index outside the local vector")
        # outer matrix: the communicating rows, compressed, with their scatter map (:99-103)
        self.scatter = np.nonzero(outer_row)[0].astype(np.int32)
        self.outer_pos = np.concatenate([[0], np.cumsum(counts[self.scatter])]).astype(np.int32)
        self.outer_val = val[entry_outer]
        opid, olidx = pid[entry_outer], lidx[entry_outer]
        # unique (rank, local index) pairs in ascending order = layout of the receive buffer (gIdx2unique_idx, :115-118)
        span = int(olidx.max()) + 1 if olidx.size else 1
        uniq, inverse = np.unique(opid * span + olidx, return_inverse=True)
        self.outer_idx = inverse.astype(np.int32)
        upid, ulidx = uniq // span, uniq % span
        self.recv_counts = np.bincount(upid, minlength=size).astype(np.int32)
        self.buffer_size = int(uniq.size)
        self.requests = [ulidx[upid == p].astype(np.int32) for p in range(size)]
        self.send_idx, self.send_counts = None, None

    def is_communicating(self):
        return self.buffer_size > 0

    def set_sends(self, asked):
        """asked[p]: the local indices rank p needs from this rank (its requests[this rank])"""
        assert len(asked) == self.size
        self.send_counts = np.array([len(a) for a in asked], dtype=np.int32)
        self.send_idx = (np.concatenate([np.asarray(a, dtype=np.int32) for a in asked]) if self.send_counts.sum() else
                         np.zeros(0, dtype=np.int32))
        if self.send_idx.size and (self.send_idx.min() < 0 or self.send_idx.max() >= self.local_cols):
            raise ValueError("dg::Error: a rank asked for an element outside this rank's vector")
        if not np.array_equal(np.asarray(asked[self.rank], dtype=np.int32), self.requests[self.rank]):
            raise ValueError("the message of a rank to itself does not match its own request")


def exchange_requests(requests, rank, size, group=None):
    """every rank learns what the others need from it (the setup step of MPIGather, mpi_gather.h:476-520); host data through
    torch.distributed (any backend) -- setup only, the data path never goes through here"""
    if size == 1:
        return [requests[0]]
    import torch.distributed as dist
    everyone = [None] * size
    dist.all_gather_object(everyone, [np.asarray(r).tolist() for r in requests], group=group)
    return [np.asarray(everyone[p][rank], dtype=np.int32) for p in range(size)]


class DistCsr:
    """dg::MPIDistMat in row_dist mode on the device; comm is a feltor_b200.dist.Comm"""

    def __init__(self, comm, pos, idx, val, global2local, local_cols, asked=None, group=None):
        import torch
        from ._dev import dvec
        self.comm = comm
        p = DistCsrPlan(comm.rank, comm.size, pos, idx, val, global2local, local_cols)
        p.set_sends(exchange_requests(p.requests, comm.rank, comm.size, group) if asked is None else asked)
        self.plan = p
        self.inner = tuple(dvec(a) for a in (p.inner_pos, p.inner_idx, p.inner_val))
        self.outer = tuple(dvec(a) for a in (p.outer_pos, p.outer_idx, p.outer_val))
        self.scatter, self.send_idx = dvec(p.scatter), dvec(p.send_idx)
        self.send_buf = torch.empty(max(1, p.send_idx.size), dtype=torch.float64, device="cuda")
        self.recv_buf = torch.empty(max(1, p.buffer_size), dtype=torch.float64, device="cuda")
        self._sc = (C.c_int * comm.size)(*[int(v) for v in p.send_counts])
        self._rc = (C.c_int * comm.size)(*[int(v) for v in p.recv_counts])
        self._side = torch.cuda.Stream()
        self._packed, self._arrived = torch.cuda.Event(), torch.cuda.Event()

    # the four steps; tests drive them one by one to emulate several ranks in one process
    def pack(self, x):
        from ._lib import lib
        from ._dev import ptr, stream
        lib().gather_indexed(self.plan.send_idx.size, ptr(self.send_idx), ptr(x), ptr(self.send_buf), stream())

    def exchange(self):
        """global_gather_init: on a side stream so that the inner product overlaps the transfer"""
        import torch
        from ._lib import lib
        from ._dev import ptr, stream
        main = torch.cuda.current_stream()
        self._packed.record(main)
        self._side.wait_event(self._packed)
        with torch.cuda.stream(self._side):
            lib().comm_gather(self.comm.h, ptr(self.send_buf), self._sc, ptr(self.recv_buf), self._rc, stream())
            self._arrived.record(self._side)

    def apply_inner(self, x, y):
        from ._lib import lib
        from ._dev import ptr, stream
        lib().csr_spmv(self.plan.num_rows, self.plan.local_cols, *[ptr(a) for a in self.inner], C.c_double(1.), ptr(x), C.c_double(0.),
                       ptr(y), stream())

    def apply_outer(self, y):
        from ._lib import lib
        from ._dev import ptr, stream
        lib().csr_spmv_scatter_add(self.plan.scatter.size, *[ptr(a) for a in self.outer], ptr(self.recv_buf), ptr(self.scatter), ptr(y),
                                   stream())

    def symv(self, x, y):
        """y = M x on this rank's rows (mpi_matrix.h:478-523)"""
        import torch
        if not self.plan.is_communicating() and self.plan.send_idx.size == 0:
            return self.apply_inner(x, y)
        self.pack(x)
        self.exchange()
        self.apply_inner(x, y)
        torch.cuda.current_stream().wait_event(self._arrived)
        self.apply_outer(y)


class DistCsrAllreduce:
    """dg::MPIDistMat in allreduce mode (mpi_matrix.h:438-441,487-490), the form dg::Average takes when the averaged axis is
    distributed: every rank holds the COLUMN block of the matrix that belongs to its part of the vector (local column indices,
    all rows), applies it, and the partial results are summed over the ranks.  The sum runs in rank order on every rank
    (dgb_comm_gather of the partial vectors + dgb_sum_ranks), so all ranks end up with identical bits; it differs from the
    single-GPU product only by the rounding of that regrouping (the reference's MPI_Allreduce differs the same way)."""

    def __init__(self, comm, num_rows, local_cols, pos, idx, val):
        import torch
        from ._dev import dvec
        self.comm, self.num_rows, self.local_cols = comm, int(num_rows), int(local_cols)
        self.mat = tuple(dvec(np.ascontiguousarray(a, dtype=t)) for a, t in ((pos, np.int32), (idx, np.int32), (val, np.float64)))
        self.partial = torch.empty(max(1, self.num_rows), dtype=torch.float64, device="cuda")
        self.send = torch.empty(max(1, self.num_rows * comm.size), dtype=torch.float64, device="cuda")
        self.parts = torch.empty(max(1, self.num_rows * comm.size), dtype=torch.float64, device="cuda")
        self._cnt = (C.c_int * comm.size)(*([self.num_rows] * comm.size))

    def apply_local(self, x):
        from ._lib import lib
        from ._dev import ptr, stream
        lib().csr_spmv(self.num_rows, self.local_cols, *[ptr(a) for a in self.mat], C.c_double(1.), ptr(x), C.c_double(0.), ptr(self.partial), stream())

    def reduce(self, y):
        from ._lib import lib
        from ._dev import ptr, stream
        lib().sum_ranks(self.comm.size, self.num_rows, ptr(self.parts), ptr(y), stream())

    def symv(self, x, y):
        """y = M x with x distributed over the ranks and y replicated on all of them"""
        from ._lib import lib
        from ._dev import ptr, stream
        self.apply_local(x)
        if self.comm.size == 1:
            y.copy_(self.partial[:self.num_rows])
            return
        m = self.num_rows
        for r in range(self.comm.size):        # the same partial vector goes to every rank
            self.send[r * m:(r + 1) * m].copy_(self.partial[:m])
        lib().comm_gather(self.comm.h, ptr(self.send), self._cnt, ptr(self.parts), self._cnt, stream())
        self.reduce(y)

