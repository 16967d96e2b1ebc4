"""Block matrices distributed along THEIR OWN dimension on N GPUs: dg::MPISparseBlockMat + dg::make_mpi_sparseblockmat +
dg::MPIKroneckerGather (inc/dg/backend/mpi_matrix.h:100-330, mpi_gather_kron.h:150-289) on the C ABI -- the general form of
what feltor_b200/dist.py / dist_toefl.py do for y-slabs with ghost rows: any 1-d block matrix (derivative, jump, projection,
interpolation; x, y or z) with its Kronecker sizes left / right, any contiguous distribution of its block rows and columns.

As in the reference a block row that touches a block column of another rank moves as a whole into the OUTER matrix
(CooSparseBlockMat; "we need to grab the entire row to ensure reproducibility", mpi_matrix.h:253) and the INNER matrix
(EllSparseBlockMat on local column indices) keeps that row without entries.  symv follows mpi_matrix.h:183-217:

    1. pack the block columns the others asked for (dgb_gather_indexed: chunk layout [q][s][j] of CooSparseBlockMat) and start
       the exchange on a side stream (dgb_comm_gather)
    2. y = alpha inner x + beta y                          (dgb_ell_symv; communicating rows become beta y)
    3. wait
    4. y += alpha outer buffer                              (dgb_coo_symv, beta = 1)

Every block row is accumulated block by block in slot order in both matrices, so the result equals dgb_ell_symv of the global
matrix on the global vector bit for bit, for any rank count (tests/test_gpu_dist.py::test_dist_ell_equals_global,
tools/dist_check.py).  DistEllPlan is host-only (numpy)."""
import ctypes as C
import numpy as np


class DistEllPlan:
    """dg::make_mpi_sparseblockmat (mpi_matrix.h:242-330) for the block rows [row_off, row_off + rows) of `m` (an object with the
    EllSparseBlockMat fields: num_rows, num_cols, bpl, n, data, cols_idx, data_idx) when block columns are distributed as
    col_part = [(offset, count), ...]; left / right: the LOCAL Kronecker sizes."""

    def __init__(self, rank, size, m, row_off, rows, col_part, left, right):
        self.rank, self.size, self.n, self.bpl = rank, size, int(m.n), int(m.bpl)
        self.left, self.right, self.rows = int(left), int(right), int(rows)
        self.data = np.ascontiguousarray(m.data, dtype=np.float64)
        cols = np.asarray(m.cols_idx, dtype=np.int64).reshape(m.num_rows, m.bpl)[row_off:row_off + rows]
        didx = np.asarray(m.data_idx, dtype=np.int32).reshape(m.num_rows, m.bpl)[row_off:row_off + rows]
        c0, nc = col_part[rank]
        self.local_cols = int(nc)
        valid = cols >= 0
        ends = np.array([o + c for o, c in col_part], dtype=np.int64)
        offs = np.array([o for o, _ in col_part], dtype=np.int64)
        pid = np.searchsorted(ends, np.where(valid, cols, c0), side="right")
        if np.any(valid & (cols >= ends[-1])):
            raise ValueError("dg::Error: block column outside the distributed vector")
        remote = valid & (pid != rank)
        outer_row = remote.any(axis=1)
        # inner: local indices, communicating rows emptied (invalid index -1)
        inner_cols = np.where(valid & ~outer_row[:, None], cols - c0, -1).astype(np.int32)
        self.inner_cols = np.ascontiguousarray(inner_cols).reshape(-1)
        self.inner_didx = np.ascontiguousarray(np.where(inner_cols >= 0, didx, 0).astype(np.int32)).reshape(-1)
        # outer: the entries of the communicating rows, row by row in slot order
        r_idx, d_idx = np.nonzero(valid & outer_row[:, None])          # row-major: rows ascending, slots ascending
        o_pid, o_lcol = pid[r_idx, d_idx], cols[r_idx, d_idx] - offs[pid[r_idx, d_idx]]
        span = int(o_lcol.max()) + 1 if o_lcol.size else 1
        uniq, inverse = np.unique(o_pid * span + o_lcol, return_inverse=True)
        self.coo_rows = r_idx.astype(np.int32)
        self.coo_cols = inverse.astype(np.int32)                       # index into the pointer table = chunk of the buffer
        self.coo_didx = didx[r_idx, d_idx].astype(np.int32)
        upid, ulcol = uniq // span, uniq % span
        self.recv_blocks = np.bincount(upid, minlength=size).astype(np.int64)
        self.num_chunks = int(uniq.size)
        self.requests = [ulcol[upid == p].astype(np.int32) for p in range(size)]
        self.chunk = self.n * self.left * self.right                  # doubles per block column
        self.send_blocks, self.send_idx = None, None

    def set_sends(self, asked):
        """asked[p]: local block columns rank p needs from this rank -> element indices of the pack, chunk layout [q][s][j]"""
        n, L, R = self.n, self.left, self.right
        self.send_blocks = np.array([len(a) for a in asked], dtype=np.int64)
        blocks = np.concatenate([np.asarray(a, dtype=np.int64) for a in asked]) if self.send_blocks.sum() else np.zeros(0, dtype=np.int64)
        if blocks.size and (blocks.min() < 0 or blocks.max() >= self.local_cols):
            raise ValueError("dg::Error: a rank asked for a block column outside this rank's vector")
        q, s, j = np.meshgrid(np.arange(n), np.arange(L), np.arange(R), indexing="ij")
        # element (s, block, q, j) of the local vector: ((s * local_cols + block) * n + q) * right + j   (sparseblockmat.h:60-75)
        idx = ((s[None] * self.local_cols + blocks[:, None, None, None]) * n + q[None]) * R + j[None]
        if idx.size and idx.max() > np.iinfo(np.int32).max:
            raise ValueError("dist_ell: local vector too large for int indices")
        self.send_idx = idx.reshape(-1).astype(np.int32)


class DistEll:
    """dg::MPISparseBlockMat on the device; comm is a feltor_b200.dist.Comm (or, in single-process tests, any object with
    rank / size / h)"""

    def __init__(self, comm, m, row_off, rows, col_part, left, right, asked=None, group=None):
        import torch
        from ._dev import dvec
        from . import blas2
        from .dist_csr import exchange_requests
        self.comm = comm
        p = DistEllPlan(comm.rank, comm.size, m, row_off, rows, col_part, left, right)
        p.set_sends(exchange_requests(p.requests, comm.rank, comm.size, group) if asked is None else asked)
        self.plan = p
        self.inner = blas2.Ell(rows, p.local_cols, p.bpl, p.n, left, right, p.data, p.inner_cols, p.inner_didx)
        self.send_idx = dvec(p.send_idx)
        self.send_buf = torch.empty(max(1, p.send_idx.size), dtype=torch.float64, device="cuda")
        self.recv_buf = torch.empty(max(1, p.num_chunks * p.chunk), dtype=torch.float64, device="cuda")
        self._sc = (C.c_int * comm.size)(*[int(v) * p.chunk for v in p.send_blocks])
        self._rc = (C.c_int * comm.size)(*[int(v) * p.chunk for v in p.recv_blocks])
        self.data_dev = dvec(p.data)
        self.coo_rows, self.coo_cols, self.coo_didx = dvec(p.coo_rows), dvec(p.coo_cols), dvec(p.coo_didx)
        base = self.recv_buf.data_ptr()
        self.table = torch.tensor([base + 8 * p.chunk * c for c in range(max(1, p.num_chunks))], dtype=torch.int64, device="cuda")
        self._side = torch.cuda.Stream()
        self._packed, self._arrived = torch.cuda.Event(), torch.cuda.Event()

    class _Coo(C.Structure):
        _fields_ = [("num_rows", C.c_int), ("num_cols", C.c_int), ("num_entries", C.c_int), ("n", C.c_int), ("left_size", C.c_int),
                    ("right_size", C.c_int), ("data", C.c_void_p), ("rows_idx", C.c_void_p), ("cols_idx", C.c_void_p), ("data_idx", C.c_void_p)]

    def pack(self, x):
        from ._lib import lib
        from ._dev import ptr, stream
        lib().gather_indexed(self.plan.send_idx.size, ptr(self.send_idx), ptr(x), ptr(self.send_buf), stream())

    def exchange(self):
        import torch
        from ._lib import lib
        from ._dev import ptr, stream
        main = torch.cuda.current_stream()
        self._packed.record(main)
        self._side.wait_event(self._packed)
        with torch.cuda.stream(self._side):
            lib().comm_gather(self.comm.h, ptr(self.send_buf), self._sc, ptr(self.recv_buf), self._rc, stream())
            self._arrived.record(self._side)

    def apply_inner(self, alpha, x, beta, y):
        self.inner.symv(alpha, x, beta, y)

    def apply_outer(self, alpha, y):
        from ._lib import lib
        from ._dev import ptr, stream
        p = self.plan
        if p.coo_rows.size == 0:
            return
        m = DistEll._Coo(p.rows, max(1, p.num_chunks), p.coo_rows.size, p.n, p.left, p.right, self.data_dev.data_ptr(),
                         self.coo_rows.data_ptr(), self.coo_cols.data_ptr(), self.coo_didx.data_ptr())
        lib().coo_symv(C.byref(m), C.c_double(alpha), ptr(self.table), C.c_double(1.), ptr(y), stream())

    def symv(self, alpha, x, beta, y):
        """y = alpha M x + beta y on this rank's block rows (mpi_matrix.h:183-217)"""
        import torch
        p = self.plan
        if p.num_chunks == 0 and p.send_idx.size == 0:
            return self.apply_inner(alpha, x, beta, y)
        self.pack(x)
        self.exchange()
        self.apply_inner(alpha, x, beta, y)
        torch.cuda.current_stream().wait_event(self._arrived)
        self.apply_outer(alpha, y)
