"""Multi-GPU harness: one process per GPU, the 2-d grid is cut into slabs of cell rows (y direction).
Mirrors MPI_Vector / MPISparseBlockMat usage of the reference (inc/dg/backend/mpi_vector.h, mpi_matrix.h) on top of
dgb_comm_* / dgb_elliptic2d_set_slab / dgb_pcg_solve_elliptic2d_dist.  torch.distributed is only used to ship the
NCCL unique id; all data-path communication happens inside libdgb200.so."""
import ctypes as C
import numpy as np
import torch
from ._lib import lib, DgbError
from ._dev import ptr, stream, dvec
from . import topology as T

d = C.c_double


def partition(ncells, size):
    """contiguous, as even as possible split of `ncells` cell rows over `size` ranks -> [(offset, rows)]"""
    base, rem = divmod(ncells, size)
    out, off = [], 0
    for r in range(size):
        rows = base + (1 if r < rem else 0)
        out.append((off, rows))
        off += rows
    return out


class Comm:
    def __init__(self, rank=0, size=1, unique_id=None):
        self.rank, self.size = rank, size
        self.h = C.c_void_p()
        idbuf = (C.c_char * 128).from_buffer_copy(unique_id if unique_id is not None else bytes(128))
        lib().comm_create(C.byref(self.h), idbuf, rank, size)

    @staticmethod
    def unique_id():
        buf = (C.c_char * 128)()
        lib().comm_unique_id(buf)
        return bytes(buf.raw)

    @classmethod
    def from_torch_distributed(cls):
        """rank 0 draws the NCCL id, torch.distributed broadcasts the 128 bytes"""
        import torch.distributed as dist
        rank, size = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.tensor(list(cls.unique_id()), dtype=torch.uint8)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = t.to(dev)
        dist.broadcast(t, 0)
        return cls(rank, size, bytes(t.cpu().numpy().tobytes()))

    @property
    def peer_memory(self):
        """True if dots and halos travel by CUDA-IPC peer stores over NVLink, False if by NCCL"""
        pm = C.c_int()
        lib().comm_info(self.h, None, None, C.byref(pm))
        return bool(pm.value)

    def halo_rows(self, padded, row_len, nrows, ghost_rows, periodic):
        interior = C.c_void_p(padded.data_ptr() + ghost_rows * row_len * 8)
        lib().comm_halo_rows(self.h, interior, row_len, nrows, ghost_rows, int(periodic), stream())

    def allreduce_dot(self, record):
        """record: int64 CUDA tensor of 41 words per dot (dgb_dot_result)"""
        lib().comm_allreduce_dot(self.h, ptr(record), record.numel() // 41, stream())

    def __del__(self):
        try:
            lib().comm_destroy(self.h)
        except Exception:
            pass


class SlabElliptic2d:
    """Elliptic2d on this rank's slab of a global Cartesian grid (decomposition in y)."""
    GHOST = 2  # cell rows; covers forward/backward (1) and centered (2) discretisations

    def __init__(self, comm, g, bcx=None, bcy=None, direction=T.FORWARD, jfactor=1.0):
        bcx = g.bc[0] if bcx is None else bcx
        bcy = g.bc[1] if bcy is None else bcy
        self.comm, self.grid = comm, g
        self.periodic_y = bcy == T.PER
        self.yoff, self.rows = partition(g.N[1], comm.size)[comm.rank]
        mats = dict(
            leftx=T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(direction)),
            lefty=T.derivative(1, g, T.inverse_bc(bcy), T.inverse_dir(direction)),
            rightx=T.derivative(0, g, bcx, direction), righty=T.derivative(1, g, bcy, direction),
            jumpx=T.jump(0, g, bcx), jumpy=T.jump(1, g, bcy))
        hs = {k: m.host_struct() for k, m in mats.items()}
        self._mats = mats
        self.h = C.c_void_p()
        lib().elliptic2d_create(C.byref(self.h), *[C.byref(hs[k]) for k in ("leftx", "lefty", "rightx", "righty",
                                                                           "jumpx", "jumpy")], d(jfactor), 0)
        lib().elliptic2d_set_slab(self.h, self.yoff, self.rows, self.GHOST)
        n = g.n[0]
        self.row_len = g.N[0] * n
        self.nrows = self.rows * n
        self.size = self.nrows * self.row_len
        self.ghost_rows = self.GHOST * n
        self._sigma_pad = torch.ones((self.nrows + 2 * self.ghost_rows) * self.row_len, dtype=torch.float64, device="cuda")
        self._sigma = self._sigma_pad[self.ghost_rows * self.row_len:][:self.size]
        lib().elliptic2d_set_sigma(self.h, ptr(self._sigma))
        wx, wy = g.weights1d(0), g.weights1d(1)[self.yoff * n:(self.yoff + self.rows) * n]
        self._weights = dvec((wy[:, None] * wx[None, :]).reshape(-1))  # = rows of create::weights (w_x[i] * w_y[j])
        self._precond = torch.ones(self.size, dtype=torch.float64, device="cuda")
        self._pad = None

    def local(self, global_host_vector):
        """this rank's rows of a global host vector"""
        a = np.asarray(global_host_vector).reshape(-1, self.row_len)
        n = self.grid.n[0]
        return np.ascontiguousarray(a[self.yoff * n:(self.yoff + self.rows) * n]).reshape(-1)

    def evaluate(self, f):
        """dg::evaluate restricted to this rank's rows (same abscissas as the global grid)"""
        n = self.grid.n[0]
        ax, ay = self.grid.abscissas(0), self.grid.abscissas(1)[self.yoff * n:(self.yoff + self.rows) * n]
        Y, X = np.meshgrid(ay, ax, indexing="ij")
        return np.ascontiguousarray(np.broadcast_to(f(X, Y), X.shape).reshape(-1), dtype=np.float64)

    def weights(self):
        return self._weights

    def precond(self):
        return self._precond

    def set_chi(self, sigma_local):
        """elliptic.h:324-333 on the slab; the ghost rows of sigma are exchanged once here"""
        from . import blas1
        blas1.copy(sigma_local, self._sigma)
        self.comm.halo_rows(self._sigma_pad, self.row_len, self.nrows, self.ghost_rows, self.periodic_y)
        blas1.pointwiseDivide(torch.ones_like(sigma_local), sigma_local, self._precond)

    def symv(self, *a):
        """y = A x or y = alpha A x + beta y; x is staged in a padded buffer and its halo exchanged (MPISparseBlockMat::symv,
        mpi_matrix.h:183-217)"""
        alpha, x_local, beta, y_local = (1., a[0], 0., a[1]) if len(a) == 2 else a
        if self._pad is None:
            self._pad = torch.zeros((self.nrows + 2 * self.ghost_rows) * self.row_len, dtype=torch.float64, device="cuda")
        pad = self._pad
        interior = pad[self.ghost_rows * self.row_len:][:self.size]
        interior.copy_(x_local)
        self.comm.halo_rows(pad, self.row_len, self.nrows, self.ghost_rows, self.periodic_y)
        lib().elliptic2d_symv(self.h, d(alpha), ptr(interior), d(beta), ptr(y_local), stream())

    def __del__(self):
        try:
            lib().elliptic2d_destroy(self.h)
        except Exception:
            pass


class DistPCG:
    """dg::PCG on MPI_Vector-like slabs; bit-identical to the single-GPU solve for any number of ranks"""

    def __init__(self, comm, size, max_iterations):
        self.comm, self.max_iter, self.throw_on_fail = comm, max_iterations, True
        self.h = C.c_void_p()
        lib().pcg_create(C.byref(self.h), size)

    def solve(self, A, x, b, P, W, eps=1e-12, nrmb_correction=1.0, test_frequency=1):
        it = C.c_int()
        try:
            lib().pcg_solve_elliptic2d_dist(self.h, self.comm.h, A.h, ptr(x), ptr(b), ptr(P), ptr(W), d(eps),
                                            d(nrmb_correction), test_frequency, self.max_iter, C.byref(it), stream())
        except DgbError as e:
            if e.code == -4 and not self.throw_on_fail:
                return it.value
            raise
        return it.value

    def __del__(self):
        try:
            lib().pcg_destroy(self.h)
        except Exception:
            pass
