"""toefl (config 3) on N GPUs: the right-hand side toefl::Explicit (feltor_b200/toefl.py, itself a call-by-call mirror of
src/toefl/toefl.h) on a y-decomposed grid -- the MPI_Vector / MPISparseBlockMat / MPI multigrid usage of the reference
(inc/dg/backend/mpi_matrix.h:183-330, inc/dg/multigrid.h with MPI grids) on top of the C ABI:

  * block matrices are cut into this rank's rows on the host (`slab_ell`): an x-operator only shrinks its left size, a
    y-operator keeps its blocks and gets its column indices remapped into a vector with ghost rows -- the same data the
    reference's MPI matrices hold (inner + outer part), applied by the same dgb_ell_symv kernels, so every row is computed
    with the single-GPU arithmetic;
  * ghost rows travel with dgb_comm_halo_rows (one exchange per field, shared by all y-operators applied to it);
  * the Helmholtz / polarisation operators are slab plans of the fused Elliptic kernel, solved by the distributed PCG
    (peer-memory dots and halos inside the solver's kernels);
  * the nested-iteration multigrid (multigrid.h:197-245, 617-658) is restated here with slab transfers: projection and
    interpolation are cell-local, so they need no communication when the slab boundaries fall on coarse cell boundaries.

Harness code (tests, bench), not the product.  Results are BITWISE those of the single-GPU run for any number of ranks
(tests/test_gpu_dist.py with a communicator of size 1, tools/dist_check.py under torchrun)."""
import ctypes as C
import numpy as np
import torch
from ._lib import lib
from ._dev import ptr, stream, dvec
from . import blas1, blas2, topology as T
from .dist import SlabElliptic2d, DistPCG, partition
from . import toefl as TF

d = C.c_double
GHOST = SlabElliptic2d.GHOST


def _zeros(n):
    return torch.zeros(n, dtype=torch.float64, device="cuda")


def slab_ell(m, coord, yoff, rows, ghost=0, periodic=False, col_off=None, col_rows=None):
    """rows [yoff, yoff + rows) of the cell-row space of the 2-d block matrix m (a blas2.Ell-like host object).
    coord 0: m acts along x (y is its left index): only the left size changes.
    coord 1: m acts along y: keep the block rows of the slab; column c becomes c - col_off (+ ghost), wrapped periodically
    into the ghost rows; the operand has col_rows (+ 2 ghost) cell rows."""
    if coord == 0:
        ny_n = m.left_size // m._cells_y
        return blas2.Ell(m.num_rows, m.num_cols, m.bpl, m.n, rows * ny_n, m.right_size, m.data, m.cols_idx, m.data_idx)
    col_off = yoff if col_off is None else col_off
    col_rows = rows if col_rows is None else col_rows
    cols = m.cols_idx.reshape(m.num_rows, m.bpl)[yoff:yoff + rows].astype(np.int64)
    didx = m.data_idx.reshape(m.num_rows, m.bpl)[yoff:yoff + rows]
    valid = cols >= 0
    loc = cols - col_off + ghost
    if periodic:
        loc = np.where(loc < 0, loc + m.num_cols, loc)
        loc = np.where(loc >= col_rows + 2 * ghost, loc - m.num_cols, loc)
    if np.any(valid & ((loc < 0) | (loc >= col_rows + 2 * ghost))):
        raise ValueError("slab_ell: a block column lies outside the slab and its ghost rows")
    loc = np.where(valid, loc, cols)
    return blas2.Ell(rows, col_rows + 2 * ghost, m.bpl, m.n, m.left_size, m.right_size, m.data, loc.astype(np.int32).reshape(-1),
                     np.ascontiguousarray(didx).reshape(-1))


class Slab:
    """this rank's rows of a global 2-d grid and the padded buffers / halo exchange of its vectors"""

    def __init__(self, comm, g, yoff=None, rows=None):
        self.comm, self.grid = comm, g
        if yoff is None:
            yoff, rows = partition(g.N[1], comm.size)[comm.rank]
        self.yoff, self.rows = yoff, rows
        self.n = g.n[0]
        self.row_len = g.N[0] * self.n
        self.nrows = rows * self.n
        self.size = self.nrows * self.row_len
        self.ghost_rows = GHOST * self.n
        self.periodic = g.bc[1] == T.PER
        self._pads = {}

    def evaluate(self, f):
        n = self.n
        ax, ay = self.grid.abscissas(0), self.grid.abscissas(1)[self.yoff * n:(self.yoff + self.rows) * n]
        Y, X = np.meshgrid(ay, ax, indexing="ij")
        return np.ascontiguousarray(np.broadcast_to(f(X, Y), X.shape).reshape(-1), dtype=np.float64)

    def local(self, v):
        a = np.asarray(v).reshape(-1, self.row_len)
        return np.ascontiguousarray(a[self.yoff * self.n:(self.yoff + self.rows) * self.n]).reshape(-1)

    def padded(self, x, slot=0):
        """x with its ghost rows exchanged; returns the padded buffer (valid until the next call with the same slot)"""
        pad = self._pads.get(slot)
        if pad is None:
            pad = self._pads[slot] = _zeros((self.nrows + 2 * self.ghost_rows) * self.row_len)
        pad[self.ghost_rows * self.row_len:][:self.size].copy_(x)
        self.comm.halo_rows(pad, self.row_len, self.nrows, self.ghost_rows, self.periodic)
        return pad

    def dx(self, bc, direction):
        m = T.derivative(0, self.grid, bc, direction)
        m._cells_y = self.grid.N[1]
        return slab_ell(m, 0, self.yoff, self.rows)

    def dy(self, bc, direction):
        return slab_ell(T.derivative(1, self.grid, bc, direction), 1, self.yoff, self.rows, GHOST, self.periodic)


class DistAdvection:
    """dg::Advection::upwind (advection.h:112-120) on a slab: one halo exchange of f serves both y-derivatives"""

    def __init__(self, S, bcx, bcy):
        self.S = S
        self.dxf, self.dxb = S.dx(bcx, T.FORWARD), S.dx(bcx, T.BACKWARD)
        self.dyf, self.dyb = S.dy(bcy, T.FORWARD), S.dy(bcy, T.BACKWARD)
        self.t0, self.t1 = _zeros(S.size), _zeros(S.size)

    def upwind(self, alpha, vx, vy, f, beta, result):
        n = f.numel()
        self.dxb.symv(1., f, 0., self.t0)
        self.dxf.symv(1., f, 0., self.t1)
        lib().upwind_axpby(n, d(alpha), ptr(vx), ptr(self.t0), ptr(self.t1), d(beta), ptr(result), stream())
        fp = self.S.padded(f)
        self.dyb.symv(1., fp, 0., self.t0)
        self.dyf.symv(1., fp, 0., self.t1)
        lib().upwind_axpby(n, d(alpha), ptr(vy), ptr(self.t0), ptr(self.t1), d(1.), ptr(result), stream())


class SlabTransfer:
    """MultiMatrix X-then-Y product of fast_projection / fast_interpolation (fast_interpolation.h:71-84,380-398) on slabs"""

    def __init__(self, Sfrom, Sto, projection):
        g = Sfrom.grid
        if projection:
            gx = T.Grid(g.x0, g.x1, g.n, [g.N[0] // 2, g.N[1]], g.bc)
            mx, my = T.fast_projection(0, g, 1, 2), T.fast_projection(1, gx, 1, 2)
        else:
            gx = T.Grid(g.x0, g.x1, g.n, [g.N[0] * 2, g.N[1]], g.bc)
            mx, my = T.fast_interpolation(0, g, 1, 2), T.fast_interpolation(1, gx, 1, 2)
        mx._cells_y = g.N[1]
        self.mx = slab_ell(mx, 0, Sfrom.yoff, Sfrom.rows)
        self.my = slab_ell(my, 1, Sto.yoff, Sto.rows, 0, False, col_off=Sfrom.yoff, col_rows=Sfrom.rows)
        self.temp = _zeros(self.mx.total_rows)

    def symv(self, alpha, x, beta, y):
        self.mx.symv(1., x, 0., self.temp)
        self.my.symv(alpha, self.temp, beta, y)


class DistMultigridCG2d:
    """dg::MultigridCG2d on y-slabs: same nested iteration, distributed PCG on every stage"""

    def __init__(self, comm, g, stages):
        self.comm, self.stages = comm, stages
        y0, rows = partition(g.N[1], comm.size)[comm.rank]
        if g.N[1] % comm.size or rows % (1 << (stages - 1)) or g.N[0] % (1 << (stages - 1)):
            raise ValueError("DistMultigridCG2d: the cell rows of every rank must be divisible by 2^(stages-1)")
        self.slabs, self.grids = [], []
        for u in range(stages):
            gu = T.Grid(g.x0, g.x1, g.n, [g.N[0] >> u, g.N[1] >> u], g.bc)
            self.grids.append(gu)
            self.slabs.append(Slab(comm, gu, y0 >> u, rows >> u))
        self.sizes = [S.size for S in self.slabs]
        self.project_ = [SlabTransfer(self.slabs[u], self.slabs[u + 1], True) for u in range(stages - 1)]
        self.inter_ = [SlabTransfer(self.slabs[u + 1], self.slabs[u], False) for u in range(stages - 1)]
        self.x, self.r, self.b, self.w = ([_zeros(s) for s in self.sizes] for _ in range(4))
        self.pcg = [DistPCG(comm, self.sizes[u], self.grids[u].size) for u in range(stages)]

    def grid(self, u):
        return self.grids[u]

    def slab(self, u):
        return self.slabs[u]

    def project(self, src):
        out = [torch.empty(s, dtype=torch.float64, device="cuda") for s in self.sizes]
        blas1.copy(src, out[0])
        for u in range(self.stages - 1):
            self.project_[u].symv(1., out[u], 0., out[u + 1])
        return out

    def solve(self, ops, x, b, eps):
        """nested_iterations (multigrid.h:197-245) with the stage solvers of MultigridCG2d::solve (:640-648)"""
        S = self.stages
        eps = [eps] * S if np.isscalar(eps) else list(eps)
        X, R, B, W = self.x, self.r, self.b, self.w
        ops[0].symv(x, R[0])
        blas1.axpby(1., b, -1., R[0])
        blas1.copy(x, X[0])
        for u in range(S - 1):
            self.project_[u].symv(1., R[u], 0., R[u + 1])
            self.project_[u].symv(1., X[u], 0., X[u + 1])
            ops[u + 1].symv(X[u + 1], B[u + 1])
            blas1.axpby(1., B[u + 1], 1., R[u + 1], B[u + 1])
            blas1.copy(X[u + 1], W[u + 1])
        numbers = [0] * S
        for u in range(S - 1, 0, -1):
            numbers[u] = self.pcg[u].solve(ops[u], X[u], B[u], ops[u].precond(), ops[u].weights(), eps[u], 1., 10)
            blas1.axpby(1., X[u], -1., W[u], X[u])
            self.inter_[u - 1].symv(1., X[u], 1., X[u - 1])
        blas1.copy(X[0], x)
        numbers[0] = self.pcg[0].solve(ops[0], x, b, ops[0].precond(), ops[0].weights(), eps[0], 1., 1)
        return numbers


class DistExplicit(TF.Explicit):
    """toefl::Explicit on this rank's slab: same members and call sequence as feltor_b200.toefl.Explicit"""

    def __init__(self, comm, p):
        self.p, self.comm = p, comm
        g = T.Grid([0., 0.], [p.lx, p.ly], p.n, [p.Nx, p.Ny], [p.bcx, p.bcy])
        self.grid = g
        self.multigrid = DistMultigridCG2d(comm, g, p.num_stages)
        S = self.slab = self.multigrid.slab(0)
        n = S.size
        self.chi, self.omega, self.uE2 = _zeros(n), _zeros(n), _zeros(n)
        from fractions import Fraction
        a, b = Fraction(p.kappa), Fraction(1. - p.kappa * p.posX * p.lx)
        line = np.array([float(a * Fraction(float(x)) + b) for x in g.abscissas(0)])
        self.binv = dvec(np.ascontiguousarray(np.broadcast_to(line, (S.nrows, g.shape(0))).reshape(-1)))
        self.phi = [_zeros(n), _zeros(n)]
        self.dxphi, self.dyphi = [_zeros(n), _zeros(n)], [_zeros(n), _zeros(n)]
        self.ype, self.lapy, self.v = [_zeros(n), _zeros(n)], [_zeros(n), _zeros(n)], [_zeros(n), _zeros(n)]
        self.gamma_n = _zeros(n)
        self.laplaceM = SlabElliptic2d(comm, g, p.bcx, p.bcy, p.diff_dir, 1.)
        self.adv = DistAdvection(S, p.bcx, p.bcy)
        self.old_phi, self.old_psi, self.old_gammaN = (TF.Extrapolation(2, self.chi) for _ in range(3))
        self.multi_chi = self.multigrid.project(self.chi)
        self.multi_pol = [SlabElliptic2d(comm, self.multigrid.grid(u), p.bcx, p.bcy, p.pol_dir, 1.) for u in range(p.num_stages)]
        self.multi_gamma1 = [TF.Helmholtz(-0.5 * p.tau, SlabElliptic2d(comm, self.multigrid.grid(u), p.bcx, p.bcy, p.pol_dir, 1.))
                             for u in range(p.num_stages)]
        self._cdx, self._cdy = S.dx(p.bcx, T.CENTERED), S.dy(p.bcy, T.CENTERED)
        outer = self

        class _Dy:   # blas2::symv(m_centered[1], phi, dyphi) with the halo exchange of its operand
            def symv(self, alpha, x, beta, y):
                outer._cdy.symv(alpha, S.padded(x), beta, y)
        self.centered = [self._cdx, _Dy()]
        # Elliptic::variation of the finest polarisation operator: its right derivatives on the slab
        self._rx, self._ry = S.dx(p.bcx, p.pol_dir), S.dy(p.bcy, p.pol_dir)
        self._tx, self._ty = _zeros(n), _zeros(n)
        self.ncalls = 0
        self.numbers = {}

    def initial_condition(self):
        p, S = self.p, self.slab
        x0, y0, s = p.posX * p.lx, p.posY * p.ly, p.sigma
        gauss = S.evaluate(lambda x, y: p.amp * np.exp(-((x - x0) * (x - x0) / 2. / s / s + (y - y0) * (y - y0) / 2. / s / s)))
        y = [dvec(gauss), dvec(gauss)]
        if p.tau != 0 and p.flr == "gamma_inv":
            self.gamma_inv().symv(y[0], y[1])
        return y

    def _variation(self, phi, out):
        """elliptic.h:497-502 with lambda = 1 and the identity metric"""
        self._rx.symv(1., phi, 0., self._tx)
        self._ry.symv(1., self.slab.padded(phi), 0., self._ty)
        lib().tensor_dot2d(phi.numel(), d(1.), None, d(1.), ptr(self._tx), ptr(self._ty), None, None, None, None, None, d(1.),
                           ptr(self._tx), ptr(self._ty), d(0.), ptr(out), stream())
