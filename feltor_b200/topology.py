"""Host-side topology of the harness: grids, weights, evaluate and the DG block matrices, all produced by the
C++ builders in libdgb200.so (feltor_b200/csrc/topology.cu; reference inc/dg/topology/*)."""
import ctypes as C
import math
import numpy as np
from ._lib import lib
from .blas2 import Ell

PER, DIR, DIR_NEU, NEU_DIR, NEU = 0, 1, 2, 3, 4
FORWARD, BACKWARD, CENTERED = 0, 1, 2


def inverse_bc(bc):
    """inc/dg/enums.h:62-78"""
    return {PER: PER, DIR: NEU, NEU: DIR, DIR_NEU: NEU_DIR, NEU_DIR: DIR_NEU}[bc]


def inverse_dir(d):
    """inc/dg/enums.h:104-112"""
    return {FORWARD: BACKWARD, BACKWARD: FORWARD, CENTERED: CENTERED}[d]


class CGrid(C.Structure):
    """mirror of dgb_grid"""
    _fields_ = [("ndim", C.c_int), ("x0", C.c_double * 3), ("x1", C.c_double * 3), ("n", C.c_int * 3),
                ("N", C.c_int * 3), ("bc", C.c_int * 3)]


class Grid:
    """dg::RealGrid<double,Nd> (inc/dg/topology/grid.h); x is the fastest varying dimension."""

    def __init__(self, x0, x1, n, N, bc):
        self.ndim = len(N)
        self.x0, self.x1, self.N, self.bc = list(map(float, x0)), list(map(float, x1)), list(map(int, N)), list(bc)
        self.n = [int(n)] * self.ndim if np.isscalar(n) else list(map(int, n))
        self.c = CGrid()
        self.c.ndim = self.ndim
        for u in range(self.ndim):
            self.c.x0[u], self.c.x1[u], self.c.n[u], self.c.N[u], self.c.bc[u] = (self.x0[u], self.x1[u], self.n[u],
                                                                               self.N[u], self.bc[u])

    def ref(self):
        return C.byref(self.c)

    def shape(self, u):
        return self.n[u] * self.N[u]

    def h(self, u):
        return (self.x1[u] - self.x0[u]) / float(self.N[u])

    @property
    def size(self):
        s = 1
        for u in range(self.ndim):
            s *= self.shape(u)
        return s

    def abscissas(self, u):
        out = np.empty(self.shape(u))
        lib().topo_abscissas(self.ref(), u, out.ctypes.data)
        return out

    def weights1d(self, u):
        out = np.empty(self.shape(u))
        lib().topo_weights1d(self.ref(), u, out.ctypes.data)
        return out

    def weights(self):
        """dg::create::weights (inc/dg/topology/weights.h:60)"""
        out = np.empty(self.size)
        lib().topo_weights(self.ref(), out.ctypes.data)
        return out

    def evaluate(self, f, vectorized=True):
        """dg::evaluate (inc/dg/topology/evaluation.h:74): f(x[,y[,z]]) on the tensor grid, x fastest.
        vectorized=False calls f point by point with Python floats (libm functions => bit-identical to C)."""
        ax = [self.abscissas(u) for u in range(self.ndim)]
        if vectorized:
            mesh = np.meshgrid(*ax[::-1], indexing="ij")[::-1]
            return np.ascontiguousarray(np.broadcast_to(f(*mesh), mesh[0].shape).reshape(-1), dtype=np.float64)
        out = np.empty(self.size)
        if self.ndim == 1:
            for i, x in enumerate(ax[0]):
                out[i] = f(float(x))
        elif self.ndim == 2:
            k = 0
            for y in ax[1]:
                for x in ax[0]:
                    out[k] = f(float(x), float(y))
                    k += 1
        else:
            k = 0
            for z in ax[2]:
                for y in ax[1]:
                    for x in ax[0]:
                        out[k] = f(float(x), float(y), float(z))
                        k += 1
        return out

    def multiplied(self, fn, fN):
        """grid with n*fn coefficients and N*fN cells per dimension (x and y only for ndim == 3)"""
        n, N = list(self.n), list(self.N)
        for u in range(min(self.ndim, 2)):
            n[u] = int(n[u] * fn)
            N[u] = int(N[u] * fN)
        return Grid(self.x0, self.x1, n, N, self.bc)


def _take(handle):
    """Copy a dgb_ellh into an Ell (host numpy arrays) and free the C++ object."""
    v = Ell._Host()
    lib().ellh_view(handle, C.byref(v))
    nd = v.num_blocks * v.n * v.n
    ni = v.num_rows * v.blocks_per_line
    data = np.ctypeslib.as_array(C.cast(v.data, C.POINTER(C.c_double)), shape=(nd,)).copy()
    cols = np.ctypeslib.as_array(C.cast(v.cols_idx, C.POINTER(C.c_int)), shape=(ni,)).copy()
    didx = np.ctypeslib.as_array(C.cast(v.data_idx, C.POINTER(C.c_int)), shape=(ni,)).copy()
    m = Ell(v.num_rows, v.num_cols, v.blocks_per_line, v.n, v.left_size, v.right_size, data, cols, didx,
            (v.right_range[0], v.right_range[1]))
    lib().ellh_destroy(handle)
    return m


def dx1d(n, N, h, bc, direction=CENTERED):
    h_ = C.c_void_p()
    lib().topo_dx(C.byref(h_), n, N, C.c_double(h), bc, direction)
    return _take(h_)


def jump1d(n, N, h, bc):
    h_ = C.c_void_p()
    lib().topo_jump(C.byref(h_), n, N, C.c_double(h), bc)
    return _take(h_)


def derivative(coord, g, bc=None, direction=CENTERED):
    """dg::create::derivative (inc/dg/topology/derivatives.h:47); dx/dy/dz = coord 0/1/2"""
    h_ = C.c_void_p()
    lib().topo_derivative(C.byref(h_), g.ref(), coord, g.bc[coord] if bc is None else bc, direction)
    return _take(h_)


def jump(coord, g, bc=None):
    """dg::create::jump (inc/dg/topology/derivatives.h:68)"""
    h_ = C.c_void_p()
    lib().topo_jump_nd(C.byref(h_), g.ref(), coord, g.bc[coord] if bc is None else bc)
    return _take(h_)


def fast_projection(coord, g, dividen, divideN):
    h_ = C.c_void_p()
    lib().topo_fast_projection(C.byref(h_), g.ref(), coord, dividen, divideN)
    return _take(h_)


def fast_interpolation(coord, g, multiplyn, multiplyN):
    h_ = C.c_void_p()
    lib().topo_fast_interpolation(C.byref(h_), g.ref(), coord, multiplyn, multiplyN)
    return _take(h_)


def dlt(which, n):
    out = np.empty(n * n if which >= 2 else n)
    lib().topo_dlt(which, n, out.ctypes.data)
    return out


def window_stencil(g, window):
    """dg::create::window_stencil (inc/dg/topology/stencil.h:177-237): CSR arrays (row_offsets, cols int32; vals float64) of the
    neighbourhood matrix blas2.stencil runs on; window = points per axis"""
    window = [int(window)] * g.ndim if np.isscalar(window) else [int(w) for w in window]
    rows, per_row = g.size, int(np.prod(window))
    pos, idx, val = np.empty(rows + 1, dtype=np.int32), np.empty(rows * per_row, dtype=np.int32), np.empty(rows * per_row)
    lib().topo_window_stencil(g.ref(), (C.c_int * g.ndim)(*window), pos.ctypes.data, idx.ctypes.data, val.ctypes.data)
    return pos, idx, val


def limiter_stencil(g, direction=0, bound=None):
    """dg::create::limiter_stencil (inc/dg/topology/stencil.h:89-137,199-256): CSR arrays of the matrix blas2.stencil("slope", ...)
    = dg::CSRSlopeLimiter runs on; 1-d grid, or along `direction` (0 x, 1 y) of a 2-d grid; bound defaults to the grid's"""
    bound = g.bc[direction] if bound is None else bound
    rows = g.size
    pos, idx, val = np.empty(rows + 1, dtype=np.int32), np.empty(3 * rows, dtype=np.int32), np.empty(3 * rows)
    lib().topo_limiter_stencil(g.ref(), int(direction), int(bound), pos.ctypes.data, idx.ctypes.data, val.ctypes.data)
    return pos, idx, val

