"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libdgref.so (the unmodified reference,
OpenMP backend, wrapped by oracle/ref_wrap.cpp).  Only tests/, bench.py's reference/cpu_baseline legs,
__graft_entry__.smoke() and tests/golden/make_golden.py may import this module."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdgref.so")

PER, DIR, DIR_NEU, NEU_DIR, NEU = 0, 1, 2, 3, 4
FORWARD, BACKWARD, CENTERED = 0, 1, 2

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)


class RefGrid(C.Structure):
    _fields_ = [("ndim", C.c_int), ("x0", C.c_double * 3), ("x1", C.c_double * 3),
                ("n", C.c_int * 3), ("N", C.c_int * 3), ("bc", C.c_int * 3)]


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_round.restype = C.c_double
        _lib.ref_elliptic2d_time.restype = C.c_double
        for name in ("ref_ell_create", "ref_ell_from_arrays", "ref_elliptic2d_create", "ref_multigrid_create"):
            getattr(_lib, name).restype = C.c_void_p
    return _lib


def dp(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_ip)


def lp(a):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_lp)


def grid(x0, x1, n, N, bc):
    """x0,x1,N,bc: sequences of length ndim; n scalar."""
    g = RefGrid()
    g.ndim = len(N)
    for u in range(g.ndim):
        g.x0[u], g.x1[u], g.n[u], g.N[u], g.bc[u] = x0[u], x1[u], n, N[u], bc[u]
    if g.ndim == 3:
        g.n[2] = 1  # dg::CartesianGrid3d has one coefficient per cell in z (inc/dg/topology/grid.h:738)
    return g


def grid_size(g):
    s = 1
    for u in range(g.ndim):
        s *= g.n[u] * g.N[u]
    return s


def abscissas(g, u):
    out = np.empty(g.n[u] * g.N[u])
    lib().ref_abscissas(C.byref(g), u, dp(out))
    return out


def weights1d(g, u):
    out = np.empty(g.n[u] * g.N[u])
    lib().ref_weights1d(C.byref(g), u, dp(out))
    return out


def weights(g):
    out = np.empty(grid_size(g))
    lib().ref_weights(C.byref(g), dp(out))
    return out


FUNC2 = {"zero": 0, "one": 1, "sinsin": 2, "cosxsiny": 3, "cosysinx": 4, "sincos": 5, "cossin": 6, "pol": 7,
         "rhs": 8, "expexp": 9, "shear": 10}
FUNC3 = {"zero": 0, "sin3": 1, "cosx3": 2, "cosy3": 3, "cosz3": 4, "exp3": 5}
FUNC1 = {"exp": 0, "sin": 1}


def evaluate(g, name):
    f = {1: FUNC1, 2: FUNC2, 3: FUNC3}[g.ndim][name]
    out = np.empty(grid_size(g))
    lib().ref_evaluate(C.byref(g), f, dp(out))
    return out


def dlt(which, n):
    out = np.empty(n * n if which >= 2 else n)
    lib().ref_dlt(which, n, dp(out))
    return out


class Ell:
    """Host copy of an EllSparseBlockMat plus a live reference handle."""

    def __init__(self, handle):
        self.h = C.c_void_p(handle)
        meta = np.zeros(10, dtype=np.int32)
        lib().ref_ell_meta(self.h, ip(meta))
        (self.num_rows, self.num_cols, self.bpl, self.n, self.left_size, self.right_size, self.nblocks, rr0,
         rr1) = [int(v) for v in meta[:9]]
        self.right_range = (rr0, rr1)
        self.data = np.empty(self.nblocks * self.n * self.n)
        self.cols_idx = np.empty(self.num_rows * self.bpl, dtype=np.int32)
        self.data_idx = np.empty(self.num_rows * self.bpl, dtype=np.int32)
        lib().ref_ell_arrays(self.h, dp(self.data), ip(self.cols_idx), ip(self.data_idx))

    @property
    def total_rows(self):
        return self.num_rows * self.n * self.left_size * self.right_size

    @property
    def total_cols(self):
        return self.num_cols * self.n * self.left_size * self.right_size

    def meta(self):
        return np.array([self.num_rows, self.num_cols, self.bpl, self.n, self.left_size, self.right_size,
                         self.nblocks, self.right_range[0], self.right_range[1], 0], dtype=np.int32)

    def symv(self, alpha, x, beta, y):
        assert x.size == self.total_cols and y.size == self.total_rows
        return lib().ref_ell_symv(self.h, C.c_double(alpha), dp(x), C.c_double(beta), dp(y))

    def __del__(self):
        try:
            lib().ref_ell_free(self.h)
        except Exception:
            pass


def ell_create(g, kind, coord, bc=PER, direction=CENTERED, a=1, b=1):
    kinds = {"derivative": 0, "jump": 1, "fast_projection": 2, "fast_interpolation": 3}
    return Ell(lib().ref_ell_create(C.byref(g), kinds[kind], coord, bc, direction, a, b))


def ell_from_arrays(meta, data, cols, didx):
    return Ell(lib().ref_ell_from_arrays(ip(np.ascontiguousarray(meta, dtype=np.int32)), dp(data), ip(cols), ip(didx)))


def csr_symv(nrows, ncols, pos, idx, val, alpha, x, beta, y):
    return lib().ref_csr_symv(nrows, ncols, len(val), ip(pos), ip(idx), dp(val), C.c_double(alpha), dp(x),
                              C.c_double(beta), dp(y))


class Csr:
    """the reference's IDMatrix built once (Fieldaligned keeps its interpolation matrices), symv per call"""

    def __init__(self, nrows, ncols, pos, idx, val):
        lib().ref_csr_create.restype = C.c_void_p
        self.nrows, self.ncols = nrows, ncols
        self.h = C.c_void_p(lib().ref_csr_create(nrows, ncols, len(val), ip(pos), ip(idx), dp(val)))

    def symv(self, alpha, x, beta, y):
        return lib().ref_csr_apply(self.h, self.nrows, self.ncols, C.c_double(alpha), dp(x), C.c_double(beta), dp(y))

    def __del__(self):
        try:
            lib().ref_csr_free(self.h)
        except Exception:
            pass


def dot2(x, y):
    acc = np.zeros(39, dtype=np.int64)
    st = lib().ref_dot2(x.size, dp(x), dp(y), lp(acc))
    return acc, st


def dot3(x, w, y):
    acc = np.zeros(39, dtype=np.int64)
    st = lib().ref_dot3(x.size, dp(x), dp(w), dp(y), lp(acc))
    return acc, st


def vdot(kind, x, y):
    """dg::blas1::vdot (blas1.h:90): kind 0 dg::Product, 1 a user functor a*b + a/4"""
    lib().ref_vdot.restype = C.c_double
    return float(lib().ref_vdot(int(kind), x.size, dp(np.ascontiguousarray(x)), dp(np.ascontiguousarray(y))))


def reduce(kind, x):
    """dg::blas1::reduce (blas1.h:215): kind 0 sum of squares, 1 max |x|, 2 min"""
    lib().ref_reduce.restype = C.c_double
    return float(lib().ref_reduce(int(kind), x.size, dp(np.ascontiguousarray(x))))


def round_acc(acc):
    return float(lib().ref_round(lp(np.ascontiguousarray(acc, dtype=np.int64))))


def d(v):
    return C.c_double(v)


class Elliptic2d:
    def __init__(self, g, bcx, bcy, direction=FORWARD, jfactor=1.0, chi_weight_jump=False):
        self.h = C.c_void_p(lib().ref_elliptic2d_create(C.byref(g), bcx, bcy, direction, d(jfactor),
                                                        int(chi_weight_jump)))
        self.size = grid_size(g)

    def set_chi(self, sigma):
        lib().ref_elliptic2d_set_chi(self.h, dp(sigma))

    def symv(self, alpha, x, beta, y):
        lib().ref_elliptic2d_symv(self.h, d(alpha), dp(x), d(beta), dp(y))

    def variation(self, alpha, lam, phi, beta, sigma):
        lib().ref_elliptic2d_variation(self.h, d(alpha), dp(lam), dp(phi), d(beta), dp(sigma))

    def weights(self):
        out = np.empty(self.size)
        lib().ref_elliptic2d_weights(self.h, dp(out))
        return out

    def precond(self):
        out = np.empty(self.size)
        lib().ref_elliptic2d_precond(self.h, dp(out))
        return out

    def pcg_solve(self, x, b, P, W, eps, nrmb_correction=1.0, test_frequency=1, max_iter=None):
        sec = C.c_double(0)
        it = lib().ref_pcg_solve(self.h, dp(x), dp(b), dp(P), dp(W), d(eps), d(nrmb_correction), test_frequency,
                                 max_iter or self.size, C.byref(sec))
        return it, sec.value

    def time_symv(self, x, y, reps):
        return float(lib().ref_elliptic2d_time(self.h, dp(x), dp(y), reps))

    def __del__(self):
        try:
            lib().ref_elliptic2d_free(self.h)
        except Exception:
            pass


class Multigrid:
    def __init__(self, g, stages, direction=FORWARD, jfactor=1.0):
        self.h = C.c_void_p(lib().ref_multigrid_create(C.byref(g), stages, direction, d(jfactor)))
        self.stages = stages
        self.sizes = [lib().ref_multigrid_stage_size(self.h, u) for u in range(stages)]

    def project(self, src):
        outs = [np.empty(s) for s in self.sizes]
        arr = (c_dp * self.stages)(*[dp(o) for o in outs])
        lib().ref_multigrid_project(self.h, dp(src), arr)
        return outs

    def set_chi(self, chi):
        lib().ref_multigrid_set_chi(self.h, dp(chi))

    def solve(self, x, b, eps):
        eps = np.ascontiguousarray(eps, dtype=np.float64)
        num = np.zeros(self.stages, dtype=np.int32)
        sec = C.c_double(0)
        st = lib().ref_multigrid_solve(self.h, dp(x), dp(b), dp(eps), ip(num), C.byref(sec))
        return st, [int(v) for v in num], sec.value

    def __del__(self):
        try:
            lib().ref_multigrid_free(self.h)
        except Exception:
            pass


def elliptic3d_symv(g, cylindrical, direction, jfactor, chi_weight_jump, chi, alpha, x, beta, y):
    """dg::Elliptic3d with set_compute_in_2d(true) (elliptic.h:557-797); g: RefGrid (ndim 3); returns (y, weights, precond)"""
    n = x.size
    out, w, p = np.array(y, copy=True), np.empty(n), np.empty(n)
    lib().ref_elliptic3d_symv(C.byref(g), int(cylindrical), int(direction), C.c_double(jfactor), int(chi_weight_jump),
                              dp(chi) if chi is not None else None, C.c_double(alpha), dp(np.ascontiguousarray(x)),
                              C.c_double(beta), dp(out), dp(w), dp(p))
    return out, w, p


def elliptic3d_symv_mode(g, cylindrical, direction, jfactor, chi_weight_jump, compute_in_2d, chi, alpha, x, beta, y, variation=False):
    """dg::Elliptic3d in its full 3-d mode (compute_in_2d False: elliptic.h:688-697,727-746) or restricted to the planes;
    returns y, or (y, variation(x)) with variation=True"""
    n = x.size
    out = np.array(y, copy=True)
    var = np.empty(n) if variation else None
    lib().ref_elliptic3d_symv_mode(C.byref(g), int(cylindrical), int(direction), C.c_double(jfactor), int(chi_weight_jump),
                                   int(compute_in_2d), dp(chi) if chi is not None else None, C.c_double(alpha),
                                   dp(np.ascontiguousarray(x)), C.c_double(beta), dp(out), dp(var) if variation else None)
    return (out, var) if variation else out


def elliptic1d_symv(g, bcx, direction, jfactor, chi, alpha, x, beta, y):
    """dg::Elliptic1d (elliptic.h:65-200); g: RefGrid (ndim 1); returns (y, weights, precond)"""
    n = x.size
    out, w, p = np.array(y, copy=True), np.empty(n), np.empty(n)
    lib().ref_elliptic1d_symv(C.byref(g), int(bcx), int(direction), C.c_double(jfactor), dp(chi) if chi is not None else None,
                              C.c_double(alpha), dp(np.ascontiguousarray(x)), C.c_double(beta), dp(out), dp(w), dp(p))
    return out, w, p
