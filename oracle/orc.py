"""TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/libdgoracle.so (plain-C restatement of the
reference algorithms, oracle/dgoracle.c).  Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module; nothing under feltor_b200/ does."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdgoracle.so")
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_lp = C.POINTER(C.c_int64)
BIN_COUNT = 39


def build():
    src = os.path.join(_HERE, "dgoracle.c")
    if not os.path.exists(LIB_PATH) or os.path.getmtime(src) > os.path.getmtime(LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "libdgoracle.so"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_round.restype = C.c_double
        _lib.orc_dot2.restype = C.c_double
        _lib.orc_dot3.restype = C.c_double
    return _lib


def dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_ip)


def lp(a):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_lp)


def d(v):
    return C.c_double(v)


# ---------------------------------------------------------------- blas1
def copy(x, y): lib().orc_copy(x.size, dp(x), dp(y))
def scal(x, a): lib().orc_scal(x.size, dp(x), d(a))
def plus(x, a): lib().orc_plus(x.size, dp(x), d(a))
def axpby(a, x, b, y): lib().orc_axpby(x.size, d(a), dp(x), d(b), dp(y))
def axpbyz(a, x, b, y, z): lib().orc_axpbyz(x.size, d(a), dp(x), d(b), dp(y), dp(z))
def axpbypgz(a, x, b, y, g, z): lib().orc_axpbypgz(x.size, d(a), dp(x), d(b), dp(y), d(g), dp(z))
def pointwiseDot(a, x, y, b, z): lib().orc_pointwiseDot(x.size, d(a), dp(x), dp(y), d(b), dp(z))
def axypby(a, x, b, y): lib().orc_axypby(x.size, d(a), dp(x), d(b), dp(y))
def pointwiseDot_xy(x, y, z): lib().orc_pointwiseDot_xy(x.size, dp(x), dp(y), dp(z))
def pointwiseDot3(a, x1, x2, x3, b, y): lib().orc_pointwiseDot3(x1.size, d(a), dp(x1), dp(x2), dp(x3), d(b), dp(y))


def pointwiseDot2(a, x1, y1, b, x2, y2, g, z):
    lib().orc_pointwiseDot2(z.size, d(a), dp(x1), dp(y1), d(b), dp(x2), dp(y2), d(g), dp(z))


def pointwiseDivide(a, x, y, b, z): lib().orc_pointwiseDivide(x.size, d(a), dp(x), dp(y), d(b), dp(z))
def pointwiseDivide_xy(x, y, z): lib().orc_pointwiseDivide_xy(x.size, dp(x), dp(y), dp(z))


def tensor_multiply2d(lam, t, in0, in1, mu, out0, out1):
    """lam: array or float; t = (t00,t01,t10,t11) arrays or None."""
    larr = lam if isinstance(lam, np.ndarray) else None
    ls = 1.0 if larr is not None else float(lam)
    t = t or (None, None, None, None)
    lib().orc_tensor_multiply2d(in0.size, dp(larr), d(ls), dp(t[0]), dp(t[1]), dp(t[2]), dp(t[3]), dp(in0), dp(in1),
                                d(mu), dp(out0), dp(out1))


def tensor_multiply3d(lam, t, ins, mu, outs):
    """lam: array or float; t = 9 arrays/None (row major) or None; ins, outs: 3 arrays each."""
    larr = lam if isinstance(lam, np.ndarray) else None
    ls = 1.0 if larr is not None else float(lam)
    T = (c_dp * 9)(*[dp(a) for a in (t or [None] * 9)])
    lib().orc_tensor_multiply3d(ins[0].size, dp(larr), d(ls), T, (c_dp * 3)(*[dp(a) for a in ins]), d(mu),
                                (c_dp * 3)(*[dp(a) for a in outs]))


def ds_apply(kind, alpha, a, b, c, G, bphi, delta, beta, g):
    """inc/geometries/ds.h:743-1000; G = (sqrtGm, sqrtG, sqrtGp) or None, bphi = (bphiM, bphi, bphiP)"""
    G = G or (None, None, None)
    lib().orc_ds_apply(kind, a.size, d(alpha), dp(a), dp(b), dp(c), dp(G[0]), dp(G[1]), dp(G[2]), dp(bphi[0]), dp(bphi[1]),
                       dp(bphi[2]), d(delta), d(beta), dp(g))


def assign_bc_along_field(order, neu, delta, fm, f, fp, hbm, hbp, bbm, bbo, bbp, bv, fmg, fpg):
    """inc/geometries/ds.h:169-296; order 2 (fm, f, fp) or 1 (fm, fp; f None); bv = (boundary value minus, plus)"""
    lib().orc_assign_bc_along_field(order, int(neu), fm.size, d(delta), dp(fm), dp(f), dp(fp), dp(hbm), dp(hbp), dp(bbm), dp(bbo),
                                    dp(bbp), d(bv[0]), d(bv[1]), dp(fmg), dp(fpg))


def csr_stencil(kind, pos, idx, val, alpha, x, y):
    """blas2::stencil with CSRMedianFilter (0) / CSRSWMFilter(alpha) (1) / CSRAverageFilter (2) / CSRSymvFilter (3) /
    CSRSlopeLimiter(alpha) (4; writes only the rows of y the limiter touches)"""
    lib().orc_csr_stencil(kind, len(pos) - 1, ip(pos), ip(idx), dp(val), d(alpha), dp(x), dp(y))


def embedded_pair_sum(y, yt, b0, bt0, b, bt, ks):
    b = np.ascontiguousarray(b, dtype=np.float64)
    bt = np.ascontiguousarray(bt, dtype=np.float64)
    arr = (c_dp * len(ks))(*[dp(k) for k in ks])
    lib().orc_embedded_pair_sum(y.size, dp(y), dp(yt), d(b0), d(bt0), len(ks), dp(b), dp(bt), arr)


# ---------------------------------------------------------------- exblas
def exdot2(x, y):
    acc = np.zeros(BIN_COUNT, dtype=np.int64)
    st = lib().orc_exdot2(x.size, dp(x), dp(y), lp(acc))
    return acc, st


def exdot3(x, w, y):
    acc = np.zeros(BIN_COUNT, dtype=np.int64)
    st = lib().orc_exdot3(x.size, dp(x), dp(w), dp(y), lp(acc))
    return acc, st


def normalize(acc):
    a = np.array(acc, dtype=np.int64)
    lib().orc_normalize(lp(a))
    return a


def round_acc(acc):
    a = np.array(acc, dtype=np.int64)
    return float(lib().orc_round(lp(a)))


def dot2(x, y):
    st = C.c_int(0)
    v = lib().orc_dot2(x.size, dp(x), dp(y), C.byref(st))
    return float(v), st.value


def dot3(x, w, y):
    st = C.c_int(0)
    v = lib().orc_dot3(x.size, dp(x), dp(w), dp(y), C.byref(st))
    return float(v), st.value


def superacc_add(acc, other):
    a = np.array(acc, dtype=np.int64)
    lib().orc_superacc_add(lp(a), lp(np.ascontiguousarray(other, dtype=np.int64)))
    return a


# ---------------------------------------------------------------- sparse
def ell_symv(m, alpha, x, beta, y):
    """m: any object with meta()/data/cols_idx/data_idx (feltor_b200.topology.Ell or refwrap.Ell)."""
    meta = np.ascontiguousarray(m.meta(), dtype=np.int32)
    lib().orc_ell_symv(ip(meta), dp(m.data), ip(m.cols_idx), ip(m.data_idx), d(alpha), dp(x), d(beta), dp(y))


def coo_symv(meta, data, rows_idx, cols_idx, data_idx, alpha, xs, y):
    meta = np.ascontiguousarray(meta, dtype=np.int32)
    arr = (c_dp * len(xs))(*[dp(v) for v in xs])
    lib().orc_coo_symv(ip(meta), dp(data), ip(rows_idx), ip(cols_idx), ip(data_idx), d(alpha), arr, dp(y))


def csr_spmv(pos, idx, val, alpha, x, beta, y):
    lib().orc_csr_spmv(len(pos) - 1, ip(pos), ip(idx), dp(val), d(alpha), dp(x), d(beta), dp(y))


class _OrcEll(C.Structure):
    _fields_ = [("meta", c_ip), ("data", c_dp), ("cols", c_ip), ("didx", c_ip)]


class _OrcElliptic(C.Structure):
    _fields_ = [("leftx", _OrcEll), ("lefty", _OrcEll), ("rightx", _OrcEll), ("righty", _OrcEll),
                ("jumpx", _OrcEll), ("jumpy", _OrcEll), ("sigma", c_dp), ("vol", c_dp), ("chi_xx", c_dp),
                ("chi_xy", c_dp), ("chi_yx", c_dp), ("chi_yy", c_dp), ("jfactor", C.c_double),
                ("chi_weight_jump", C.c_int), ("size", C.c_int)]


class Elliptic2d:
    """Oracle Elliptic2d: mats = dict(leftx, lefty, rightx, righty, jumpx, jumpy) of Ell-like objects."""

    def __init__(self, mats, sigma, vol=None, chi=None, jfactor=1.0, chi_weight_jump=False):
        self._keep = []
        self.s = _OrcElliptic()
        for name in ("leftx", "lefty", "rightx", "righty", "jumpx", "jumpy"):
            m = mats[name]
            meta = np.ascontiguousarray(m.meta(), dtype=np.int32)
            self._keep += [meta, m.data, m.cols_idx, m.data_idx]
            setattr(self.s, name, _OrcEll(ip(meta), dp(m.data), ip(m.cols_idx), ip(m.data_idx)))
        self.size = sigma.size
        self.sigma = np.ascontiguousarray(sigma)
        self.vol = vol
        self.chi = chi or (None, None, None, None)
        self.s.sigma = dp(self.sigma)
        self.s.vol = dp(vol)
        self.s.chi_xx, self.s.chi_xy, self.s.chi_yx, self.s.chi_yy = [dp(c) for c in self.chi]
        self.s.jfactor = jfactor
        self.s.chi_weight_jump = int(chi_weight_jump)
        self.s.size = self.size
        self.work = np.empty(6 * self.size)

    def symv(self, alpha, x, beta, y):
        lib().orc_elliptic2d_symv(C.byref(self.s), d(alpha), dp(x), d(beta), dp(y), dp(self.work))

    def pcg_solve(self, x, b, P, W, eps, nrmb_correction=1.0, test_frequency=1, max_iter=None, residuals=None):
        return lib().orc_pcg_solve_elliptic2d(C.byref(self.s), dp(x), dp(b), dp(P), dp(W), d(eps), d(nrmb_correction),
                                              test_frequency, max_iter or self.size, dp(self.work), dp(residuals))


def spgemm(B_rows, C_cols, B, Cm):
    """dg::detail::spgemm_cpu_kernel (sparsematrix_cpu.h:19-95): B, Cm = (pos, idx, val) -> (pos, idx, val) of B Cm"""
    L = lib()
    L.orc_spgemm.restype = C.c_longlong
    bp, bi, bv = (np.ascontiguousarray(B[0], dtype=np.int32), np.ascontiguousarray(B[1], dtype=np.int32), np.ascontiguousarray(B[2], dtype=np.float64))
    cp, ci, cv = (np.ascontiguousarray(Cm[0], dtype=np.int32), np.ascontiguousarray(Cm[1], dtype=np.int32), np.ascontiguousarray(Cm[2], dtype=np.float64))
    pos = np.zeros(B_rows + 1, dtype=np.int32)
    nnz = L.orc_spgemm(B_rows, C_cols, ip(bp), ip(bi), dp(bv), ip(cp), ip(ci), dp(cv), ip(pos), None, None)
    idx, val = np.zeros(max(nnz, 1), dtype=np.int32), np.zeros(max(nnz, 1))
    L.orc_spgemm(B_rows, C_cols, ip(bp), ip(bi), dp(bv), ip(cp), ip(ci), dp(cv), ip(pos), ip(idx), dp(val))
    return pos, idx[:nnz], val[:nnz]

