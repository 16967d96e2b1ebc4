"""dg::blas1 on device tensors.  Names, argument order and the call-site shortcuts follow inc/dg/blas1.h of the
reference (file:line cited per function); the arithmetic is done by libdgb200.so (feltor_b200/csrc/blas1.cu)."""
import ctypes as C
from ._lib import lib
from ._dev import ptr, stream

d = C.c_double


def _n(*ts):
    n = ts[0].numel()
    for t in ts:
        if t.numel() != n:
            raise ValueError("blas1: vector sizes differ")  # dg::Error in blas1_dispatch_shared.h:113
    return n


def copy(x, y):
    """blas1.h:243: y = x (x may be a scalar)."""
    if isinstance(x, (int, float)):
        lib().fill(y.numel(), d(x), ptr(y), stream())
    else:
        lib().copy(_n(x, y), ptr(x), ptr(y), stream())


def scal(x, alpha):
    """blas1.h:324"""
    if alpha == 1.0:
        return
    lib().scal(x.numel(), ptr(x), d(alpha), stream())


def plus(x, alpha):
    """blas1.h:341"""
    if alpha == 0.0:
        return
    lib().plus(x.numel(), ptr(x), d(alpha), stream())


def axpby(alpha, x, beta, y, z=None):
    """blas1.h:306-317 (2-vector) and blas1.h:373-386 (3-vector) including their shortcuts."""
    if z is None:
        if alpha == 0.0:
            return scal(y, beta)
        if x.data_ptr() == y.data_ptr():
            return scal(y, alpha + beta)
        lib().axpby(_n(x, y), d(alpha), ptr(x), d(beta), ptr(y), stream())
    else:
        lib().axpbyz(_n(x, y, z), d(alpha), ptr(x), d(beta), ptr(y), ptr(z), stream())


def axpbypgz(alpha, x, beta, y, gamma, z):
    """blas1.h:339-363"""
    if alpha == 0.0:
        return axpby(beta, y, gamma, z)
    if beta == 0.0:
        return axpby(alpha, x, gamma, z)
    if x.data_ptr() == y.data_ptr():
        return axpby(alpha + beta, x, gamma, z)
    if x.data_ptr() == z.data_ptr():
        return axpby(beta, y, alpha + gamma, z)
    if y.data_ptr() == z.data_ptr():
        return axpby(alpha, x, beta + gamma, z)
    lib().axpbypgz(_n(x, y, z), d(alpha), ptr(x), d(beta), ptr(y), d(gamma), ptr(z), stream())


def pointwiseDot(*a):
    """blas1.h:405-479: (alpha,x1,x2,beta,y) | (x1,x2,y) | (alpha,x1,x2,x3,beta,y) | (alpha,x1,y1,beta,x2,y2,gamma,z)"""
    if len(a) == 3:
        x1, x2, y = a
        lib().pointwise_dot_xy(_n(x1, x2, y), ptr(x1), ptr(x2), ptr(y), stream())
    elif len(a) == 5:
        alpha, x1, x2, beta, y = a
        if alpha == 0.0:
            return scal(y, beta)
        lib().pointwise_dot(_n(x1, x2, y), d(alpha), ptr(x1), ptr(x2), d(beta), ptr(y), stream())
    elif len(a) == 6:
        alpha, x1, x2, x3, beta, y = a
        if alpha == 0.0:
            return scal(y, beta)
        lib().pointwise_dot3(_n(x1, x2, x3, y), d(alpha), ptr(x1), ptr(x2), ptr(x3), d(beta), ptr(y), stream())
    elif len(a) == 8:
        alpha, x1, y1, beta, x2, y2, gamma, z = a
        if alpha == 0.0:
            return pointwiseDot(beta, x2, y2, gamma, z)
        if beta == 0.0:
            return pointwiseDot(alpha, x1, y1, gamma, z)
        lib().pointwise_dot2(_n(x1, y1, x2, y2, z), d(alpha), ptr(x1), ptr(y1), d(beta), ptr(x2), ptr(y2), d(gamma),
                             ptr(z), stream())
    else:
        raise TypeError("pointwiseDot: unsupported signature")


def pointwiseDivide(*a):
    """blas1.h:493-528: (alpha,x1,x2,beta,y) | (x1,x2,y)"""
    if len(a) == 3:
        x1, x2, y = a
        lib().pointwise_divide_xy(_n(x1, x2, y), ptr(x1), ptr(x2), ptr(y), stream())
    else:
        alpha, x1, x2, beta, y = a
        if alpha == 0.0:
            return scal(y, beta)
        lib().pointwise_divide(_n(x1, x2, y), d(alpha), ptr(x1), ptr(x2), d(beta), ptr(y), stream())


def tensor_multiply2d(lam, t, in0, in1, mu, out0, out1):
    """dg::tensor::multiply2d, inc/dg/topology/multiply.h:215; lam: tensor or float, t = 4 tensors/None or None."""
    t = t or (None, None, None, None)
    larr = None if isinstance(lam, (int, float)) else lam
    ls = float(lam) if larr is None else 1.0
    lib().tensor_multiply2d(_n(in0, in1, out0, out1), ptr(larr), d(ls), ptr(t[0]), ptr(t[1]), ptr(t[2]), ptr(t[3]),
                            ptr(in0), ptr(in1), d(mu), ptr(out0), ptr(out1), stream())


def tensor_multiply3d(lam, t, ins, mu, outs):
    """dg::tensor::multiply3d, inc/dg/topology/multiply.h:34-58,243; t = 9 tensors/None (row major) or None."""
    larr = None if isinstance(lam, (int, float)) else lam
    ls = float(lam) if larr is None else 1.0
    T = (C.c_void_p * 9)(*[None if a is None else a.data_ptr() for a in (t or [None] * 9)])
    I = (C.c_void_p * 3)(*[a.data_ptr() for a in ins])
    O = (C.c_void_p * 3)(*[a.data_ptr() for a in outs])
    lib().tensor_multiply3d(_n(*ins, *outs), ptr(larr), d(ls), T, I, d(mu), O, stream())


def embedded_pair_sum(y, yt, b0, bt0, b, bt, ks):
    """subroutines.h:179-204 as used by ERKStep (runge_kutta.h:35-62)."""
    nk = len(ks)
    B = (C.c_double * nk)(*b)
    BT = (C.c_double * nk)(*bt)
    K = (C.c_void_p * nk)(*[k.data_ptr() for k in ks])
    lib().embedded_pair_sum(_n(y, yt, *ks), ptr(y), ptr(yt), d(b0), d(bt0), nk, B, BT, K, stream())


OPS = {"exp": 0, "ln": 1, "sqrt": 2, "invert": 3, "abs": 4, "square": 5, "invsqrt": 6}


def transform(x, y, op):
    """blas1.h:585 with a functor of inc/dg/functors.h named by `op`."""
    lib().transform(_n(x, y), OPS[op], ptr(x), ptr(y), stream())


REDUCE = {"sum": 0, "max": 1, "min": 2, "or": 3}
UNARY = {"identity": 0, "abs": 1, "square": 2, "isnan": 3, "isnotfinite": 4}


def reduce(x, init, op, unary="identity"):
    """blas1.h:213-223 for the closed set of (binary, unary) functors of include/dgb200.h"""
    out = C.c_double()
    lib().reduce(x.numel(), ptr(x), REDUCE[op], UNARY[unary], d(init), C.byref(out), stream())
    return out.value
