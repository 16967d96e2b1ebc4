"""dg::blas1-shaped adapters over (a) the C oracle, (b) the wrapped reference, so that one test body
(tests/util.blas1_sequence and the randomized parity cases) runs unchanged on oracle, reference and GPU."""
import ctypes as C
import math
import numpy as np
from oracle import orc, refwrap


class OracleBlas1:
    """call-site shortcuts of inc/dg/blas1.h + functor arithmetic of oracle/dgoracle.c"""

    def copy(self, x, y): orc.copy(x, y)

    def scal(self, x, a):
        if a != 1.0:
            orc.scal(x, a)

    def plus(self, x, a):
        if a != 0.0:
            orc.plus(x, a)

    def axpby(self, a, x, b, y, z=None):
        if z is not None:
            return orc.axpbyz(a, x, b, y, z)
        if a == 0.0:
            return self.scal(y, b)
        if x is y:
            return self.scal(y, a + b)
        orc.axpby(a, x, b, y)

    def axpbypgz(self, a, x, b, y, g, z):
        if a == 0.0:
            return self.axpby(b, y, g, z)
        if b == 0.0:
            return self.axpby(a, x, g, z)
        if x is y:
            return self.axpby(a + b, x, g, z)
        if x is z:
            return self.axpby(b, y, a + g, z)
        if y is z:
            return self.axpby(a, x, b + g, z)
        orc.axpbypgz(a, x, b, y, g, z)

    def pointwiseDot(self, *a):
        if len(a) == 3:
            return orc.pointwiseDot_xy(*a)
        if len(a) == 5:
            al, x1, x2, be, y = a
            if al == 0.0:
                return self.scal(y, be)
            if x1 is y:
                return orc.axypby(al, x2, be, y)
            if x2 is y:
                return orc.axypby(al, x1, be, y)
            return orc.pointwiseDot(al, x1, x2, be, y)
        if len(a) == 6:
            if a[0] == 0.0:
                return self.scal(a[5], a[4])
            return orc.pointwiseDot3(*a)
        al, x1, y1, be, x2, y2, ga, z = a
        if al == 0.0:
            return self.pointwiseDot(be, x2, y2, ga, z)
        if be == 0.0:
            return self.pointwiseDot(al, x1, y1, ga, z)
        return orc.pointwiseDot2(*a)

    def pointwiseDivide(self, *a):
        if len(a) == 3:
            return orc.pointwiseDivide_xy(*a)
        al, x1, x2, be, y = a
        if al == 0.0:
            return self.scal(y, be)
        if x1 is y:
            return orc.lib().orc_pointwiseDivide_alias(y.size, orc.d(al), orc.dp(x2), orc.d(be), orc.dp(y))
        return orc.pointwiseDivide(al, x1, x2, be, y)

    def transform(self, x, y, op):
        f = {"exp": math.exp, "ln": math.log, "sqrt": math.sqrt, "invert": lambda v: 1. / v, "abs": abs,
             "square": lambda v: v * v, "invsqrt": lambda v: 1. / math.sqrt(v)}[op]
        for i in range(x.size):
            y[i] = f(float(x[i]))


class RefBlas1:
    """the unmodified reference's dg::blas1 through oracle/_ref/libdgref.so"""

    def __init__(self, lib=None):
        self.l = lib if lib is not None else refwrap.lib()
        self.d, self.p = C.c_double, refwrap.dp

    def copy(self, x, y): self.l.ref_copy(x.size, self.p(x), self.p(y))
    def scal(self, x, a): self.l.ref_scal(x.size, self.p(x), self.d(a))
    def plus(self, x, a): self.l.ref_plus(x.size, self.p(x), self.d(a))

    def axpby(self, a, x, b, y, z=None):
        if z is None:
            self.l.ref_axpby(x.size, self.d(a), self.p(x), self.d(b), self.p(y))
        else:
            self.l.ref_axpbyz(x.size, self.d(a), self.p(x), self.d(b), self.p(y), self.p(z))

    def axpbypgz(self, a, x, b, y, g, z):
        self.l.ref_axpbypgz(x.size, self.d(a), self.p(x), self.d(b), self.p(y), self.d(g), self.p(z))

    def pointwiseDot(self, *a):
        d, p = self.d, self.p
        if len(a) == 3:
            self.l.ref_pointwiseDot_xy(a[0].size, p(a[0]), p(a[1]), p(a[2]))
        elif len(a) == 5:
            self.l.ref_pointwiseDot(a[1].size, d(a[0]), p(a[1]), p(a[2]), d(a[3]), p(a[4]))
        elif len(a) == 6:
            self.l.ref_pointwiseDot3(a[1].size, d(a[0]), p(a[1]), p(a[2]), p(a[3]), d(a[4]), p(a[5]))
        else:
            self.l.ref_pointwiseDot2(a[1].size, d(a[0]), p(a[1]), p(a[2]), d(a[3]), p(a[4]), p(a[5]), d(a[6]), p(a[7]))

    def pointwiseDivide(self, *a):
        d, p = self.d, self.p
        if len(a) == 3:
            self.l.ref_pointwiseDivide_xy(a[0].size, p(a[0]), p(a[1]), p(a[2]))
        else:
            self.l.ref_pointwiseDivide(a[1].size, d(a[0]), p(a[1]), p(a[2]), d(a[3]), p(a[4]))

    def transform(self, x, y, op):
        assert op == "exp"
        self.l.ref_transform_exp(x.size, self.p(x), self.p(y))


def np_make(a):
    return np.array(a, dtype=np.float64)


def np_get(a):
    return a
