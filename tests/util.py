"""Shared helpers of the parity tests."""
import numpy as np


def bits(a):
    """int64 bit pattern(s) of double(s) -- the representation the reference's golden tests compare (blas1_t.cpp:41)"""
    return np.array(a, dtype=np.float64, copy=True).view(np.int64)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


def rng(seed=42):
    return np.random.default_rng(seed)


def wide(r, n, lo=-30, hi=30):
    """random doubles with a wide dynamic range (exercises many superaccumulator words)"""
    return r.uniform(-1, 1, n) * np.exp2(r.integers(lo, hi, n).astype(np.float64))


# --------------------------------------------------------------------------------------------------
# the reference's own blas1 known-answer sequence (inc/dg/blas1_t.cpp:102-184), backend-agnostic:
# `B` provides the dg::blas1 functions, make(v) creates a backend vector from numpy, get(v) reads it back.
BLAS1_GOLDEN = [
    ("copy", 4617316080911554445), ("scal", 4474825110624711575), ("plus", 4476275821608249130),
    ("fma", 4633360230582305548), ("axpby", 4408573477492505937), ("axpbyz", 4468869610430797025),
    ("axpbypgz", 4617320336812948958), ("pDot", 4413077932784031586), ("pDot_ab", 4556605413983777388),
    ("pDot2", 4601058031075598447), ("pDot3", 4550507856334720009), ("pDivide", 4810082017219139146),
    ("pDivide_ab", 4820274520177585116), ("exp", 4620007020034741378),
]


def blas1_sequence(B, make, get, n=500):
    """returns {name: int64 bit pattern array} following blas1_t.cpp:102-176 step by step"""
    out = {}
    v1, v2, v3, v4 = (make(np.full(n, c)) for c in (2.0002, 3.00003, 5.0005, 4.00004))
    nan = np.full(n, 5.0005)
    nan[0] = np.nan
    res = make(nan)
    B.copy(v3, res)
    out["copy"] = bits(get(res))
    B.scal(v3, 3e-10)
    out["scal"] = bits(get(v3))
    B.plus(v3, 3e-10)
    out["plus"] = bits(get(v3))
    B.axpby(3e+10, v3, 1., v4)
    out["fma"] = bits(get(v4))
    B.axpby(3e-10, v1, -2e-10, v2)
    out["axpby"] = bits(get(v2))
    v5 = make(nan)
    B.axpby(3e-10, v1, -2., v2, v5)
    out["axpbyz"] = bits(get(v5))
    B.axpbypgz(2.5, v1, 7.e+10, v2, -0.125, v3)
    out["axpbypgz"] = bits(get(v3))
    v3 = make(nan)
    B.pointwiseDot(v1, v2, v3)
    out["pDot"] = bits(get(v3))
    B.pointwiseDot(0.2, v1, v2, +0.4e10, v3)
    out["pDot_ab"] = bits(get(v3))
    v5 = make(np.full(n, 4.00004))
    B.pointwiseDot(-0.2, v1, v2, 0.4, v3, v4, 0.1, v5)
    out["pDot2"] = bits(get(v5))
    B.pointwiseDot(0.2, v1, v2, v4, 0.4, v3)
    out["pDot3"] = bits(get(v3))
    v5 = make(nan)
    B.pointwiseDivide(v1, v2, v5)
    out["pDivide"] = bits(get(v5))
    B.pointwiseDivide(5., v1, v2, -1., v3)
    out["pDivide_ab"] = bits(get(v3))
    v3 = make(nan)
    B.transform(v1, v3, "exp")
    out["exp"] = bits(get(v3))
    return out
