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


def elliptic3d_full_symv(B, T, x0, x1, N, bc, direction, jfactor, cwj, cyl, chi, alpha, x, beta, y, compute_in_2d=False):
    """dg::Elliptic3d::symv of the reference (inc/dg/elliptic.h:680-749) in its FULL 3-d mode, restated call by
    call on a backend B of dg-shaped primitives -- the C oracle on numpy arrays (tests/test_elliptic3d_oracle.py, pinned against
    the live reference and committed fixtures) or the C ABI of libdgb200.so on device vectors (tests/test_gpu_elliptic3d.py).
    B: make(np) -> vector, symv(m, a, x, b, y), tensor_multiply3d / tensor_multiply2d, pointwiseDot, pointwiseDivide, axpby,
    axpbypgz.  Grids: n = 3 in x and y, 1 in z => g.nz() == 1: centered dz on both sides and NO jump term in z (elliptic.h:606-611)."""
    g = T.Grid(x0, x1, [3, 3, 1], N, bc)
    n = g.size
    inv_d = T.inverse_dir(direction)
    leftx, lefty = T.derivative(0, g, T.inverse_bc(bc[0]), inv_d), T.derivative(1, g, T.inverse_bc(bc[1]), inv_d)
    rightx, righty = T.derivative(0, g, bc[0], direction), T.derivative(1, g, bc[1], direction)
    jumpx, jumpy = T.jump(0, g, bc[0]), T.jump(1, g, bc[1])
    rightz, leftz = T.derivative(2, g, bc[2], T.CENTERED), T.derivative(2, g, T.inverse_bc(bc[2]), T.CENTERED)
    ones, zeros = np.ones(n), np.zeros(n)
    if cyl:   # CylindricalGrid3d: g^pp = 1/R/R (base_geometry.h:336-344), vol = 1/sqrt(det) (multiply.h:389)
        R = np.ascontiguousarray(np.broadcast_to(g.abscissas(0), (g.shape(2), g.shape(1), g.shape(0))).reshape(-1))
        gpp = (1. / R) / R
        vol = 1. / np.sqrt(gpp)
    else:
        gpp, vol = ones, ones.copy()
    # the metric as SparseTensor: value 0 = zeros, 1 = ones (tensor.h): every entry is a vector, the arithmetic is carried out
    t9 = [B.make(v) for v in (ones, zeros, zeros, zeros, ones, zeros, zeros, zeros, gpp)]
    t4 = [t9[0], t9[1], t9[3], t9[4]]
    vold = B.make(vol)
    sigma = B.make(np.zeros(n))
    B.pointwiseDot(chi, vold, sigma)                                  # set_chi: elliptic.h:638
    tx, ty, tz, temp = (B.make(np.zeros(n)) for _ in range(4))
    B.symv(rightx, 1., x, 0., tx)
    B.symv(righty, 1., x, 0., ty)
    if not compute_in_2d:
        B.symv(rightz, 1., x, 0., tz)
        B.tensor_multiply3d(sigma, t9, (tx, ty, tz), 0., (tx, ty, tz))
        B.symv(leftz, -1., tz, 0., temp)
        B.symv(lefty, -1., ty, 1., temp)
    else:
        B.tensor_multiply2d(sigma, t4, tx, ty, 0., tx, ty)
        B.symv(lefty, -1., ty, 0., temp)
    B.symv(leftx, -1., tx, 1., temp)
    if jfactor != 0:
        if cwj:
            B.symv(jumpx, jfactor, x, 0., tx)
            B.symv(jumpy, jfactor, x, 0., ty)
            B.tensor_multiply2d(sigma, t4, tx, ty, 0., tx, ty)
            B.axpbypgz(1., tx, 1., ty, 1., temp)
        else:
            B.symv(jumpx, jfactor, x, 1., temp)
            B.symv(jumpy, jfactor, x, 1., temp)
    B.pointwiseDivide(alpha, temp, vold, beta, y)
