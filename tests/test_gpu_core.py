"""GPU parity (through the C ABI): blas1, exblas dot, Ell/Coo symv vs the oracle, the reference's golden vectors
and the committed reference fixtures.  Bit-exact everywhere."""
import numpy as np
import pytest
from oracle import orc
from util import bits, same_bits, blas1_sequence, BLAS1_GOLDEN, rng, wide
from backends import OracleBlas1
import kat

pytestmark = pytest.mark.gpu
OB = OracleBlas1()


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


def test_blas1_reference_goldens(G):
    from feltor_b200 import blas1
    out = blas1_sequence(blas1, G.make, G.get)
    for name, g in BLAS1_GOLDEN:
        if name == "exp":  # device libm vs host libm: the reference pins exp per compiler only (evaluation_t.cpp:36-40)
            assert np.all(np.abs(out[name] - g) <= 2), name
        else:
            assert np.all(out[name] == g), name


@pytest.mark.parametrize("n", [1, 2, 3, 255, 1031, 4096, 100003])
def test_blas1_vs_oracle(G, n):
    from feltor_b200 import blas1
    r = rng(n)
    v = [r.uniform(-2, 2, n) for _ in range(5)]
    v[1][np.abs(v[1]) < 1e-3] = 0.5
    ops = [
        lambda B, w: B.copy(w[0], w[1]),
        lambda B, w: B.scal(w[0], 0.37),
        lambda B, w: B.plus(w[0], -0.37),
        lambda B, w: B.axpby(0.7, w[0], -1.3, w[1]),
        lambda B, w: B.axpby(0.7, w[0], 0., w[1]),
        lambda B, w: B.axpby(0.7, w[0], -1.3, w[1], w[2]),
        lambda B, w: B.axpby(0.7, w[0], -1.3, w[1], w[0]),  # output aliases input
        lambda B, w: B.axpbypgz(0.7, w[0], -1.3, w[1], 0.4, w[2]),
        lambda B, w: B.pointwiseDot(w[0], w[1], w[2]),
        lambda B, w: B.pointwiseDot(0.7, w[0], w[1], -1.3, w[2]),
        lambda B, w: B.pointwiseDot(0.7, w[0], w[1], -1.3, w[1]),
        lambda B, w: B.pointwiseDot(0.7, w[0], w[1], w[2], -1.3, w[3]),
        lambda B, w: B.pointwiseDot(0.7, w[0], w[1], -1.3, w[2], w[3], 0.4, w[4]),
        lambda B, w: B.pointwiseDivide(w[0], w[1], w[2]),
        lambda B, w: B.pointwiseDivide(0.7, w[0], w[1], -1.3, w[2]),
        lambda B, w: B.pointwiseDivide(0.7, w[2], w[1], -1.3, w[2]),
    ]
    for k, f in enumerate(ops):
        a = [u.copy() for u in v]
        f(OB, a)
        b = [G.make(u) for u in v]
        f(blas1, b)
        assert same_bits(np.stack(a), np.stack([G.get(t) for t in b])), ("op", k, n)


@pytest.mark.parametrize("kind", ["median", "swm", "average", "symv"])
def test_csr_stencil(G, kind):
    """blas2::stencil with the library's CSR filters: fixture of the unmodified reference + random stencils vs the oracle"""
    import os
    from feltor_b200 import blas2, DgbError
    from feltor_b200._dev import dvec
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ds_golden.npz"))
    k = blas2.STENCILS[kind]
    pos, idx, val = dvec(gold["stencil/pos"]), dvec(gold["stencil/idx"]), dvec(gold["stencil/val"])
    x, y = G.make(gold["stencil/x"]), G.make(np.full(gold["stencil/x"].size, np.nan))
    blas2.stencil(kind, pos, idx, val, x, y, alpha=1.5)
    assert same_bits(G.get(y), gold[f"stencil/kind{k}"])
    r = rng(50 + k)
    n = 3001
    counts = r.integers(1, 30, n)
    hp = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    hi = np.concatenate([np.clip(i + r.integers(-40, 41, c), 0, n - 1) for i, c in enumerate(counts)]).astype(np.int32)
    hv, hx = r.uniform(-1, 1, hp[-1]), np.round(r.uniform(-3, 3, n), 2) + 0.0
    want = np.zeros(n)
    orc.csr_stencil(k, hp, hi, hv, 0.8, hx, want)
    dx, dy = G.make(hx), G.make(np.full(n, np.nan))
    dp_, di_, dv_ = dvec(hp), dvec(hi), dvec(hv)
    blas2.stencil(kind, dp_, di_, dv_, dx, dy, alpha=0.8)
    assert same_bits(G.get(dy), want)
    with pytest.raises(DgbError):
        blas2.stencil(kind, dp_, di_, dv_, dx, dx, alpha=0.8)


def test_slope_limiter(G):
    """blas2::stencil( CSRSlopeLimiter( mod), limiter_stencil, x, y) (filter.h:288-336, stencil.h:89-256): golden vectors of the
    unmodified reference for every boundary condition, 1-d and 2-d, both directions; larger random cases against the oracle"""
    import os
    from feltor_b200 import blas2, topology as T
    from feltor_b200._dev import dvec
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "limiter_golden.npz"))
    names = sorted({k.split("/")[0] for k in gold.files})
    assert len(names) == 10
    for name in names:
        pos, idx, val, x = (gold[name + "/" + k] for k in ("pos", "idx", "val", "x"))
        for mod in (0., 0.3):
            y = G.make(np.full(x.size, np.nan))
            blas2.stencil("slope", dvec(pos), dvec(idx), dvec(val), G.make(x), y, alpha=mod)
            assert same_bits(G.get(y), gold[name + "/y_mod%g" % mod]), (name, mod)
    r = rng(77)
    for n, N, bc, direction in ((3, [200, 150], [1, 0], 0), (3, [200, 150], [1, 0], 1), (4, [64, 33], [2, 3], 1), (2, [31, 17], [4, 1], 0)):
        g = T.Grid([0., 0.], [1., 2.], n, N, bc)
        pos, idx, val = T.limiter_stencil(g, direction)
        t = np.linspace(0, 1, g.size)
        x = np.sin(40 * t) + (t > 0.3) * 1.1 + 0.2 * r.uniform(-1, 1, g.size)
        for mod in (0., 0.05):
            want = np.full(g.size, np.nan)
            orc.csr_stencil(4, pos, idx, val, mod, x, want)
            y = G.make(np.full(g.size, np.nan))
            blas2.stencil("slope", dvec(pos), dvec(idx), dvec(val), G.make(x), y, alpha=mod)
            assert same_bits(G.get(y), want), (n, N, bc, direction, mod)
            assert not np.isnan(want).any()


def test_tensor_multiply3d(G):
    """TensorMultiply3d (multiply.h:34-58): fixture of the unmodified reference, oracle, aliasing, identity, odd/unaligned"""
    import os
    from feltor_b200 import blas1
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ds_golden.npz"))
    t, ins, lam = [G.make(a) for a in gold["t3d/t"]], [G.make(a) for a in gold["t3d/in"]], G.make(gold["t3d/lambda"])
    o = [G.make(a) for a in gold["t3d/out0"]]
    blas1.tensor_multiply3d(lam, t, ins, 0.3, o)
    assert same_bits(np.stack([G.get(a) for a in o]), gold["t3d/out"])
    o = [G.make(a) for a in gold["t3d/in"]]
    blas1.tensor_multiply3d(lam, t, o, 0., o)
    assert same_bits(np.stack([G.get(a) for a in o]), gold["t3d/out_alias"])
    # sparse tensor (some components implicit), scalar lambda, operands offset by one element -> scalar path
    r = rng(31)
    n = 777
    hv = [r.uniform(-2, 2, n + 1) for _ in range(10)]
    dv = [G.make(a) for a in hv]
    ht = [hv[0][1:].copy(), None, hv[1][1:].copy(), None, None, hv[2][1:].copy(), hv[3][1:].copy(), None, None]
    dt = [dv[0][1:], None, dv[1][1:], None, None, dv[2][1:], dv[3][1:], None, None]
    hi, ho = [hv[4 + k][1:].copy() for k in range(3)], [hv[7 + k][1:].copy() for k in range(3)]
    orc.tensor_multiply3d(-1.5, ht, hi, 0.25, ho)
    do = [dv[7 + k][1:] for k in range(3)]
    blas1.tensor_multiply3d(-1.5, dt, [dv[4 + k][1:] for k in range(3)], 0.25, do)
    assert same_bits(np.stack(ho), np.stack([G.get(a) for a in do]))


def test_blas1_unaligned_views_and_tensor(G):
    """operands offset by one element (8-byte aligned only) take the scalar path; TensorMultiply2d; EmbeddedPairSum"""
    from feltor_b200 import blas1
    r = rng(3)
    n = 1001
    v = [r.uniform(-2, 2, n + 1) for _ in range(9)]
    dv = [G.make(u) for u in v]
    a = [u[1:].copy() for u in v]
    b = [t[1:] for t in dv]
    OB.axpbypgz(0.7, a[0], -1.3, a[1], 0.4, a[2])
    blas1.axpbypgz(0.7, b[0], -1.3, b[1], 0.4, b[2])
    assert same_bits(a[2], G.get(b[2]))
    orc.tensor_multiply2d(a[0], (a[1], a[2], a[3], a[4]), a[5], a[6], 0.3, a[7], a[8])
    blas1.tensor_multiply2d(b[0], (b[1], b[2], b[3], b[4]), b[5], b[6], 0.3, b[7], b[8])
    assert same_bits(a[7], G.get(b[7])) and same_bits(a[8], G.get(b[8]))
    # identity tensor, in-place (the Elliptic call: multiply2d(sigma, chi, tx, ty, 0., tx, ty), elliptic.h:435)
    a[7][5] = np.nan
    dv[7][6] = float("nan")
    orc.tensor_multiply2d(a[0], None, a[5], a[6], 0., a[5], a[6])
    blas1.tensor_multiply2d(b[0], None, b[5], b[6], 0., b[5], b[6])
    assert same_bits(a[5], G.get(b[5])) and same_bits(a[6], G.get(b[6]))
    ks = [a[1], a[2], a[3]]
    orc.embedded_pair_sum(a[7][:], a[8][:], 0.5, -0.25, [0.1, 0.2, 0.3], [0.3, 0.2, 0.1], ks)
    blas1.embedded_pair_sum(b[7], b[8], 0.5, -0.25, [0.1, 0.2, 0.3], [0.3, 0.2, 0.1], [b[1], b[2], b[3]])
    # a[7] contains a NaN at index 5 on both sides (same bits expected)
    assert same_bits(a[7], G.get(b[7])) and same_bits(a[8], G.get(b[8]))


def test_blas1_empty(G):
    import torch
    from feltor_b200 import blas1
    e = torch.zeros(0, dtype=torch.float64, device="cuda")
    blas1.axpby(1., e, 2., e.clone())
    assert G.dot2(e, e) == 0.0


# ------------------------------------------------------------------------------------------------ dot
@pytest.mark.parametrize("case", list(kat.evaluation_cases()), ids=lambda c: c[0])
def test_dot_reference_goldens(G, case):
    name, kind, ops, gold = case
    dv = [G.make(o) for o in ops]
    val = G.dot2(*dv) if kind == "dot2" else G.dot3(*dv)
    oval = orc.dot2(*ops)[0] if kind == "dot2" else orc.dot3(*ops)[0]
    assert same_bits([val], [oval])
    assert abs(int(bits([val])[0]) - gold) < 2, name


@pytest.mark.parametrize("name", ["small", "wide", "mid"])
def test_dot_fixtures(G, golden, name):
    from feltor_b200 import blas2
    x, w, y = (golden[f"dot/{name}/{k}"] for k in "xwy")
    dx, dw, dy = G.make(x), G.make(w), G.make(y)
    a2, v2, s2 = blas2.superacc(dx, dy)
    a3, v3, s3 = blas2.superacc(dx, dw, dy)
    assert s2 == 0 and s3 == 0
    assert np.array_equal(a2, golden[f"dot/{name}/acc2"]) and np.array_equal(a3, golden[f"dot/{name}/acc3"])
    assert same_bits([v2], golden[f"dot/{name}/val2"]) and same_bits([v3], golden[f"dot/{name}/val3"])


@pytest.mark.parametrize("n,lo,hi", [(1, -5, 5), (2, -5, 5), (31, -1000, 1000), (1000, -500, 500), (65537, -60, 60),
                                     (1 << 20, -20, 20), (3000001, -2, 2)])
def test_dot_vs_oracle_random(G, n, lo, hi):
    from feltor_b200 import blas2
    r = rng(n)
    x, w, y = wide(r, n, lo, hi), wide(r, n, lo // 4, hi // 4), wide(r, n, lo // 4, hi // 4)
    dx, dw, dy = G.make(x), G.make(w), G.make(y)
    a2, v2, s2 = blas2.superacc(dx, dy)
    a3, v3, s3 = blas2.superacc(dx, dw, dy)
    o2, os2 = orc.exdot2(x, y)
    o3, os3 = orc.exdot3(x, w, y)
    assert (s2, s3) == (os2, os3)
    if os2 == 0:
        assert np.array_equal(a2, o2) and same_bits([v2], [orc.round_acc(o2)])
    if os3 == 0:
        assert np.array_equal(a3, o3) and same_bits([v3], [orc.round_acc(o3)])


def test_dot_cancellation_and_scalars(G):
    """catastrophic cancellation is summed exactly; scalar operands (dot(1., v)); unaligned operands"""
    from feltor_b200 import blas2
    n = 100001
    r = rng(5)
    x = wide(r, n, -200, 200)
    xx = np.concatenate([x, -x, [3.0]])
    one = np.ones(xx.size)
    d = G.make(xx)
    assert G.dot2(d, G.make(one)) == 3.0
    a, v, s = blas2.superacc(1.0, d)
    assert v == 3.0 and s == 0
    a, v, s = blas2.superacc(d, 2.0, 0.5)
    assert v == 3.0
    dd = G.make(np.concatenate([[7.0], xx]))
    assert G.dot2(dd[1:], G.make(np.concatenate([[1.0], one]))[1:]) == 3.0


def test_dot_nonfinite(G):
    x = np.ones(1000)
    x[777] = np.nan
    with pytest.raises(FloatingPointError):
        G.dot2(G.make(x), G.make(x))
    big = G.make(np.full(4, 1e300))
    with pytest.raises(FloatingPointError):
        G.dot3(big, big, big)
    # and the workspace is usable afterwards
    one = G.make(np.ones(10))
    assert G.dot2(one, one) == 10.0


def test_dot_linearity_full_size(G):
    """size-independent property at the benchmark size (n=3, 1024^2): dot(x, y1) + dot(x, y2) accumulators add up
    EXACTLY to dot over the concatenation (integer superaccumulator algebra, mpi_accumulate.h:94-125)"""
    from feltor_b200 import blas2
    import torch
    n = 9 * 1024 * 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    y = torch.rand(n, dtype=torch.float64, device="cuda", generator=g) - 0.5
    h = n // 3 + 1
    a_all, v_all, _ = blas2.superacc(x, y)
    a1, _, _ = blas2.superacc(x[:h].contiguous(), y[:h].contiguous())
    a2, _, _ = blas2.superacc(x[h:].contiguous(), y[h:].contiguous())
    assert np.array_equal(orc.superacc_add(a1, a2), a_all)
    assert v_all == orc.round_acc(a_all)


# ------------------------------------------------------------------------------------------------ Ell / Coo symv
@pytest.mark.parametrize("case", list(kat.derivative_cases(three_d=True)), ids=lambda c: c[0])
def test_derivative_reference_goldens(G, case):
    from feltor_b200 import blas1
    got, gold, gh = kat.run_derivative_case(case, G.make, G.dot2, G.dot3, G.symv, blas1.pointwiseDot)
    ogot, _, _ = kat.run_derivative_case(case, np.array, lambda x, y: orc.dot2(x, y)[0], lambda x, w, y: orc.dot3(x, w, y)[0],
                                         lambda m, a, x, b, y: orc.ell_symv(m, a, x, b, y), orc.pointwiseDot_xy)
    assert got == ogot
    assert abs(got - gold) < 2 or abs(got - gh) < 2


def _ell_cases():
    from feltor_b200 import topology as T
    out = []
    for n in (1, 2, 3, 4, 5, 6):
        g = T.Grid([0, 0.1], [np.pi, 2 * np.pi + 0.1], n, [9, 7], [T.DIR, T.PER])
        for coord in (0, 1):
            for bc in (T.PER, T.DIR, T.NEU_DIR):
                for d in (T.FORWARD, T.BACKWARD, T.CENTERED):
                    out.append((f"n{n}-d{coord}-bc{bc}-dir{d}", T.derivative(coord, g, bc, d), g.size))
            out.append((f"n{n}-jump{coord}", T.jump(coord, g, T.DIR_NEU), g.size))
    g3 = T.Grid([0, 0, 0], [1, 1, 1], [3, 3, 1], [5, 4, 6], [T.DIR, T.PER, T.NEU])
    for coord in range(3):
        out.append((f"3d-d{coord}", T.derivative(coord, g3, g3.bc[coord], T.CENTERED), g3.size))
    g = T.Grid([0, 0], [1, 1], 3, [8, 12], [T.DIR, T.PER])
    for coord in (0, 1):
        out.append((f"proj{coord}", T.fast_projection(coord, g, 1, 2), None))
        out.append((f"proj4-{coord}", T.fast_projection(coord, g, 1, 4), None))
        out.append((f"interp{coord}", T.fast_interpolation(coord, g, 1, 2), None))
        out.append((f"projn{coord}", T.fast_projection(coord, g, 3, 1), None))
    return out


@pytest.mark.parametrize("case", _ell_cases(), ids=lambda c: c[0])
def test_ell_symv_vs_oracle(G, case):
    name, m, _ = case
    r = rng(11)
    x = r.uniform(-1, 1, m.total_cols)
    y0 = r.uniform(-1, 1, m.total_rows)
    dx = G.make(x)
    for al, be in ((1., 0.), (-1., 1.), (0.5, -2.)):
        y = y0.copy()
        if be == 0.:
            y[::7] = np.nan  # must be overwritten
        orc.ell_symv(m, al, x, be, y)
        for generic in (False, True):
            dy = G.make(y0 if be != 0. else np.where(np.arange(y0.size) % 7 == 0, np.nan, y0))
            G.symv(m, al, dx, be, dy, generic=generic)
            assert same_bits(y, G.get(dy)), (name, al, be, generic)


def test_ell_symv_fixtures(G, golden):
    from feltor_b200 import topology as T
    g2 = T.Grid([0, 0.1], [np.pi, 2 * np.pi + 0.1], 3, [8, 6], [T.DIR, T.PER])
    x2, y2 = golden["ell/x2"], golden["ell/y2"]
    dx = G.make(x2)
    for coord, bc in ((0, T.DIR), (1, T.PER), (0, T.NEU_DIR), (1, T.NEU)):
        for d in range(3):
            m = T.derivative(coord, g2, bc, d)
            for al, be in ((1., 0.), (-1., 1.), (0.5, -2.)):
                dy = G.make(y2)
                G.symv(m, al, dx, be, dy)
                assert same_bits(G.get(dy), golden[f"ell/d{coord}/bc{bc}/dir{d}/a{al}b{be}"])


def test_ell_right_range_and_errors(G):
    """set_right_range restricts the update to columns [a,b) (sparseblockmat.h:152-166); size mismatch raises"""
    from feltor_b200 import topology as T
    g = T.Grid([0, 0], [1, 1], 3, [6, 5], [T.DIR, T.PER])
    m = T.derivative(1, g, T.PER, T.FORWARD)
    r = rng(2)
    x, y0 = r.uniform(-1, 1, g.size), r.uniform(-1, 1, g.size)
    m.set_right_range(4, 11)
    y = y0.copy()
    orc.ell_symv(m, 1.5, x, 0., y)
    dy = G.make(y0)
    G.symv(m, 1.5, G.make(x), 0., dy)
    assert same_bits(y, G.get(dy))
    with pytest.raises(ValueError):
        G.symv(m, 1., G.make(x[:-1]), 0., dy)
    import feltor_b200 as fb
    with pytest.raises(fb.DgbError):
        G.symv(m, 1., dy, 0., dy)  # x aliases y


def test_ell_full_size_properties(G):
    """benchmark-size (n=3, 1024^2) size-independent checks: generic == fast kernel bit for bit; dx of a constant
    vanishes in the interior; linearity symv(x1+x2) ~ symv(x1)+symv(x2)"""
    import torch
    from feltor_b200 import topology as T
    g = T.Grid([0, 0], [1, 1], 3, [1024, 1024], [T.DIR, T.PER])
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(g.size, dtype=torch.float64, device="cuda", generator=gen)
    for coord, bc, d in ((0, T.DIR, T.FORWARD), (1, T.PER, T.CENTERED), (0, T.DIR, T.CENTERED), (1, T.PER, T.BACKWARD)):
        m = T.derivative(coord, g, bc, d)
        y1 = torch.full_like(x, float("nan"))
        y2 = torch.full_like(x, float("nan"))
        G.symv(m, 1., x, 0., y1)
        G.symv(m, 1., x, 0., y2, generic=True)
        assert torch.equal(y1.view(torch.int64), y2.view(torch.int64))
    m = T.derivative(1, g, T.PER, T.CENTERED)
    c = torch.ones_like(x)
    y = torch.empty_like(x)
    G.symv(m, 1., c, 0., y)
    assert float(y.abs().max()) < 1e-9


@pytest.mark.parametrize("n", [1, 31, 1000, 300001])
def test_reduce_closed_set(G, n):
    """blas1::reduce for max / min / logical-or (order independent: exact) and sum (to rounding); blas1_t.cpp reduce snippets"""
    from feltor_b200 import blas1
    r = rng(n)
    x = r.uniform(-3, 3, n)
    dx = G.make(x)
    assert blas1.reduce(dx, -1e300, "max") == x.max()
    assert blas1.reduce(dx, 1e10, "min") == min(1e10, x.min())
    assert blas1.reduce(dx, 0., "max", "abs") == np.abs(x).max()
    assert blas1.reduce(dx, 0., "or", "isnan") == 0.
    assert abs(blas1.reduce(dx, 0., "sum", "square") - np.sum(x * x)) <= 1e-12 * np.sum(x * x)
    x[n // 2] = np.nan
    assert blas1.reduce(G.make(x), 0., "or", "isnan") == 1.
    x[n // 2] = np.inf
    assert blas1.reduce(G.make(x), 0., "or", "isnan") == 0. and blas1.reduce(G.make(x), 0., "or", "isnotfinite") == 1.


def test_vdot_is_the_exact_dot_with_a_scalar_operand(G):
    """blas1::vdot(identity, x) / dot(1., x): the superaccumulator kernels accept scalar operands"""
    from feltor_b200 import blas2
    x = wide(rng(5), 5000, -40, 40)
    got = blas2.dot(1., G.make(x))
    import math
    assert got == math.fsum(x)


# ------------------------------------------------------------------------------------------------ CooSparseBlockMat
class _Coo(__import__("ctypes").Structure):
    import ctypes as _C
    _fields_ = [("num_rows", _C.c_int), ("num_cols", _C.c_int), ("num_entries", _C.c_int), ("n", _C.c_int),
                ("left_size", _C.c_int), ("right_size", _C.c_int), ("data", _C.c_void_p), ("rows_idx", _C.c_void_p),
                ("cols_idx", _C.c_void_p), ("data_idx", _C.c_void_p)]


@pytest.mark.parametrize("n,left,right,rows,chunks,entries", [(3, 1, 1, 7, 2, 5), (3, 4, 24, 12, 3, 9), (2, 5, 1, 9, 1, 4),
                                                              (4, 1, 30, 6, 4, 11), (5, 3, 7, 5, 2, 6), (3, 64, 9, 20, 2, 0)])
def test_coo_symv_vs_oracle(G, n, left, right, rows, chunks, entries):
    """dgb_coo_symv == CooSparseBlockMat::symv (sparseblockmat_omp_kernels.h:354-380): the outer (communicating) part of
    MPISparseBlockMat -- entries hit the same row repeatedly, x is a table of chunk pointers laid out [q][s][j]; bit-exact"""
    import ctypes as C
    import torch
    import feltor_b200 as fb
    from feltor_b200._dev import ptr, stream
    r = rng(n * 100 + left + entries)
    nblocks = 3
    data = r.uniform(-1, 1, nblocks * n * n)
    rows_idx = r.integers(0, rows, entries).astype(np.int32)
    cols_idx = r.integers(0, chunks, entries).astype(np.int32)
    data_idx = r.integers(0, nblocks, entries).astype(np.int32)
    xs = [r.uniform(-1, 1, n * left * right) for _ in range(chunks)]
    y0 = r.uniform(-1, 1, left * rows * n * right)
    yo = y0.copy()
    orc.coo_symv([rows, chunks, entries, n, left, right], data, rows_idx, cols_idx, data_idx, -0.7, xs, yo)
    d_data, d_y = G.make(data), G.make(y0)
    d_rows, d_cols, d_didx = (torch.from_numpy(a).cuda() for a in (rows_idx, cols_idx, data_idx))
    d_xs = [G.make(x) for x in xs]
    table = torch.tensor([t.data_ptr() for t in d_xs], dtype=torch.int64, device="cuda")
    m = _Coo(rows, chunks, entries, n, left, right, d_data.data_ptr(), d_rows.data_ptr(), d_cols.data_ptr(), d_didx.data_ptr())
    fb.lib().coo_symv(C.byref(m), C.c_double(-0.7), ptr(table), C.c_double(1.), ptr(d_y), stream())
    assert same_bits(G.get(d_y), yo)
    if entries:
        with pytest.raises(fb.DgbError):     # beta != 1 is rejected like the reference's assert (sparseblockmat.h:324)
            fb.lib().coo_symv(C.byref(m), C.c_double(1.), ptr(table), C.c_double(0.), ptr(d_y), stream())
