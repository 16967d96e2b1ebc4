"""GPU parity of the Fieldaligned shift (ePlus/eMinus incl. ghost-cell boundary fix-up) and the DS formulas against a
numpy/oracle restatement of inc/geometries/fieldaligned.h:850-912 and ds.h:744-852 (the CSR part bitwise, the
user-lambda formulas to 1e-14 relative), plus fused DS::centered == its composition bit for bit."""
import ctypes as C
import os
import numpy as np
import pytest
from oracle import orc
from util import same_bits, rng

pytestmark = pytest.mark.gpu
PER, DIR, DIR_NEU, NEU_DIR, NEU = 0, 1, 2, 3, 4


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


def stencil_csr(r, n, width, maxlen):
    """interpolation-like matrix: each row couples to a few columns near its own index"""
    counts = r.integers(1, maxlen + 1, n)
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    idx = np.concatenate([np.clip(i + r.integers(-width, width + 1, c), 0, n - 1) for i, c in enumerate(counts)]).astype(np.int32)
    val = r.uniform(-1, 1, pos[-1])
    return pos, idx, val


def oracle_shift(plus, pos, idx, val, f, nplanes, bcz, bnd, lim, dphi):
    n = len(pos) - 1
    out = np.zeros_like(f)
    for k in range(nplanes):
        src = (k + 1) % nplanes if plus else (k - 1) % nplanes
        y = out[k * n:(k + 1) * n]
        orc.csr_spmv(pos, idx, val, 1., np.ascontiguousarray(f[src * n:(src + 1) * n]), 0., y)
    if bcz != PER:
        i0 = nplanes - 1 if plus else 0
        fi, ti = f[i0 * n:(i0 + 1) * n], out[i0 * n:(i0 + 1) * n]
        ghost = np.empty(n)
        dirichlet = bcz in ((DIR, NEU_DIR) if plus else (DIR, DIR_NEU))
        if dirichlet:
            orc.axpbyz(2., bnd, -1., np.ascontiguousarray(fi), ghost)
        else:
            orc.axpbyz(dphi if plus else -dphi, bnd, 1., np.ascontiguousarray(fi), ghost)
        orc.axpbyz(1., ghost.copy(), -1., np.ascontiguousarray(ti), ghost)
        orc.pointwiseDot(1., lim, ghost, 1., ti)
    return out


def dev_csr(G, pos, idx, val):
    from feltor_b200._dev import dvec
    return dvec(pos), dvec(idx), dvec(val)


@pytest.mark.parametrize("bcz", [PER, DIR, NEU, DIR_NEU, NEU_DIR])
def test_fieldaligned_shift(G, bcz):
    from feltor_b200 import lib
    from feltor_b200._dev import ptr, stream
    r = rng(bcz)
    n, nz = 700, 7
    pos, idx, val = stencil_csr(r, n, 30, 35)
    f = r.uniform(-1, 1, n * nz)
    bnd, lim = r.uniform(-1, 1, n), r.integers(0, 2, n).astype(np.float64)
    dphi = 2 * np.pi / nz
    dpos, didx, dval = dev_csr(G, pos, idx, val)
    df, dbnd, dlim = G.make(f), G.make(bnd), G.make(lim)  # keep the device operands alive across the call
    for plus in (1, 0):
        want = oracle_shift(plus, pos, idx, val, f, nz, bcz, bnd, lim, dphi)
        out = G.make(np.full(n * nz, np.nan))
        ghost = G.make(np.zeros(n))
        lib().fa_shift(plus, n, nz, ptr(dpos), ptr(didx), ptr(dval), ptr(df), ptr(out), bcz, ptr(dbnd), ptr(dlim),
                       ptr(ghost), C.c_double(dphi), stream())
        assert same_bits(G.get(out), want), (bcz, plus)


def test_ds_formulas_and_fused_centered(G):
    from feltor_b200 import lib
    from feltor_b200._dev import ptr, stream
    r = rng(9)
    n, nz = 1000, 9
    N = n * nz
    ppos, pidx, pval = stencil_csr(r, n, 20, 40)
    mpos, midx, mval = stencil_csr(r, n, 20, 40)
    f, g0 = r.uniform(-1, 1, N), r.uniform(-1, 1, N)
    bphi, bm, bp = r.uniform(0.5, 1.5, N), r.uniform(0.5, 1.5, N), r.uniform(0.5, 1.5, N)
    dphi = 2 * np.pi / nz
    fp = oracle_shift(1, ppos, pidx, pval, f, nz, PER, None, None, dphi)
    fm = oracle_shift(0, mpos, midx, mval, f, nz, PER, None, None, dphi)
    al, be = 0.7, -0.3
    want = {0: al * bphi * (fp - f) / dphi + be * g0, 1: al * bphi * (f - fm) / dphi + be * g0,
            2: al * bphi * (fp - fm) / 2. / dphi + be * g0,
            3: al * bphi * (-3. * f + 4. * fp - fm) / 2. / dphi + be * g0,   # (fm stands in for fpp)
            4: al * bphi * (3. * f - 4. * fm + fp) / 2. / dphi + be * g0,
            5: al * bphi * (((bp + bphi) / 2.) * ((fp - f) / dphi) - ((bm + bphi) / 2.) * ((f - fm) / dphi)) / dphi + be * g0}
    args = {0: (f, fp, None), 1: (f, fm, None), 2: (fm, fp, None), 3: (f, fp, fm), 4: (f, fm, fp), 5: (fm, f, fp)}
    d = {k: G.make(v) for k, v in dict(f=f, fp=fp, fm=fm, bphi=bphi, bm=bm, bp=bp).items()}
    name = {id(f): "f", id(fp): "fp", id(fm): "fm"}
    for kind in range(6):
        a, b, c = args[kind]
        g = G.make(g0)
        lib().ds_apply(kind, N, C.c_double(al), ptr(d[name[id(a)]]), ptr(d[name[id(b)]]), ptr(d[name[id(c)]]) if c is not None else None,
                       ptr(d["bm"]), ptr(d["bphi"]), ptr(d["bp"]), C.c_double(dphi), C.c_double(be), ptr(g), stream())
        got = G.get(g)
        assert np.max(np.abs(got - want[kind]) / (np.abs(want[kind]) + 1e-300)) < 1e-13 or np.allclose(got, want[kind], rtol=1e-14, atol=1e-15), kind
    # fused DS::centered == shift + shift + ds_centered, bit for bit; beta = 0 overwrites NaN
    dp, dm = dev_csr(G, ppos, pidx, pval), dev_csr(G, mpos, midx, mval)
    for beta in (be, 0.):
        g1 = G.make(g0 if beta != 0. else np.full(N, np.nan))
        lib().ds_centered_fused(n, nz, ptr(dp[0]), ptr(dp[1]), ptr(dp[2]), ptr(dm[0]), ptr(dm[1]), ptr(dm[2]), C.c_double(al),
                                ptr(d["f"]), ptr(d["bphi"]), C.c_double(dphi), C.c_double(beta), ptr(g1), stream())
        g2 = G.make(g0 if beta != 0. else np.full(N, np.nan))
        lib().ds_apply(2, N, C.c_double(al), ptr(d["fm"]), ptr(d["fp"]), None, None, ptr(d["bphi"]), None, C.c_double(dphi),
                       C.c_double(beta), ptr(g2), stream())
        assert same_bits(G.get(g1), G.get(g2)), beta


ARGS = {0: ("f", "fp", None), 1: ("f", "fm", None), 2: ("fm", "fp", None), 3: ("f", "fp", "fpp"), 4: ("f", "fm", "fmm"),
        5: ("fm", "f", "fp"), 6: ("fm", "f", "fp"), 7: ("fm", "f", None), 8: ("f", "fp", None), 9: ("fm", "fp", None),
        10: ("fm", "fp", None)}


@pytest.mark.parametrize("kind", range(11))
def test_ds_formulas_golden_and_oracle(G, kind):
    """all eleven ds.h:743-1000 formulas: bit-identical to the oracle restatement (same rounding sequence), and within
    1e-13 of the fixture computed by the unmodified reference functions (their lambdas are contracted by its compiler);
    beta = 0 overwrites NaN"""
    import os
    from feltor_b200 import lib
    from feltor_b200._dev import ptr, stream
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ds_golden.npz"))
    a, b, c = ARGS[kind]
    dv = {k: G.make(gold["ds/" + k]) for k in ("f", "fm", "fp", "fmm", "fpp", "Gm", "G0", "Gp", "bm", "b0", "bp")}
    n, delta = gold["ds/f"].size, float(gold["ds/delta"][0])
    for beta in (-0.3, 0.0):
        g0 = gold["ds/g0"].copy() if beta != 0. else np.full(n, np.nan)
        g = G.make(g0)
        cc = ptr(dv[c]) if c else None
        if kind < 6:
            lib().ds_apply(kind, n, C.c_double(0.7), ptr(dv[a]), ptr(dv[b]), cc, ptr(dv["bm"]), ptr(dv["b0"]), ptr(dv["bp"]),
                           C.c_double(delta), C.c_double(beta), ptr(g), stream())
        else:
            lib().ds_apply_vol(kind, n, C.c_double(0.7), ptr(dv[a]), ptr(dv[b]), cc, ptr(dv["Gm"]), ptr(dv["G0"]), ptr(dv["Gp"]),
                               ptr(dv["bm"]), ptr(dv["b0"]), ptr(dv["bp"]), C.c_double(delta), C.c_double(beta), ptr(g), stream())
        got = G.get(g)
        want = g0.copy()
        orc.ds_apply(kind, 0.7, gold["ds/" + a], gold["ds/" + b], gold["ds/" + c] if c else None,
                     tuple(gold["ds/" + k] for k in ("Gm", "G0", "Gp")), tuple(gold["ds/" + k] for k in ("bm", "b0", "bp")),
                     delta, beta, want)
        assert same_bits(got, want), (kind, beta)
        ref = gold[f"ds/kind{kind}/beta{int(beta != 0)}"]
        assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.)) < 1e-13, (kind, beta)


def test_ds_apply_vol_rejects_missing_operands(G):
    from feltor_b200 import lib, DgbError
    from feltor_b200._dev import ptr, stream
    x = G.make(np.ones(8))
    with pytest.raises(DgbError):
        lib().ds_apply_vol(7, 8, C.c_double(1.), ptr(x), ptr(x), None, None, ptr(x), None, ptr(x), ptr(x), None,
                           C.c_double(1.), C.c_double(0.), ptr(x), stream())
    with pytest.raises(DgbError):
        lib().ds_apply_vol(11, 8, C.c_double(1.), ptr(x), ptr(x), None, None, None, None, None, None, None,
                           C.c_double(1.), C.c_double(0.), ptr(x), stream())


@pytest.mark.parametrize("n,nz,maxlen", [(700, 7, 35), (1000, 64, 90), (33, 1, 5), (4096, 9, 40)])
def test_gather_plan_equals_csr_kernels(G, n, nz, maxlen):
    """the sliced-ELL gather plan (dgb_gather_*) == the CSR kernels bit for bit: all-planes SpMV for every beta / shift,
    fused DS::centered; ragged rows including empty ones"""
    from feltor_b200 import lib
    from feltor_b200._dev import ptr, stream
    r = rng(n + nz)
    pos, idx, val = stencil_csr(r, n, 30, maxlen)
    counts = np.diff(pos)
    # make some rows empty
    keep = r.uniform(0, 1, n) > 0.05
    sel = np.repeat(keep, counts)
    idx, val = idx[sel], val[sel]
    pos = np.concatenate([[0], np.cumsum(counts * keep)]).astype(np.int32)
    mpos, midx, mval = stencil_csr(r, n, 25, maxlen)
    dP, dM = dev_csr(G, pos, idx, val), dev_csr(G, mpos, midx, mval)
    hp, hm = C.c_void_p(), C.c_void_p()
    lib().gather_plan_create(C.byref(hp), n, n, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), stream())
    lib().gather_plan_create(C.byref(hm), n, n, ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), stream())
    f, y0 = r.uniform(-1, 1, n * nz), r.uniform(-1, 1, n * nz)
    df = G.make(f)
    for alpha, beta, shift in ((1., 0., 1), (0.7, 1., -1), (-1.3, 0.4, 0)):
        a = G.make(y0 if beta != 0. else np.full(n * nz, np.nan))
        b = G.make(y0 if beta != 0. else np.full(n * nz, np.nan))
        lib().csr_spmv_planes(n, n, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), C.c_double(alpha), ptr(df), C.c_double(beta), ptr(a), nz, shift, stream())
        lib().gather_spmv_planes(hp, C.c_double(alpha), ptr(df), C.c_double(beta), ptr(b), nz, shift, stream())
        assert same_bits(G.get(a), G.get(b)), (alpha, beta, shift)
    bphi = G.make(r.uniform(0.5, 1.5, n * nz))
    for beta in (0., -0.3):
        a = G.make(y0 if beta != 0. else np.full(n * nz, np.nan))
        b = G.make(y0 if beta != 0. else np.full(n * nz, np.nan))
        lib().ds_centered_fused(n, nz, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), C.c_double(0.7),
                                ptr(df), ptr(bphi), C.c_double(0.1), C.c_double(beta), ptr(a), stream())
        lib().gather_ds_centered(hp, hm, nz, C.c_double(0.7), ptr(df), ptr(bphi), C.c_double(0.1), C.c_double(beta), ptr(b), stream())
        assert same_bits(G.get(a), G.get(b)), beta
    lib().gather_plan_destroy(hp)
    lib().gather_plan_destroy(hm)


# ------------------------------------------------------------------------------------------------ cell-tiled layout
def _fieldaligned_like_matrix(r, n, Nx, Ny, max_shift=3):
    """CSR matrix with the structure of Fieldaligned's I+ / I-: all n^2 rows of a target cell couple to all nodes of the same
    2..5 source cells (smoothly displaced, clipped at the boundary), columns ascending"""
    rowlen = Nx * n
    pos, idx, val = [0], [], []
    rows = {}
    for cy in range(Ny):
        for cx in range(Nx):
            sx = int(np.clip(cx + round(max_shift * np.sin(0.3 * cy + 0.1 * cx)), 0, Nx - 2))
            sy = int(np.clip(cy + round(max_shift * np.cos(0.2 * cx)), 0, Ny - 2))
            cells = {(sy, sx), (sy, sx + 1), (sy + 1, sx), (sy + 1, sx + 1)}
            k = r.integers(0, 3)
            if k == 1 and sx + 2 < Nx:
                cells.add((sy, sx + 2))
            if k == 2:
                cells = {(sy, sx), (sy, sx + 1)}
            cols = sorted((py * n + ky) * rowlen + px * n + kx for (py, px) in cells for ky in range(n) for kx in range(n))
            for ky in range(n):
                for kx in range(n):
                    rows[(cy * n + ky) * rowlen + cx * n + kx] = (cols, r.uniform(-1, 1, len(cols)))
    for i in range(n * n * Nx * Ny):
        c, v = rows[i]
        idx.extend(c)
        val.extend(v.tolist())
        pos.append(len(idx))
    return np.array(pos, dtype=np.int32), np.array(idx, dtype=np.int32), np.array(val)


@pytest.mark.parametrize("n,Nx,Ny,nz", [(3, 40, 12, 20), (2, 33, 9, 5), (3, 96, 7, 16), (4, 20, 6, 3)])
def test_celltile_plan_equals_csr_kernels(G, n, Nx, Ny, nz):
    """the cell-tiled layout (dgb_celltile_*) == the CSR kernels bit for bit: all-planes SpMV and the fused DS::centered; ragged
    strips (Nx not a multiple of 32), plane counts that are not multiples of the planes per CTA, n = 2, 3, 4"""
    import ctypes as C
    import torch
    from feltor_b200._lib import lib
    from feltor_b200._dev import ptr, stream
    r = rng(n * 10 + Nx)
    P, M = _fieldaligned_like_matrix(r, n, Nx, Ny), _fieldaligned_like_matrix(r, n, Nx, Ny, 2)
    rows = n * n * Nx * Ny
    f = r.uniform(-1, 1, rows * nz)
    df, bphi = G.make(f), G.make(r.uniform(0.5, 1.5, rows * nz))
    dP = [torch.from_numpy(a).cuda() for a in P]
    dM = [torch.from_numpy(a).cuda() for a in M]
    hp, hm = C.c_void_p(), C.c_void_p()
    lib().celltile_plan_create(C.byref(hp), n, Nx, Ny, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), stream())
    lib().celltile_plan_create(C.byref(hm), n, Nx, Ny, ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), stream())
    for alpha, beta, shift in ((1., 0., 1), (-1., 0., -1), (1., 0.5, 0), (1., 0., nz + 2)):
        y0 = r.uniform(-1, 1, rows * nz)
        a, b = G.make(y0), G.make(y0)
        lib().csr_spmv_planes(rows, rows, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), C.c_double(alpha), ptr(df), C.c_double(beta), ptr(a), nz, shift, stream())
        lib().celltile_spmv_planes(hp, C.c_double(alpha), ptr(df), C.c_double(beta), ptr(b), nz, shift, stream())
        assert same_bits(G.get(a), G.get(b)), (alpha, beta, shift)
    for beta in (0., 2.):
        y0 = r.uniform(-1, 1, rows * nz)
        a, b = G.make(y0), G.make(y0)
        lib().ds_centered_fused(rows, nz, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), C.c_double(0.7), ptr(df),
                                ptr(bphi), C.c_double(0.1), C.c_double(beta), ptr(a), stream())
        lib().celltile_ds_centered(hp, hm, nz, C.c_double(0.7), ptr(df), ptr(bphi), C.c_double(0.1), C.c_double(beta), ptr(b), stream())
        assert same_bits(G.get(a), G.get(b)), beta
    lib().celltile_plan_destroy(hp)
    lib().celltile_plan_destroy(hm)
    # a matrix without the cell structure is refused (the gather plan takes it)
    import feltor_b200 as fb
    bad = P[1].copy()
    bad[0] = (bad[0] + 1) % rows
    dbad = torch.from_numpy(bad).cuda()
    with pytest.raises(fb.DgbError):
        lib().celltile_plan_create(C.byref(hp), n, Nx, Ny, ptr(dP[0]), ptr(dbad), ptr(dP[2]), stream())


def test_celltile_ds_centered_on_reference_fieldaligned(G):
    """REAL field-line matrices: the unmodified dg::geo::Fieldaligned (circular field of inc/geometries/ds_b.cpp, built live in
    oracle/_ref/libdgref_fa.so) hands out I+ / I- and bphi; DS::centered on the cell-tiled plan == the gather plan bitwise and ==
    the reference's ds.centered to 1e-13 (the reference's compiler contracts the formula, tests/test_ds_oracle.py)"""
    import ctypes as C
    import torch
    from feltor_b200._lib import lib
    from feltor_b200._dev import ptr, stream
    from oracle import reffa
    if not reffa.available():
        pytest.skip("oracle/_ref/libdgref_fa.so not present")
    n, Nx, Ny, Nz = 3, 24, 20, 12
    F = reffa.RefFieldaligned(n, Nx, Ny, Nz, 6, 6, "dg")
    f = F.testfunction()
    gref, _ = F.ds("centered", 0.8, f, 0., np.zeros(F.size))
    P, M = F.csr("plus"), F.csr("minus")
    dP = [torch.from_numpy(a).cuda() for a in P]
    dM = [torch.from_numpy(a).cuda() for a in M]
    df, bphi = G.make(f), G.make(F.field("bphi"))
    hp, hm, gp, gm = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    lib().celltile_plan_create(C.byref(hp), n, Nx, Ny, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), stream())
    lib().celltile_plan_create(C.byref(hm), n, Nx, Ny, ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), stream())
    lib().gather_plan_create(C.byref(gp), F.plane, F.plane, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), stream())
    lib().gather_plan_create(C.byref(gm), F.plane, F.plane, ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), stream())
    a, b = G.make(np.zeros(F.size)), G.make(np.zeros(F.size))
    lib().celltile_ds_centered(hp, hm, Nz, C.c_double(0.8), ptr(df), ptr(bphi), C.c_double(F.delta_phi), C.c_double(0.), ptr(a), stream())
    lib().gather_ds_centered(gp, gm, Nz, C.c_double(0.8), ptr(df), ptr(bphi), C.c_double(F.delta_phi), C.c_double(0.), ptr(b), stream())
    assert same_bits(G.get(a), G.get(b))
    assert np.abs(G.get(a) - gref).max() <= 1e-13 * np.abs(gref).max()
    for h in (hp, hm):
        lib().celltile_plan_destroy(h)
    for h in (gp, gm):
        lib().gather_plan_destroy(h)


# ------------------------------------------------------------------------------------------------ boundary conditions along the field
@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("bound", [4, 1])
def test_assign_bc_along_field_vs_oracle_and_fixture(G, order, bound):
    """dgb_assign_bc_along_field == the oracle restatement bitwise (same left-to-right arithmetic) and == the fixtures of the
    unmodified reference functions (ds.h:169-296, user lambdas) to 1e-13; aliasing fmg = fm, fpg = fp as DS uses it"""
    import ctypes as C
    from feltor_b200._lib import lib
    from feltor_b200._dev import ptr, stream
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ds_golden.npz"))
    k = {n: gold["bc/" + n] for n in ("fm", "f", "fp", "hbm", "hbp", "bbm", "bbo", "bbp")}
    n = k["fm"].size
    delta = 2 * np.pi / 7
    fmo, fpo = np.full(n, np.nan), np.full(n, np.nan)
    orc.assign_bc_along_field(order, bound == 4, delta, k["fm"], k["f"] if order == 2 else None, k["fp"], k["hbm"], k["hbp"], k["bbm"],
                              k["bbo"], k["bbp"], (0.3, -0.2), fmo, fpo)
    d = {key: G.make(v) for key, v in k.items()}
    fmg, fpg = G.make(k["fm"]), G.make(k["fp"])   # in place
    lib().assign_bc_along_field(order, bound, n, C.c_double(delta), ptr(fmg), ptr(d["f"]), ptr(fpg), ptr(d["hbm"]), ptr(d["hbp"]),
                                ptr(d["bbm"]), ptr(d["bbo"]), ptr(d["bbp"]), C.c_double(0.3), C.c_double(-0.2), ptr(fmg), ptr(fpg), stream())
    assert same_bits(G.get(fmg), fmo) and same_bits(G.get(fpg), fpo)
    for got, name in ((G.get(fmg), "fmg"), (G.get(fpg), "fpg")):
        ref = gold[f"bc/order{order}/bound{bound}/{name}"]
        assert np.abs(got - ref).max() <= 1e-13 * max(np.abs(ref).max(), 1.)
    # swap_bc_perp: values outside the box change sign
    sm, sp = G.make(np.zeros(n)), G.make(np.zeros(n))
    lib().swap_bc_perp(n, ptr(d["fm"]), ptr(d["fp"]), ptr(d["bbm"]), ptr(d["bbo"]), ptr(d["bbp"]), ptr(sm), ptr(sp), stream())
    em = (1. - k["bbo"] - k["bbm"]) * k["fm"] + (k["bbm"] + k["bbo"]) * (-k["fm"])
    ep = (1. - k["bbo"] - k["bbp"]) * k["fp"] + (k["bbp"] + k["bbo"]) * (-k["fp"])
    assert same_bits(G.get(sm), em) and same_bits(G.get(sp), ep)
