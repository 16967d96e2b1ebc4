"""CPU: pins the oracle's parallel-derivative formulas (orc_ds_apply) and TensorMultiply3d against fixtures produced by
the UNMODIFIED reference functions (tests/golden/make_golden_ds.py -> ds_golden.npz).  TensorMultiply3d is written with
explicit DG_FMA in the reference: bit-exact.  The ds formulas are user lambdas that the reference's compiler is free to
contract (g++ -O2 -mfma does): tolerance 1e-13 relative to the operand scale, stated here."""
import os
import numpy as np
import pytest
from oracle import orc
from util import same_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = {0: ("f", "fp", None), 1: ("f", "fm", None), 2: ("fm", "fp", None), 3: ("f", "fp", "fpp"), 4: ("f", "fm", "fmm"),
        5: ("fm", "f", "fp"), 6: ("fm", "f", "fp"), 7: ("fm", "f", None), 8: ("f", "fp", None), 9: ("fm", "fp", None),
        10: ("fm", "fp", None)}
DS_TOL = 1e-13


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ds_golden.npz"))


def ds_close(got, want):
    scale = np.maximum(np.abs(want), 1.0)
    return np.max(np.abs(got - want) / scale) < DS_TOL


@pytest.mark.parametrize("kind", range(11))
def test_ds_formulas_fixture(gold, kind):
    a, b, c = ARGS[kind]
    G = tuple(gold["ds/" + k] for k in ("Gm", "G0", "Gp"))
    B = tuple(gold["ds/" + k] for k in ("bm", "b0", "bp"))
    for beta in (-0.3, 0.0):
        g = gold["ds/g0"].copy() if beta != 0. else np.full_like(gold["ds/g0"], np.nan)
        orc.ds_apply(kind, 0.7, gold["ds/" + a], gold["ds/" + b], gold["ds/" + c] if c else None, G, B,
                     float(gold["ds/delta"][0]), beta, g)
        assert ds_close(g, gold[f"ds/kind{kind}/beta{int(beta != 0)}"]), (kind, beta)


def test_tensor_multiply3d_fixture(gold):
    t, ins = list(gold["t3d/t"]), list(gold["t3d/in"])
    o = [a.copy() for a in gold["t3d/out0"]]
    orc.tensor_multiply3d(gold["t3d/lambda"], t, ins, 0.3, o)
    assert same_bits(np.stack(o), gold["t3d/out"])
    o = [a.copy() for a in ins]
    orc.tensor_multiply3d(gold["t3d/lambda"], t, o, 0., o)
    assert same_bits(np.stack(o), gold["t3d/out_alias"])
    # identity tensor, scalar lambda: out = 2 in (temp = out*mu as in the reference: no NaN overwrite for mu = 0)
    o = [np.zeros_like(a) for a in ins]
    orc.tensor_multiply3d(2., None, ins, 0., o)
    assert same_bits(np.stack(o), 2. * np.stack(ins))


@pytest.mark.parametrize("kind", range(4))
def test_csr_stencil_fixture(gold, kind):
    """blas2::stencil with CSRMedianFilter / CSRSWMFilter / CSRAverageFilter / CSRSymvFilter (filter.h:174-266): bit-exact"""
    y = np.full(gold["stencil/x"].size, np.nan)
    orc.csr_stencil(kind, gold["stencil/pos"], gold["stencil/idx"], gold["stencil/val"], 1.5, gold["stencil/x"], y)
    assert same_bits(y, gold[f"stencil/kind{kind}"])


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("bound", [4, 1])
def test_assign_bc_along_field_fixture(gold, order, bound):
    """assign_bc_along_field_1st / _2nd (ds.h:169-296), NEU (4) and DIR (1): user lambdas, tolerance 1e-13 as above"""
    k = {n: gold["bc/" + n] for n in ("fm", "f", "fp", "hbm", "hbp", "bbm", "bbo", "bbp")}
    fmg, fpg = np.full(k["fm"].size, np.nan), np.full(k["fm"].size, np.nan)
    orc.assign_bc_along_field(order, bound == 4, 2 * np.pi / 7, k["fm"], k["f"] if order == 2 else None, k["fp"], k["hbm"], k["hbp"],
                              k["bbm"], k["bbo"], k["bbp"], (0.3, -0.2), fmg, fpg)
    assert ds_close(fmg, gold[f"bc/order{order}/bound{bound}/fmg"]) and ds_close(fpg, gold[f"bc/order{order}/bound{bound}/fpg"])
    # interior points (no mask set) keep the shifted values exactly
    inner = (k["bbm"] + k["bbo"] + k["bbp"]) == 0
    assert np.array_equal(fmg[inner], k["fm"][inner]) and np.array_equal(fpg[inner], k["fp"][inner])
