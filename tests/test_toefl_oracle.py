"""CPU: pins the wrapped reference toefl (oracle/_ref/libdgref_toefl.so) to the committed golden vectors, so that a GPU
failure against the fixture cannot be a stale fixture.  Skipped when the reference tree / wrapper is not available."""
import os
import numpy as np
import pytest
from util import same_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model", ["global", "local"])
def test_reference_reproduces_golden(model):
    from oracle import reftoefl as R
    if not R.available():
        pytest.skip("oracle/_ref/libdgref_toefl.so not built (needs /root/reference)")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "toefl_golden.npz"))
    ref = R.RefToefl(R.default_params(3, 24, 24, model__type=model))
    y0, y1 = ref.init()
    assert same_bits(y0, gold[model + "_init0"]) and same_bits(y1, gold[model + "_init1"])
    a, b, _ = ref.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 3, y0, y1)
    assert same_bits(a, gold[model + "_y0"]) and same_bits(b, gold[model + "_y1"])
    assert same_bits(ref.phi(0), gold[model + "_phi0"])


def test_reference_adaptive_reproduces_golden():
    """dg::Adaptive<ERKStep> + pid_control + l2norm (adaptive.h:232-395), incl. a rejected step"""
    from oracle import reftoefl as R
    if not R.available():
        pytest.skip("oracle/_ref/libdgref_toefl.so not built (needs /root/reference)")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "toefl_golden.npz"))
    ref = R.RefToefl(R.default_params(3, 24, 24, model__type="global"))
    y0, y1 = ref.init()
    a, b, t, dts, nf = ref.adaptive("Bogacki-Shampine-4-2-3", 0., 60., 4, 1e-5, 1e-6, y0, y1)
    assert same_bits(a, gold["adaptB_y0"]) and same_bits(b, gold["adaptB_y1"]) and same_bits(dts, gold["adaptB_dts"])
    assert t == gold["adaptB_t_nfailed"][0] and nf == int(gold["adaptB_t_nfailed"][1]) == 1


def test_reference_multistep_reproduces_golden():
    """dg::ExplicitMultistep("TVB-3-3") (multistep.h:59-100)"""
    from oracle import reftoefl as R
    if not R.available():
        pytest.skip("oracle/_ref/libdgref_toefl.so not built (needs /root/reference)")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "toefl_golden.npz"))
    ref = R.RefToefl(R.default_params(3, 24, 24, model__type="global"))
    y0, y1 = ref.init()
    a, b, ts = ref.multistep("TVB-3-3", 0., 0.3, 5, y0, y1)
    assert same_bits(a, gold["msTVB_y0"]) and same_bits(b, gold["msTVB_y1"]) and same_bits(ts, gold["msTVB_ts"])


def test_toefl_host_time_arithmetic():
    """the stage times of ShuOsher::step are host expressions the reference's compiler contracts into FMAs; the harness
    reproduces them with exact rational arithmetic (no GPU needed to check the helper)"""
    from feltor_b200.toefl import _host_fma
    import math
    assert _host_fma(2. / 3., 0.75, 0.) == math.fma(2. / 3., 0.75, 0.) if hasattr(math, "fma") else True
    assert _host_fma(0.1, 3., -0.3) != 0.1 * 3. - 0.3   # one rounding, not two
    assert _host_fma(1., 0.5, 0.25) == 0.75


def test_toefl_harness_has_no_oracle_import():
    src = open(os.path.join(ROOT, "feltor_b200", "toefl.py")).read()
    assert "oracle" not in src.replace("oracle/", "")
