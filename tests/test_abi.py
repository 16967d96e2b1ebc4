"""CPU: the C-ABI library loads without a GPU and exports every symbol include/dgb200.h declares; host-only entry
points (topology, superaccumulator helpers, error reporting) work; compute entry points fail loudly without a device."""
import ctypes as C
import numpy as np
import pytest
import feltor_b200 as fb
from feltor_b200._lib import parse_header
from oracle import orc
from util import rng, wide


def test_exports_every_declared_symbol():
    L = fb.lib()
    protos = parse_header()
    assert len(protos) > 70
    assert L.missing == []
    for name in protos:
        assert hasattr(L.cdll, name), name


def test_version_and_error_string():
    L = fb.lib()
    assert L.raw["dgb_version"]() >= 100
    with pytest.raises(fb.DgbError) as e:
        L.topo_dlt(0, 99, None)
    assert "dgb_topo_dlt" in str(e.value)


def test_host_superacc_helpers_match_oracle():
    """dgb_superacc_normalize_host / round_host == accumulate.h:267-349 as restated (and pinned) in the oracle"""
    L = fb.lib()
    r = rng(1)
    for n in (1, 10, 1000):
        x, y = wide(r, n, -200, 200), wide(r, n, -50, 50)
        acc, _ = orc.exdot2(x, y)
        raw = acc.copy()
        raw[3] += 5 << 56  # de-normalise: push carries into a word
        raw[4] -= 5
        a = raw.copy()
        L.raw["dgb_superacc_normalize_host"](a.ctypes.data, None)
        assert np.array_equal(a, acc)
        assert L.raw["dgb_superacc_round_host"](raw.ctypes.data) == orc.round_acc(acc)


def test_host_superacc_negative_sum_is_not_an_error():
    """a negative accumulator (dot(x, -x)) must come back as status 0 with the sign in the out-parameter: the helper used
    to return exblas::cpu::Normalize's sign flag as if it were an error code"""
    L = fb.lib()
    x = rng(2).uniform(-1, 1, 257)
    acc, _ = orc.exdot2(x, -x)
    raw = acc.copy()
    raw[20] -= 3 << 56
    raw[21] += 3
    neg = C.c_int(-1)
    assert L.superacc_normalize_host(raw.ctypes.data_as(C.c_void_p), C.byref(neg)) == 0   # the CHECKED wrapper: no DgbError
    assert neg.value == 1 and np.array_equal(raw, orc.normalize(acc))
    assert L.raw["dgb_superacc_round_host"](raw.ctypes.data) == orc.round_acc(acc) < 0
    acc, _ = orc.exdot2(x, x)
    assert L.superacc_normalize_host(acc.ctypes.data_as(C.c_void_p), C.byref(neg)) == 0 and neg.value == 0


def test_compute_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    ws = C.c_void_p()
    with pytest.raises(fb.DgbError):
        fb.lib().dot_ws_create(C.byref(ws))
    p = C.c_void_p()
    with pytest.raises(fb.DgbError):
        fb.lib().malloc(C.byref(p), 1024)
    # kernels launched on host memory without a device: every entry point reports the CUDA error instead of computing
    a, b = np.ones(64), np.ones(64)
    pa, pb = a.ctypes.data, b.ctypes.data
    d = C.c_double
    pos, idx = np.arange(65, dtype=np.int32), np.arange(64, dtype=np.int32)
    calls = [
        lambda: fb.lib().axpby(64, d(1.), pa, d(1.), pb, None),
        lambda: fb.lib().adaptive_tolerance(64, d(1.), d(1.), pa, pb, None),
        lambda: fb.lib().csr_stencil(0, 64, pos.ctypes.data, idx.ctypes.data, None, d(0.), pa, pb, None),
        lambda: fb.lib().ds_apply_vol(10, 64, d(1.), pa, pa, None, None, None, None, None, None, None, d(1.), d(0.), pb, None),
        lambda: fb.lib().reduce(64, pa, 0, 0, d(0.), pb, None),
    ]
    for k, call in enumerate(calls):
        with pytest.raises(fb.DgbError):
            call()
        assert np.all(b == 1.), k   # nothing was computed on the host


def test_headers_compile_standalone(tmp_path):
    """include/dgb200.h is a plain C99 header (the drop-in boundary has no C++ or torch types in it) and
    include/dg_b200.hpp compiles on its own as pedantic C++17"""
    import os
    import subprocess
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    c = tmp_path / "h.c"
    c.write_text('#include "dgb200.h"\nint main(void) { return sizeof(dgb_dot_result) == 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I" + inc, str(c)])
    cpp = tmp_path / "h.cpp"
    cpp.write_text('#include "dg_b200.hpp"\n#include "dg_b200.hpp"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only", "-I" + inc, str(cpp)])
