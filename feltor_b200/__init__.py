"""feltor_b200: B200-native (sm_100a) data-parallel core of the Feltor `dg` library.

The product is the C-ABI library feltor_b200/libdgb200.so (include/dgb200.h) plus the C++ host layer in
include/dg_b200/.  This Python package is the thin test/bench harness around the C ABI: device memory and
streams come from torch (plumbing), every computation is a call into libdgb200.so.  There is no CPU path.
"""
from ._lib import lib, DgbError, LIB_PATH  # noqa: F401
from . import blas1, blas2  # noqa: F401
from .blas2 import Ell, DotWorkspace  # noqa: F401
