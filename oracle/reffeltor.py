"""TEST INFRASTRUCTURE: ctypes access to oracle/_ref/libdgref_feltor.so (the UNMODIFIED feltor::Explicit of src/feltor/feltor.h on the
reference's OpenMP backend, oracle/ref_feltor.cpp); tests clone this module with _PATH pointing at the build on the libdgb200
binding (integration/_build/libdgshim_feltor.so)."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, "oracle", "_ref", "libdgref_feltor.so")
_LIB = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(_PATH)
        _LIB.ref_feltor_create.restype = C.c_void_p
        _LIB.ref_feltor_create.argtypes = [C.c_char_p]
        for name in ("ref_feltor_free", "ref_feltor_size", "ref_feltor_state", "ref_feltor_rhs", "ref_feltor_euler", "ref_feltor_potential"):
            getattr(_LIB, name).argtypes = None
    return _LIB


def default_input():
    return open(os.path.join(ROOT, "tests", "golden", "feltor_input.json")).read()


class RefFeltor:
    def __init__(self, json_text=None):
        h = lib().ref_feltor_create((json_text or default_input()).encode())
        if not h:
            raise RuntimeError("ref_feltor_create failed")
        self.h = C.c_void_p(h)
        self.size = lib().ref_feltor_size(self.h)

    def _p(self, a):
        return a.ctypes.data_as(C.c_void_p)

    def state(self, field, species):
        out = np.empty(self.size)
        lib().ref_feltor_state(self.h, field, species, self._p(out))
        return out

    def rhs(self, t=0.):
        out = [np.empty(self.size) for _ in range(4)]
        if lib().ref_feltor_rhs(self.h, C.c_double(t), *[self._p(a) for a in out]) != 0:
            raise RuntimeError("ref_feltor_rhs failed")
        return out

    def euler(self, dt):
        lib().ref_feltor_euler(self.h, C.c_double(dt))

    def potential(self, i):
        out = np.empty(self.size)
        lib().ref_feltor_potential(self.h, i, self._p(out))
        return out

    def __del__(self):
        try:
            lib().ref_feltor_free(self.h)
        except Exception:
            pass
