"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libdgref_toefl.so (the unmodified reference toefl::Explicit)."""
import ctypes as C
import json
import os
import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libdgref_toefl.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.ref_toefl_create.restype = C.c_void_p
        _lib.ref_toefl_create.argtypes = [C.c_char_p]
        _lib.ref_toefl_free.argtypes = [C.c_void_p]
        _lib.ref_toefl_size.argtypes = [C.c_void_p]
        _lib.ref_toefl_init.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_toefl_rhs.restype = C.c_double
        _lib.ref_toefl_rhs.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 4
        _lib.ref_toefl_phi.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.ref_toefl_erk.restype = C.c_double
        _lib.ref_toefl_erk.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ref_toefl_ncalls.argtypes = [C.c_void_p]
        _lib.ref_toefl_multistep.argtypes = [C.c_void_p, C.c_char_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p]
        _lib.ref_toefl_adaptive.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_toefl_helmholtz_solve.argtypes = [C.c_void_p] * 4
        _lib.ref_toefl_pol_solve.argtypes = [C.c_void_p] * 5
        _lib.ref_toefl_upwind.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.ref_toefl_variation.argtypes = [C.c_void_p] * 3
        _lib.ref_toefl_arakawa.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
        _lib.ref_toefl_binv.argtypes = [C.c_void_p] * 2
    return _lib


def default_params(n=3, Nx=100, Ny=100, **over):
    """src/toefl/input/default.json with the grid overridden"""
    p = {"grid": {"n": n, "Nx": Nx, "Ny": Ny, "lx": 200, "ly": 200},
         "init": {"amplitude": 1.0, "sigma": 10, "posX": 0.3, "posY": 0.5, "flr": "gamma_inv"},
         "timestepper": {"tableau": "Bogacki-Shampine-4-2-3", "rtol": 1e-5, "atol": 1e-6},
         "bc": ["DIR", "PER"],
         "elliptic": {"stages": 3, "eps_pol": [1e-6, 1, 1], "eps_gamma": [1e-7, 1, 1], "direction": "centered"},
         "model": {"type": "global", "boussinesq": False, "curvature": 0.00015, "tau": 1, "nu": 1e-6},
         "output": {"type": "glfw", "itstp": 2}}
    for k, v in over.items():
        sec, key = k.split("__")
        p[sec][key] = v
    return p


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class RefToefl:
    def __init__(self, params):
        self.h = lib().ref_toefl_create(json.dumps(params).encode())
        self.size = lib().ref_toefl_size(self.h)

    def init(self):
        y0, y1 = np.empty(self.size), np.empty(self.size)
        lib().ref_toefl_init(self.h, _p(y0), _p(y1))
        return y0, y1

    def rhs(self, t, y0, y1):
        yp0, yp1 = np.empty(self.size), np.empty(self.size)
        sec = lib().ref_toefl_rhs(self.h, t, _p(np.ascontiguousarray(y0)), _p(np.ascontiguousarray(y1)), _p(yp0), _p(yp1))
        return yp0, yp1, sec

    def phi(self, i):
        out = np.empty(self.size)
        lib().ref_toefl_phi(self.h, i, _p(out))
        return out

    def erk(self, tableau, t0, dt, nsteps, y0, y1):
        a, b = np.array(y0, copy=True), np.array(y1, copy=True)
        sec = lib().ref_toefl_erk(self.h, tableau.encode(), t0, dt, nsteps, _p(a), _p(b))
        return a, b, sec

    def adaptive(self, tableau, t0, dt0, nsteps, rtol, atol, y0, y1):
        """nsteps calls of dg::Adaptive<ERKStep>::step (pid_control, l2norm); returns y0, y1, t, dts[], nfailed"""
        a, b = np.array(y0, copy=True), np.array(y1, copy=True)
        t, dt, dts = C.c_double(t0), C.c_double(dt0), np.zeros(nsteps)
        nf = lib().ref_toefl_adaptive(self.h, tableau.encode(), C.byref(t), C.byref(dt), nsteps, C.c_double(rtol),
                                      C.c_double(atol), _p(a), _p(b), _p(dts))
        return a, b, t.value, dts, nf

    def ncalls(self):
        return lib().ref_toefl_ncalls(self.h)

    def multistep(self, tableau, t0, dt, nsteps, y0, y1):
        """dg::ExplicitMultistep: init + nsteps steps; returns y0, y1, ts[]"""
        a, b, ts = np.array(y0, copy=True), np.array(y1, copy=True), np.zeros(nsteps)
        lib().ref_toefl_multistep(self.h, tableau.encode(), t0, dt, nsteps, _p(a), _p(b), _p(ts))
        return a, b, ts

    def helmholtz_solve(self, x, b):
        x = np.array(x, copy=True)
        num = (C.c_int * 8)()
        lib().ref_toefl_helmholtz_solve(self.h, _p(x), _p(np.ascontiguousarray(b)), num)
        return x, list(num)[:3]

    def pol_solve(self, chi, x, b):
        x = np.array(x, copy=True)
        num = (C.c_int * 8)()
        lib().ref_toefl_pol_solve(self.h, _p(np.ascontiguousarray(chi)), _p(x), _p(np.ascontiguousarray(b)), num)
        return x, list(num)[:3]

    def upwind(self, alpha, vx, vy, f, beta, result):
        r = np.array(result, copy=True)
        lib().ref_toefl_upwind(self.h, alpha, _p(np.ascontiguousarray(vx)), _p(np.ascontiguousarray(vy)), _p(np.ascontiguousarray(f)), beta, _p(r))
        return r

    def arakawa(self, alpha, lhs, rhs, beta, result):
        r = np.array(result, copy=True)
        lib().ref_toefl_arakawa(self.h, alpha, _p(np.ascontiguousarray(lhs)), _p(np.ascontiguousarray(rhs)), beta, _p(r))
        return r

    def variation(self, phi):
        out = np.empty(self.size)
        lib().ref_toefl_variation(self.h, _p(np.ascontiguousarray(phi)), _p(out))
        return out

    def binv(self):
        out = np.empty(self.size)
        lib().ref_toefl_binv(self.h, _p(out))
        return out

    def __del__(self):
        try:
            lib().ref_toefl_free(self.h)
        except Exception:
            pass
