"""dg::Elliptic2d / dg::PCG on device tensors (inc/dg/elliptic.h:233-516, inc/dg/pcg.h:25-199) through the C ABI."""
import ctypes as C
import numpy as np
import torch
from ._lib import lib, DgbError
from ._dev import ptr, stream, dvec
from . import blas1, topology as T

d = C.c_double


class Elliptic2d:
    """-div( chi grad ) on a Cartesian 2d grid; same constructor arguments as elliptic.h:281-283"""

    def __init__(self, g, bcx=None, bcy=None, direction=T.FORWARD, jfactor=1.0, chi_weight_jump=False):
        bcx = g.bc[0] if bcx is None else bcx
        bcy = g.bc[1] if bcy is None else bcy
        self.grid = g
        self.mats = dict(
            leftx=T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(direction)),
            lefty=T.derivative(1, g, T.inverse_bc(bcy), T.inverse_dir(direction)),
            rightx=T.derivative(0, g, bcx, direction), righty=T.derivative(1, g, bcy, direction),
            jumpx=T.jump(0, g, bcx), jumpy=T.jump(1, g, bcy))
        hs = {k: m.host_struct() for k, m in self.mats.items()}
        self.h = C.c_void_p()
        lib().elliptic2d_create(C.byref(self.h), *[C.byref(hs[k]) for k in ("leftx", "lefty", "rightx", "righty",
                                                                           "jumpx", "jumpy")], d(jfactor),
                                int(chi_weight_jump))
        self.size = g.size
        self.jfactor = jfactor
        self._weights = dvec(g.weights())                      # create::volume == weights on a Cartesian grid
        self._precond = torch.ones(self.size, dtype=torch.float64, device="cuda")
        self._vol = None                                       # Cartesian: tensor::volume == 1 (not stored)
        self._sigma = torch.ones(self.size, dtype=torch.float64, device="cuda")
        lib().elliptic2d_set_sigma(self.h, ptr(self._sigma))

    @property
    def fused(self):
        f = C.c_int()
        lib().elliptic2d_size(self.h, None, C.byref(f))
        return bool(f.value)

    KERNELS = {"auto": 0, "tile": 1, "walker": 2, "unfused": 3}

    def set_kernel(self, kernel):
        """dgb_elliptic2d_set_kernel: "auto" | "tile" | "walker" | "unfused" -- every mode gives bitwise the same result"""
        lib().elliptic2d_set_kernel(self.h, self.KERNELS[kernel])
        return self

    def kernel(self, with_dot=False):
        """the kernel the next symv (or PCG iteration, with_dot=True) on this plan launches"""
        k = C.c_int()
        lib().elliptic2d_get_kernel(self.h, int(with_dot), C.byref(k))
        return {v: n for n, v in self.KERNELS.items()}[k.value]

    def set_ordering(self, ordering):
        """dgb_elliptic2d_set_ordering: "reference" (default, bitwise the reference's rounding sequence) | "relaxed" (one FMA
        chain per output in the interior rows of the walker kernel: faster, equal to <= 1e-13 relative)"""
        lib().elliptic2d_set_ordering(self.h, {"reference": 0, "relaxed": 1}[ordering])
        return self

    def set_vol(self, vol):
        """curvilinear volume form m_vol (elliptic.h:292-296); the tensor stays borrowed by the plan"""
        self._vol = vol
        lib().elliptic2d_set_vol(self.h, ptr(vol) if vol is not None else None)

    def weights(self):
        return self._weights

    def precond(self):
        return self._precond

    def set_chi(self, sigma):
        """elliptic.h:324-333: m_sigma = sigma*vol, precond = 1/sigma"""
        if self._vol is None:
            blas1.copy(sigma, self._sigma)  # sigma * 1. is exact
        else:
            blas1.pointwiseDot(sigma, self._vol, self._sigma)
        one = torch.ones_like(sigma)
        blas1.pointwiseDivide(one, sigma, self._precond)

    def set_jfactor(self, jf):
        self.jfactor = jf
        lib().elliptic2d_set_jfactor(self.h, d(jf))

    def symv(self, *a, unfused=False):
        """symv(x, y) | symv(alpha, x, beta, y)  (elliptic.h:402-458)"""
        alpha, x, beta, y = (1., a[0], 0., a[1]) if len(a) == 2 else a
        if x.numel() != self.size or y.numel() != self.size:
            raise ValueError("dg::Error: vector size does not match the operator")
        fn = lib().elliptic2d_symv_unfused if unfused else lib().elliptic2d_symv
        fn(self.h, d(alpha), ptr(x), d(beta), ptr(y), stream())

    def __del__(self):
        try:
            lib().elliptic2d_destroy(self.h)
        except Exception:
            pass


class Elliptic1d:
    """dg::Elliptic1d (inc/dg/elliptic.h:65-200): -d/dx(chi d/dx) + jump on a 1-d grid -- three Ell symv and one pointwiseDot,
    the same calls in the same order as the reference template"""

    def __init__(self, g, bcx=None, direction=T.FORWARD, jfactor=1.0):
        assert g.ndim == 1
        bcx = g.bc[0] if bcx is None else bcx
        self.leftx = T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(direction))
        self.rightx = T.derivative(0, g, bcx, direction)
        self.jumpx = T.jump(0, g, bcx)
        self.size, self.jfactor = g.size, jfactor
        self._weights = dvec(g.weights())
        self._precond = torch.ones(self.size, dtype=torch.float64, device="cuda")
        self._sigma = torch.ones_like(self._precond)
        self._tempx = torch.ones_like(self._precond)

    def weights(self):
        return self._weights

    def precond(self):
        return self._precond

    def set_chi(self, sigma):
        blas1.copy(sigma, self._sigma)
        blas1.pointwiseDivide(torch.ones_like(sigma), sigma, self._precond)

    def symv(self, *a):
        alpha, x, beta, y = (1., a[0], 0., a[1]) if len(a) == 2 else a
        self.rightx.symv(1., x, 0., self._tempx)
        blas1.pointwiseDot(self._tempx, self._sigma, self._tempx)
        self.leftx.symv(-alpha, self._tempx, beta, y)
        if self.jfactor != 0.:
            self.jumpx.symv(self.jfactor * alpha, x, 1., y)


class Elliptic3d:
    """dg::Elliptic3d with set_compute_in_2d(true) (inc/dg/elliptic.h:557-797, the mode src/feltor/feltor.h uses) on a
    CartesianGrid3d or a CylindricalGrid3d (x = R, y = Z, z = phi: vol = R): the 2-d plan applied to every plane."""

    def __init__(self, g, direction=T.FORWARD, jfactor=1.0, chi_weight_jump=False, cylindrical=False):
        assert g.ndim == 3
        self.grid = g
        self.perp = T.Grid(g.x0[:2], g.x1[:2], g.n[0], g.N[:2], g.bc[:2])
        self.op = Elliptic2d(self.perp, direction=direction, jfactor=jfactor, chi_weight_jump=chi_weight_jump)
        self.nplanes, self.size = g.shape(2), g.size
        w = g.weights()
        if cylindrical:
            R = np.ascontiguousarray(np.broadcast_to(g.abscissas(0), (g.shape(1), g.shape(0))).reshape(-1))
            # the volume form as the reference computes it: g^pp = 1/R/R (base_geometry.h:336-344), vol = 1/sqrt(det)
            # (multiply.h:389, functors.h:82-88) -- R up to rounding, and the roundings must be the same
            R = 1. / np.sqrt((1. / R) / R)
            self._vol2d = dvec(R)
            lib().elliptic2d_set_vol(self.op.h, ptr(self._vol2d))
            self._vol = dvec(np.tile(R, self.nplanes))
            w = w * np.tile(R, self.nplanes)          # create::volume = tensor::volume(metric) * weights
        else:
            self._vol2d, self._vol = None, None
        self._weights = dvec(w)
        self._precond = torch.ones(self.size, dtype=torch.float64, device="cuda")
        self._sigma = self._vol.clone() if cylindrical else torch.ones(self.size, dtype=torch.float64, device="cuda")

    def weights(self):
        return self._weights

    def precond(self):
        return self._precond

    def set_chi(self, sigma):
        """elliptic.h:636-645: m_sigma = sigma*vol, precond = 1/sigma"""
        if self._vol is None:
            blas1.copy(sigma, self._sigma)
        else:
            blas1.pointwiseDot(sigma, self._vol, self._sigma)
        blas1.pointwiseDivide(torch.ones_like(sigma), sigma, self._precond)

    def symv(self, *a):
        alpha, x, beta, y = (1., a[0], 0., a[1]) if len(a) == 2 else a
        if x.numel() != self.size or y.numel() != self.size:
            raise ValueError("dg::Error: vector size does not match the operator")
        lib().elliptic2d_symv_planes(self.op.h, self.nplanes, ptr(self._sigma), d(alpha), ptr(x), d(beta), ptr(y), stream())


class PCG:
    """dg::PCG<DVec> (pcg.h:25-199): solve(A, x, b, P, W, eps, nrmb_correction, test_frequency) -> iterations"""

    def __init__(self, size, max_iterations):
        self.size, self.max_iter = size, max_iterations
        self.throw_on_fail = True
        self.h = C.c_void_p()
        lib().pcg_create(C.byref(self.h), size)

    def set_max(self, m):
        self.max_iter = m

    def set_throw_on_fail(self, v):
        self.throw_on_fail = v

    def solve(self, A, x, b, P, W, eps=1e-12, nrmb_correction=1.0, test_frequency=1):
        it = C.c_int()
        try:
            lib().pcg_solve_elliptic2d(self.h, A.h, ptr(x), ptr(b), ptr(P), ptr(W), d(eps), d(nrmb_correction),
                                       test_frequency, self.max_iter, C.byref(it), stream())
        except DgbError as e:
            if e.code == -4 and not self.throw_on_fail:   # dg::Fail, pcg.h:189-193
                return it.value
            raise
        return it.value

    def __del__(self):
        try:
            lib().pcg_destroy(self.h)
        except Exception:
            pass


class MultigridCG2d:
    """dg::MultigridCG2d (inc/dg/multigrid.h:500-668): nested grids, projection and the nested-iteration solve"""

    def __init__(self, g, stages):
        self.h = C.c_void_p()
        lib().multigrid2d_create(C.byref(self.h), g.ref(), stages)
        self.stages = stages
        self.grids, self.sizes = [], []
        for u in range(stages):
            cg, sz = T.CGrid(), C.c_size_t()
            lib().multigrid2d_grid(self.h, u, C.byref(cg), C.byref(sz))
            self.grids.append(T.Grid([cg.x0[0], cg.x0[1]], [cg.x1[0], cg.x1[1]], [cg.n[0], cg.n[1]], [cg.N[0], cg.N[1]],
                                     [cg.bc[0], cg.bc[1]]))
            self.sizes.append(sz.value)

    def grid(self, u):
        return self.grids[u]

    def project(self, src):
        """multigrid.h:94-110: returns the list of the projections of src onto every stage"""
        out = [torch.empty(s, dtype=torch.float64, device="cuda") for s in self.sizes]
        arr = (C.c_void_p * self.stages)(*[o.data_ptr() for o in out])
        lib().multigrid2d_project(self.h, ptr(src), arr, stream())
        return out

    def interpolate(self, coarse_stage, alpha, xc, beta, xf):
        """multigrid.h:232: xf = alpha * interpolation(coarse_stage - 1) xc + beta xf"""
        lib().multigrid2d_interpolate(self.h, coarse_stage, d(alpha), ptr(xc), d(beta), ptr(xf), stream())

    def solve(self, ops, x, b, eps):
        """multigrid.h:617-658: ops = list of Elliptic2d (one per stage); eps scalar or list; returns iteration numbers"""
        eps = [eps] * self.stages if np.isscalar(eps) else list(eps)
        if len(ops) != self.stages or len(eps) != self.stages:
            raise ValueError("dg::Error: MultigridCG2d::solve needs one operator and one accuracy per stage (%d)" % self.stages)
        A = (C.c_void_p * self.stages)(*[o.h.value for o in ops])
        P = (C.c_void_p * self.stages)(*[o.precond().data_ptr() for o in ops])
        W = (C.c_void_p * self.stages)(*[o.weights().data_ptr() for o in ops])
        E = (C.c_double * self.stages)(*eps)
        num = (C.c_int * self.stages)()
        lib().multigrid2d_solve(self.h, A, P, W, ptr(x), ptr(b), E, num, stream())
        return [int(v) for v in num]

    def __del__(self):
        try:
            lib().multigrid2d_destroy(self.h)
        except Exception:
            pass
