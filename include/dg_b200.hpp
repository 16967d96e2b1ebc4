// dg_b200.hpp -- header-only C++17 host layer over the C ABI of libdgb200.so.
// Mirrors the names, argument order and error behaviour of the reference's public interface for the hot path
// (dg::blas1 / dg::blas2 / dg::Elliptic2d / dg::PCG / dg::MultigridCG2d, inc/dg/blas1.h, blas2.h, elliptic.h, pcg.h,
// multigrid.h) on an RAII device vector.  No arithmetic lives here: every call forwards to a dgb_* entry point.
// A Feltor tree would instead plug the C ABI into its own dispatch seams; this header is the
// stand-alone way to program against the library from C++.
#pragma once
#include <algorithm>
#include <array>
#include <type_traits>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "dgb200.h"

namespace dgb200 {

struct Error : std::runtime_error {  // dg::Error (inc/dg/backend/exceptions.h)
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct Fail : Error {  // dg::Fail: solver did not converge (pcg.h:189-193)
    using Error::Error;
};
inline void check(int code) {
    if (code == DGB_OK) return;
    std::string msg = std::string("dgb200: ") + dgb_last_error();
    if (code == DGB_ERR_NOCONVERGE) throw Fail(code, msg);
    throw Error(code, msg);
}

enum bc { PER = DGB_PER, DIR = DGB_DIR, DIR_NEU = DGB_DIR_NEU, NEU_DIR = DGB_NEU_DIR, NEU = DGB_NEU };
enum direction { forward = DGB_FORWARD, backward = DGB_BACKWARD, centered = DGB_CENTERED };
inline bc inverse(bc b) {  // inc/dg/enums.h:62-78
    switch (b) { case DIR: return NEU; case NEU: return DIR; case DIR_NEU: return NEU_DIR; case NEU_DIR: return DIR_NEU; default: return PER; }
}
inline direction inverse(direction d) { return d == forward ? backward : (d == backward ? forward : centered); }

using HVec = std::vector<double>;

// thrust::device_vector<double> stand-in (dg::DVec, inc/dg/backend/typedefs.h:24)
class DVec {
    double* m_p = nullptr;
    size_t m_n = 0;
  public:
    DVec() = default;
    explicit DVec(size_t n, double value = 0.) { resize(n); if (n) check(dgb_fill(n, value, m_p, nullptr)); }
    DVec(const HVec& h) { resize(h.size()); if (m_n) { check(dgb_memcpy_h2d(m_p, h.data(), m_n * sizeof(double), nullptr)); check(dgb_stream_synchronize(nullptr)); } }
    DVec(const DVec& o) { resize(o.m_n); if (m_n) check(dgb_memcpy_d2d(m_p, o.m_p, m_n * sizeof(double), nullptr)); }
    DVec(DVec&& o) noexcept { swap(o); }
    DVec& operator=(DVec o) { swap(o); return *this; }
    ~DVec() { if (m_p) dgb_free(m_p); }
    void swap(DVec& o) noexcept { std::swap(m_p, o.m_p); std::swap(m_n, o.m_n); }
    void resize(size_t n) { if (n == m_n) return; if (m_p) dgb_free(m_p); m_p = nullptr; m_n = n; if (n) check(dgb_malloc((void**)&m_p, n * sizeof(double))); }
    size_t size() const { return m_n; }
    double* data() { return m_p; }
    const double* data() const { return m_p; }
    // element read from the host like thrust::device_vector's v[i] (blas1_t.cpp:26-38 does this); one blocking 8-byte copy
    double operator[](size_t i) const {
        double v;
        check(dgb_memcpy_d2h(&v, m_p + i, sizeof(double), nullptr));
        check(dgb_stream_synchronize(nullptr));
        return v;
    }
    void assign(size_t n, double value) { resize(n); if (n) check(dgb_fill(n, value, m_p, nullptr)); }
    HVec to_host() const { HVec h(m_n); if (m_n) { check(dgb_memcpy_d2h(h.data(), m_p, m_n * sizeof(double), nullptr)); check(dgb_stream_synchronize(nullptr)); } return h; }
};

// dg::RealGrid<double,2> / dg::CartesianGrid2d (inc/dg/topology/grid.h)
struct Grid2d {
    dgb_grid g{};
    Grid2d(double x0, double x1, double y0, double y1, unsigned n, unsigned Nx, unsigned Ny, bc bcx = PER, bc bcy = PER) {
        g.ndim = 2; g.x0[0] = x0; g.x1[0] = x1; g.x0[1] = y0; g.x1[1] = y1;
        g.n[0] = g.n[1] = (int)n; g.N[0] = (int)Nx; g.N[1] = (int)Ny; g.bc[0] = bcx; g.bc[1] = bcy;
    }
    explicit Grid2d(const dgb_grid& gg) : g(gg) {}
    size_t size() const { size_t s; check(dgb_topo_size(&g, &s)); return s; }
    bc bcx() const { return (bc)g.bc[0]; }
    bc bcy() const { return (bc)g.bc[1]; }
    HVec abscissas(int u) const { HVec a((size_t)g.n[u] * g.N[u]); check(dgb_topo_abscissas(&g, u, a.data())); return a; }
};
namespace create {
inline HVec weights(const Grid2d& g) { HVec w(g.size()); check(dgb_topo_weights(&g.g, w.data())); return w; }  // weights.h:60
}
template <class F>
HVec evaluate(F f, const Grid2d& g) {  // dg::evaluate, inc/dg/topology/evaluation.h:74
    HVec ax = g.abscissas(0), ay = g.abscissas(1), v(g.size());
    for (size_t j = 0; j < ay.size(); j++)
        for (size_t i = 0; i < ax.size(); i++) v[j * ax.size() + i] = f(ax[i], ay[j]);
    return v;
}

namespace blas1 {  // inc/dg/blas1.h (call-site shortcuts included)
inline void copy(const DVec& x, DVec& y) { check(dgb_copy(x.size(), x.data(), y.data(), nullptr)); }
inline void copy(double a, DVec& y) { check(dgb_fill(y.size(), a, y.data(), nullptr)); }
inline void scal(DVec& x, double a) { if (a != 1.) check(dgb_scal(x.size(), x.data(), a, nullptr)); }
inline void plus(DVec& x, double a) { if (a != 0.) check(dgb_plus(x.size(), x.data(), a, nullptr)); }
inline void axpby(double a, const DVec& x, double b, DVec& y) {
    if (a == 0.) return scal(y, b);
    if (&x == &y) return scal(y, a + b);
    check(dgb_axpby(x.size(), a, x.data(), b, y.data(), nullptr));
}
inline void axpby(double a, const DVec& x, double b, const DVec& y, DVec& z) { check(dgb_axpbyz(x.size(), a, x.data(), b, y.data(), z.data(), nullptr)); }
inline void axpbypgz(double a, const DVec& x, double b, const DVec& y, double g, DVec& z) {
    if (a == 0.) return axpby(b, y, g, z);
    if (b == 0.) return axpby(a, x, g, z);
    check(dgb_axpbypgz(x.size(), a, x.data(), b, y.data(), g, z.data(), nullptr));
}
inline void pointwiseDot(const DVec& x1, const DVec& x2, DVec& y) { check(dgb_pointwise_dot_xy(y.size(), x1.data(), x2.data(), y.data(), nullptr)); }
inline void pointwiseDot(double a, const DVec& x1, const DVec& x2, double b, DVec& y) {
    if (a == 0.) return scal(y, b);
    check(dgb_pointwise_dot(y.size(), a, x1.data(), x2.data(), b, y.data(), nullptr));
}
inline void pointwiseDivide(const DVec& x1, const DVec& x2, DVec& y) { check(dgb_pointwise_divide_xy(y.size(), x1.data(), x2.data(), y.data(), nullptr)); }
inline void pointwiseDivide(double a, const DVec& x1, const DVec& x2, double b, DVec& y) {
    if (a == 0.) return scal(y, b);
    check(dgb_pointwise_divide(y.size(), a, x1.data(), x2.data(), b, y.data(), nullptr));
}
namespace detail {
inline dgb_dot_ws* ws() { static dgb_dot_ws* w = nullptr; if (!w) check(dgb_dot_ws_create(&w)); return w; }
}
inline double dot(const DVec& x, const DVec& y) {  // blas1.h:152-170; throws on NaN/Inf like blas1.h:161
    double v = 0; int st = 0;
    check(dgb_dot2(detail::ws(), x.size(), x.data(), y.data(), nullptr, &v, &st, nullptr));
    return v;
}
}  // namespace blas1
namespace blas2 {  // inc/dg/blas2.h
inline double dot(const DVec& x, const DVec& w, const DVec& y) {
    double v = 0; int st = 0;
    check(dgb_dot3(blas1::detail::ws(), x.size(), x.data(), w.data(), y.data(), nullptr, &v, &st, nullptr));
    return v;
}
inline double dot(const DVec& w, const DVec& x) { return dot(x, w, x); }
}  // namespace blas2

// dg::DMatrix = EllSparseBlockMat<double, thrust::device_vector> (inc/dg/backend/sparseblockmat.h:44-188)
class DMatrix {
    dgb_ell* m_m = nullptr;
  public:
    DMatrix() = default;
    explicit DMatrix(dgb_ellh* host) {  // takes ownership of the host description
        dgb_ell_host v;
        check(dgb_ellh_view(host, &v));
        int e = dgb_ell_create(&m_m, &v);
        dgb_ellh_destroy(host);
        check(e);
    }
    DMatrix(const DMatrix&) = delete;
    DMatrix(DMatrix&& o) noexcept { std::swap(m_m, o.m_m); }
    DMatrix& operator=(DMatrix&& o) noexcept { std::swap(m_m, o.m_m); return *this; }
    ~DMatrix() { if (m_m) dgb_ell_destroy(m_m); }
    void symv(double alpha, const DVec& x, double beta, DVec& y) const { check(dgb_ell_symv(m_m, alpha, x.data(), beta, y.data(), nullptr)); }
};
namespace create {  // inc/dg/topology/derivatives.h:36-60
inline DMatrix dx(const Grid2d& g, bc b, direction d = centered) { dgb_ellh* m; check(dgb_topo_derivative(&m, &g.g, 0, b, d)); return DMatrix(m); }
inline DMatrix dy(const Grid2d& g, bc b, direction d = centered) { dgb_ellh* m; check(dgb_topo_derivative(&m, &g.g, 1, b, d)); return DMatrix(m); }
}
namespace blas2 {
inline void symv(const DMatrix& m, const DVec& x, DVec& y) { m.symv(1., x, 0., y); }
inline void symv(double alpha, const DMatrix& m, const DVec& x, double beta, DVec& y) { m.symv(alpha, x, beta, y); }
}
namespace blas1 {
inline void pointwiseDot(double a, const DVec& x1, const DVec& x2, const DVec& x3, double b, DVec& y) {  // blas1.h:457
    if (a == 0.) return scal(y, b);
    check(dgb_pointwise_dot3(y.size(), a, x1.data(), x2.data(), x3.data(), b, y.data(), nullptr));
}
enum class reduce_op { sum = DGB_REDUCE_SUM, max = DGB_REDUCE_MAX, min = DGB_REDUCE_MIN, logical_or = DGB_REDUCE_OR };
enum class unary_op { identity = DGB_UNARY_IDENTITY, abs = DGB_UNARY_ABS, square = DGB_UNARY_SQUARE, isnan = DGB_UNARY_ISNAN, isnotfinite = DGB_UNARY_ISNOTFINITE };
inline double reduce(const DVec& x, double init, reduce_op op, unary_op u = unary_op::identity) {  // blas1.h:213-223, closed functor set
    double out = 0;
    check(dgb_reduce(x.size(), x.data(), (int)op, (int)u, init, &out, nullptr));
    return out;
}
}

// dg::Elliptic2d<CartesianGrid2d, DMatrix, DVec> (inc/dg/elliptic.h:233-516)
class Elliptic2d {
    dgb_elliptic2d* m_plan = nullptr;
    DVec m_weights, m_precond, m_sigma;
    static dgb_ellh* dx(const dgb_grid& g, int coord, bc b, direction d) { dgb_ellh* m; check(dgb_topo_derivative(&m, &g, coord, b, d)); return m; }
    static dgb_ellh* jump(const dgb_grid& g, int coord, bc b) { dgb_ellh* m; check(dgb_topo_jump_nd(&m, &g, coord, b)); return m; }
  public:
    Elliptic2d() = default;
    Elliptic2d(const Grid2d& g, direction dir = forward, double jfactor = 1., bool chi_weight_jump = false)
        : Elliptic2d(g, g.bcx(), g.bcy(), dir, jfactor, chi_weight_jump) {}
    Elliptic2d(const Grid2d& g, bc bcx, bc bcy, direction dir = forward, double jfactor = 1., bool chi_weight_jump = false) {
        dgb_ellh* m[6] = {dx(g.g, 0, inverse(bcx), inverse(dir)), dx(g.g, 1, inverse(bcy), inverse(dir)), dx(g.g, 0, bcx, dir),
                          dx(g.g, 1, bcy, dir), jump(g.g, 0, bcx), jump(g.g, 1, bcy)};  // elliptic.h:285-290
        dgb_ell_host v[6];
        for (int k = 0; k < 6; k++) check(dgb_ellh_view(m[k], &v[k]));
        int e = dgb_elliptic2d_create(&m_plan, &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], jfactor, chi_weight_jump);
        for (int k = 0; k < 6; k++) dgb_ellh_destroy(m[k]);
        check(e);
        m_weights = DVec(create::weights(g));
        m_precond = DVec(g.size(), 1.);
        m_sigma = DVec(g.size(), 1.);
        check(dgb_elliptic2d_set_sigma(m_plan, m_sigma.data()));
    }
    Elliptic2d(const Elliptic2d&) = delete;
    Elliptic2d(Elliptic2d&& o) noexcept { *this = std::move(o); }
    Elliptic2d& operator=(Elliptic2d&& o) noexcept {
        std::swap(m_plan, o.m_plan); m_weights.swap(o.m_weights); m_precond.swap(o.m_precond); m_sigma.swap(o.m_sigma);
        return *this;
    }
    ~Elliptic2d() { if (m_plan) dgb_elliptic2d_destroy(m_plan); }
    const DVec& weights() const { return m_weights; }
    const DVec& precond() const { return m_precond; }
    void set_chi(const DVec& sigma) {  // elliptic.h:324-333 (Cartesian: vol == 1)
        blas1::copy(sigma, m_sigma);
        DVec one(sigma.size(), 1.);
        blas1::pointwiseDivide(one, sigma, m_precond);
    }
    void set_jfactor(double j) { check(dgb_elliptic2d_set_jfactor(m_plan, j)); }
    void symv(const DVec& x, DVec& y) { symv(1., x, 0., y); }
    void symv(double alpha, const DVec& x, double beta, DVec& y) { check(dgb_elliptic2d_symv(m_plan, alpha, x.data(), beta, y.data(), nullptr)); }
    // sigma = alpha lambda^2 (grad phi . chi . grad phi) + beta sigma   (elliptic.h:469-502)
    void variation(const DVec& phi, DVec& sigma) { variation(1., nullptr, phi, 0., sigma); }
    void variation(double alpha, const DVec* lambda, const DVec& phi, double beta, DVec& sigma) {
        check(dgb_elliptic2d_variation(m_plan, alpha, lambda ? lambda->data() : nullptr, phi.data(), beta, sigma.data(), nullptr));
    }
    dgb_elliptic2d* plan() { return m_plan; }
    const dgb_elliptic2d* plan() const { return m_plan; }
};

// dg::Helmholtz<Geometry, DMatrix, DVec> = GeneralHelmholtz<Elliptic2d, DVec> (inc/dg/helmholtz.h:27-95): chi x - alpha Elliptic x.
// The Elliptic plan carries the Helmholtz term, so PCG / MultigridCG2d run on it unchanged.
class Helmholtz {
    double m_alpha;
    Elliptic2d m_matrix;
    DVec m_chi;
    bool m_has_chi = false;
  public:
    Helmholtz(double alpha, Elliptic2d&& matrix) : m_alpha(alpha), m_matrix(std::move(matrix)) {
        check(dgb_elliptic2d_set_helmholtz(m_matrix.plan(), 1, m_alpha, nullptr));
    }
    Helmholtz(Helmholtz&&) = default;
    const DVec& weights() const { return m_matrix.weights(); }
    const DVec& precond() const { return m_matrix.precond(); }
    double alpha() const { return m_alpha; }
    void set_chi(const DVec& chi) {
        m_chi = chi; m_has_chi = true;
        check(dgb_elliptic2d_set_helmholtz(m_matrix.plan(), 1, m_alpha, m_chi.data()));
    }
    void symv(const DVec& x, DVec& y) { m_matrix.symv(x, y); }
    Elliptic2d& matrix() { return m_matrix; }
    dgb_elliptic2d* plan() { return m_matrix.plan(); }
};

// dg::Advection<Geometry, DMatrix, DVec> (inc/dg/advection.h:60-120)
class Advection {
    DVec m_temp0, m_temp1;
    DMatrix m_dxf, m_dyf, m_dxb, m_dyb;
  public:
    explicit Advection(const Grid2d& g) : Advection(g, g.bcx(), g.bcy()) {}
    Advection(const Grid2d& g, bc bcx, bc bcy)
        : m_temp0(g.size(), 1.), m_temp1(g.size(), 1.), m_dxf(create::dx(g, bcx, forward)), m_dyf(create::dy(g, bcy, forward)),
          m_dxb(create::dx(g, bcx, backward)), m_dyb(create::dy(g, bcy, backward)) {}
    void upwind(double alpha, const DVec& vx, const DVec& vy, const DVec& f, double beta, DVec& result) {
        blas2::symv(m_dxb, f, m_temp0);
        blas2::symv(m_dxf, f, m_temp1);
        check(dgb_upwind_axpby(f.size(), alpha, vx.data(), m_temp0.data(), m_temp1.data(), beta, result.data(), nullptr));
        blas2::symv(m_dyb, f, m_temp0);
        blas2::symv(m_dyf, f, m_temp1);
        check(dgb_upwind_axpby(f.size(), alpha, vy.data(), m_temp0.data(), m_temp1.data(), 1., result.data(), nullptr));
    }
};

// dg::ArakawaX<Geometry, DMatrix, DVec> (inc/dg/arakawa.h:30-170), Cartesian grids (perpendicular volume 1)
class ArakawaX {
    DVec m_dxlhs, m_dxrhs, m_dylhs, m_dyrhs, m_chi;
    DMatrix m_bdxf, m_bdyf;
  public:
    explicit ArakawaX(const Grid2d& g) : ArakawaX(g, g.bcx(), g.bcy()) {}
    ArakawaX(const Grid2d& g, bc bcx, bc bcy)
        : m_dxlhs(g.size(), 1.), m_dxrhs(g.size(), 1.), m_dylhs(g.size(), 1.), m_dyrhs(g.size(), 1.), m_chi(g.size(), 1.),
          m_bdxf(create::dx(g, bcx, centered)), m_bdyf(create::dy(g, bcy, centered)) {}
    void operator()(const DVec& lhs, const DVec& rhs, DVec& result) { (*this)(1., lhs, rhs, 0., result); }
    void operator()(double alpha, const DVec& lhs, const DVec& rhs, double beta, DVec& result) {
        blas2::symv(m_bdxf, lhs, m_dxlhs);
        blas2::symv(m_bdyf, lhs, m_dylhs);
        blas2::symv(m_bdxf, rhs, m_dxrhs);
        blas2::symv(m_bdyf, rhs, m_dyrhs);
        check(dgb_arakawa_functor(lhs.size(), lhs.data(), rhs.data(), m_dxlhs.data(), m_dylhs.data(), m_dxrhs.data(), m_dyrhs.data(), nullptr));
        blas2::symv(1., m_bdxf, m_dylhs, 1., m_dyrhs);
        blas2::symv(1., m_bdyf, m_dxrhs, 1., m_dyrhs);
        blas1::pointwiseDot(alpha, m_chi, m_dyrhs, beta, result);
    }
};

// dg::Extrapolation<DVec> (inc/dg/extrapolation.h:225-460), constant / linear extrapolation
class Extrapolation {
    unsigned m_max = 0, m_counter = 0;
    std::vector<DVec> m_x;
    std::vector<double> m_t;
  public:
    Extrapolation() = default;
    Extrapolation(unsigned max, const DVec& copyable) { set_max(max, copyable); }
    void set_max(unsigned max, const DVec& copyable) { m_counter = 0; m_x.assign(max, copyable); m_t.assign(max, 0.); m_max = max; }
    void extrapolate(double t, DVec& new_x) const {
        if (m_counter == 0) return blas1::copy(0., new_x);
        if (m_counter == 1) return blas1::copy(m_x[0], new_x);
        if (m_counter == 3) throw Error(DGB_ERR_UNSUPPORTED, "dgb200: parabolic extrapolation is not implemented");
        double f0 = (t - m_t[1]) / (m_t[0] - m_t[1]), f1 = (t - m_t[0]) / (m_t[1] - m_t[0]);
        blas1::axpby(f0, m_x[0], f1, m_x[1], new_x);
    }
    void update(double t_new, const DVec& new_entry) {
        if (m_max == 0) return;
        for (unsigned i = 0; i < m_counter; i++)
            if (std::abs(t_new - m_t[i]) < 1e-14) { blas1::copy(new_entry, m_x[i]); return; }
        if (m_counter < m_max) m_counter++;
        std::rotate(m_x.rbegin(), m_x.rbegin() + 1, m_x.rend());
        std::rotate(m_t.rbegin(), m_t.rbegin() + 1, m_t.rend());
        m_t[0] = t_new;
        blas1::copy(new_entry, m_x[0]);
    }
};

// dg::PCG<DVec> (inc/dg/pcg.h:25-199)
class PCG {
    dgb_pcg* m_pcg = nullptr;
    unsigned m_max = 0;
    bool m_throw = true;
    DVec m_r, m_p, m_ap;  // work vectors of the generic solve (allocated on first use)
    // blas2::symv(M, x, y) for the operator kinds the generic solve accepts: a diagonal DVec, anything with
    // symv(x, y), or a callable (x, y)
    template <class M>
    static void apply(M&& m, const DVec& x, DVec& y) {
        using T = std::decay_t<M>;
        if constexpr (std::is_same<T, DVec>::value) blas1::pointwiseDot(m, x, y);
        else if constexpr (std::is_invocable<M, const DVec&, DVec&>::value) m(x, y);
        else m.symv(x, y);
    }
  public:
    PCG() = default;
    PCG(const DVec& copyable, unsigned max_iterations) : m_max(max_iterations) { check(dgb_pcg_create(&m_pcg, copyable.size())); }
    PCG(const PCG&) = delete;
    PCG(PCG&& o) noexcept { std::swap(m_pcg, o.m_pcg); m_max = o.m_max; m_throw = o.m_throw; m_r.swap(o.m_r); m_p.swap(o.m_p); m_ap.swap(o.m_ap); }
    ~PCG() { if (m_pcg) dgb_pcg_destroy(m_pcg); }
    void set_max(unsigned m) { m_max = m; }
    unsigned get_max() const { return m_max; }
    void set_throw_on_fail(bool t) { m_throw = t; }
    unsigned solve(Elliptic2d& A, DVec& x, const DVec& b, const DVec& P, const DVec& W, double eps = 1e-12,
                   double nrmb_correction = 1., int test_frequency = 1) {
        int it = 0;
        int e = dgb_pcg_solve_elliptic2d(m_pcg, A.plan(), x.data(), b.data(), P.data(), W.data(), eps, nrmb_correction,
                                         test_frequency, (int)m_max, &it, nullptr);
        if (e == DGB_ERR_NOCONVERGE && !m_throw) return (unsigned)it;
        check(e);
        return (unsigned)it;
    }
    unsigned solve(Helmholtz& A, DVec& x, const DVec& b, const DVec& P, const DVec& W, double eps = 1e-12,
                   double nrmb_correction = 1., int test_frequency = 1) {
        return solve(A.matrix(), x, b, P, W, eps, nrmb_correction, test_frequency);  // the plan carries the Helmholtz term
    }
    // Any other self-adjoint operator / preconditioner (pcg.h:136-195): the same weighted iteration, one blas call per
    // step instead of the three fused kernels.  Bit-identical to the fused solve when A is an Elliptic2d.
    template <class Matrix, class Preconditioner>
    unsigned solve(Matrix&& A, DVec& x, const DVec& b, Preconditioner&& P, const DVec& W, double eps = 1e-12,
                   double nrmb_correction = 1., int test_frequency = 1) {
        if (m_r.size() != x.size()) { m_r.resize(x.size()); m_p.resize(x.size()); m_ap.resize(x.size()); }
        const double nrmb = std::sqrt(blas2::dot(W, b)), tol = eps * (nrmb + nrmb_correction);
        if (nrmb == 0) { blas1::copy(0., x); return 0; }
        apply(A, x, m_r);
        blas1::axpby(1., b, -1., m_r);
        if (std::sqrt(blas2::dot(W, m_r)) < tol) return 0;
        apply(P, m_r, m_p);
        double rz = blas2::dot(m_p, W, m_r);
        for (unsigned i = 1; i < m_max; i++) {
            apply(A, m_p, m_ap);
            const double alpha = rz / blas2::dot(m_p, W, m_ap);
            blas1::axpby(alpha, m_p, 1., x);
            blas1::axpby(-alpha, m_ap, 1., m_r);
            if (i % test_frequency == 0 && std::sqrt(blas2::dot(W, m_r)) < tol) return i;
            apply(P, m_r, m_ap);
            const double rz_new = blas2::dot(m_ap, W, m_r);
            blas1::axpby(1., m_ap, rz_new / rz, m_p);
            rz = rz_new;
        }
        if (m_throw) throw Fail(DGB_ERR_NOCONVERGE, "dg::Fail: PCG did not converge within max_iterations");
        return m_max;
    }
};

// dg::MultigridCG2d<Geometry, DMatrix, DVec> (inc/dg/multigrid.h:500-668)
class MultigridCG2d {
    dgb_multigrid2d* m_mg = nullptr;
    unsigned m_stages = 0;
  public:
    MultigridCG2d(const Grid2d& g, unsigned stages) : m_stages(stages) { check(dgb_multigrid2d_create(&m_mg, &g.g, (int)stages)); }
    MultigridCG2d(const MultigridCG2d&) = delete;
    ~MultigridCG2d() { if (m_mg) dgb_multigrid2d_destroy(m_mg); }
    unsigned stages() const { return m_stages; }
    Grid2d grid(unsigned u) const { dgb_grid g; check(dgb_multigrid2d_grid(m_mg, (int)u, &g, nullptr)); return Grid2d(g); }
    std::vector<DVec> project(const DVec& src) const {
        std::vector<DVec> out;
        std::vector<double*> ptrs;
        for (unsigned u = 0; u < m_stages; u++) { out.emplace_back(grid(u).size()); }
        for (auto& v : out) ptrs.push_back(v.data());
        check(dgb_multigrid2d_project(m_mg, src.data(), ptrs.data(), nullptr));
        return out;
    }
    std::vector<unsigned> solve(std::vector<Elliptic2d>& ops, DVec& x, const DVec& b, std::vector<double> eps) {
        std::vector<dgb_elliptic2d*> A;
        std::vector<const double*> P, W;
        for (auto& o : ops) { A.push_back(o.plan()); P.push_back(o.precond().data()); W.push_back(o.weights().data()); }
        if (ops.size() != m_stages || eps.size() != m_stages) throw Error(DGB_ERR_INVALID, "dg::Error: MultigridCG2d::solve needs one operator and one accuracy per stage");
        std::vector<int> num(m_stages);
        check(dgb_multigrid2d_solve(m_mg, A.data(), P.data(), W.data(), x.data(), b.data(), eps.data(), num.data(), nullptr));
        return std::vector<unsigned>(num.begin(), num.end());
    }
    std::vector<unsigned> solve(std::vector<Elliptic2d>& ops, DVec& x, const DVec& b, double eps) {
        return solve(ops, x, b, std::vector<double>(m_stages, eps));
    }
    std::vector<unsigned> solve(std::vector<Helmholtz>& ops, DVec& x, const DVec& b, std::vector<double> eps) {
        std::vector<dgb_elliptic2d*> A;
        std::vector<const double*> P, W;
        for (auto& o : ops) { A.push_back(o.plan()); P.push_back(o.precond().data()); W.push_back(o.weights().data()); }
        if (ops.size() != m_stages || eps.size() != m_stages) throw Error(DGB_ERR_INVALID, "dg::Error: MultigridCG2d::solve needs one operator and one accuracy per stage");
        std::vector<int> num(m_stages);
        check(dgb_multigrid2d_solve(m_mg, A.data(), P.data(), W.data(), x.data(), b.data(), eps.data(), num.data(), nullptr));
        return std::vector<unsigned>(num.begin(), num.end());
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Time steppers on std::array<DVec,2> (the container of toefl / feltor right-hand sides)
// ---------------------------------------------------------------------------------------------------------------
using DVec2 = std::array<DVec, 2>;

// blas1::dot of recursive vectors: the superaccumulators of the components are summed, normalised and rounded ONCE
// (blas1_dispatch_vector.h:153-176) -- not the sum of two rounded dots
inline double dot(const DVec2& x, const DVec2& y) {
    int64_t acc[DGB_BIN_COUNT] = {0};
    for (int q = 0; q < 2; q++) {
        int64_t part[DGB_BIN_COUNT];
        double v;
        int status = 0;
        check(dgb_dot2(blas1::detail::ws(), x[q].size(), x[q].data(), y[q].data(), part, &v, &status, nullptr));
        if (status != 0) throw Error(DGB_ERR_NOTFINITE, "dg::Error: dot product failed since one of the inputs contains NaN or Inf");
        for (int k = 0; k < DGB_BIN_COUNT; k++) acc[k] += part[k];
    }
    check(dgb_superacc_normalize_host(acc, nullptr));
    return dgb_superacc_round_host(acc);
}
inline double l2norm(const DVec2& x) { return std::sqrt(dot(x, x)); }  // adaptive.h:22

// controllers of inc/dg/adaptive.h:37-85
inline double i_control(std::array<double, 3> dt, std::array<double, 3> eps, unsigned embedded_order, unsigned) {
    return dt[0] * std::pow(eps[0], -1. / (double)embedded_order);
}
inline double pi_control(std::array<double, 3> dt, std::array<double, 3> eps, unsigned embedded_order, unsigned order) {
    if (dt[1] == 0) return i_control(dt, eps, embedded_order, order);
    double factor = std::pow(eps[0], -0.8 / (double)embedded_order) * std::pow(eps[1], 0.31 / (double)embedded_order);
    return dt[0] * factor;
}
inline double pid_control(std::array<double, 3> dt, std::array<double, 3> eps, unsigned embedded_order, unsigned order) {
    if (dt[1] == 0) return i_control(dt, eps, embedded_order, order);
    if (dt[2] == 0) return pi_control(dt, eps, embedded_order, order);
    double q = (double)embedded_order;
    double factor = std::pow(eps[0], -0.58 / q) * std::pow(eps[1], 0.21 / q) * std::pow(eps[2], -0.1 / q);
    return dt[0] * factor;
}

// dg::ButcherTableau (inc/dg/tableau.h): the embedded explicit tableaus wired up here
struct ButcherTableau {
    unsigned s, embedded_order, order;
    std::vector<double> a, b, bt, c;  // a row major s x s
    bool fsal;
    static ButcherTableau get(const std::string& name) {
        if (name == "Bogacki-Shampine-4-2-3")  // tableau.h:394-405
            return {4, 2, 3, {0, 0, 0, 0, 0.5, 0, 0, 0, 0, 0.75, 0, 0, 2. / 9., 1. / 3., 4. / 9., 0.}, {2. / 9., 1. / 3., 4. / 9., 0.},
                    {7. / 24., 1. / 4., 1. / 3., 1. / 8.}, {0., 0.5, 3. / 4., 1.}, true};
        throw Error(DGB_ERR_UNSUPPORTED, "dgb200: tableau " + name + " is not wired up");
    }
};

// dg::ERKStep<std::array<DVec,2>> (inc/dg/runge_kutta.h:300-400): embedded explicit Runge-Kutta step with FSAL; the stage
// sums are the dense gemv (blas2_densematrix.h:38-74, chunks of 8/4/2/1 columns) and EmbeddedPairSum kernels
class ERKStep {
    ButcherTableau m_rk;
    std::vector<DVec2> m_k;
    double m_t1 = 1e300;
    static void dense_gemv(double alpha, const std::vector<const double*>& cols, const double* x, double beta, DVec& y) {
        const size_t size = cols.size();
        auto pair_sum = [&](size_t first, int n, double b) {
            std::vector<double> a(x + first, x + first + n);
            check(dgb_pair_sum_axpby(y.size(), alpha, n, a.data(), cols.data() + first, b, y.data(), nullptr));
        };
        size_t i = 0;
        for (; i < size / 8; i++) pair_sum(i * 8, 8, i == 0 ? beta : 1.);
        size_t l = 0, k = 0;
        if (size % 8 >= 4) { pair_sum(i * 8, 4, size < 8 ? beta : 1.); l = 1; }
        if ((size % 8) % 4 >= 2) { pair_sum(i * 8 + l * 4, 2, size < 4 ? beta : 1.); k = 1; }
        if (((size % 8) % 4) % 2 == 1) {
            size_t j = i * 8 + l * 4 + k * 2;
            check(dgb_axpby(y.size(), alpha * x[j], cols[j], size < 2 ? beta : 1., y.data(), nullptr));
        }
    }
  public:
    ERKStep() = default;
    ERKStep(const std::string& tableau, const DVec2& copyable) : m_rk(ButcherTableau::get(tableau)), m_k(m_rk.s, copyable) {}
    unsigned order() const { return m_rk.order; }
    unsigned embedded_order() const { return m_rk.embedded_order; }
    const DVec2& copyable() const { return m_k[0]; }
    template <class RHS>
    void step(RHS& rhs, double t0, const DVec2& u0, double& t1, DVec2& u1, double dt, DVec2& delta) {
        const unsigned s = m_rk.s;
        if (t0 != m_t1) rhs(t0, u0, m_k[0]);
        for (unsigned i = 1; i < s; i++) {
            const double tu = std::fma(dt, m_rk.c[i], t0);
            for (int q = 0; q < 2; q++) {
                blas1::copy(u0[q], delta[q]);
                std::vector<const double*> cols;
                for (unsigned l = 0; l < i; l++) cols.push_back(m_k[l][q].data());
                dense_gemv(dt, cols, &m_rk.a[i * s], 1., delta[q]);
            }
            rhs(tu, delta, m_k[i]);
        }
        std::vector<double> b(s), d(s);
        for (unsigned j = 0; j < s; j++) { b[j] = dt * m_rk.b[j]; d[j] = dt * (m_rk.b[j] - m_rk.bt[j]); }
        for (int q = 0; q < 2; q++) {
            blas1::copy(u0[q], u1[q]);
            std::vector<const double*> cols;
            for (unsigned j = 0; j < s; j++) cols.push_back(m_k[j][q].data());
            check(dgb_embedded_pair_sum(u1[q].size(), u1[q].data(), delta[q].data(), 1., 0., (int)s, b.data(), d.data(), cols.data(), nullptr));
        }
        m_t1 = t1 = t0 + dt;
        if (!m_rk.fsal) rhs(t1, u1, m_k[0]);
        else std::swap(m_k[0], m_k[s - 1]);
    }
};

// dg::Adaptive<dg::ERKStep<...>> (inc/dg/adaptive.h:232-395)
class Adaptive {
    ERKStep m_stepper;
    DVec2 m_next, m_delta;
    double m_size = 0, m_eps0 = 1, m_eps1 = 1, m_eps2 = 1, m_t_next = 0, m_dt0 = 0, m_dt1 = 0, m_dt2 = 0;
    bool m_failed = false;
    unsigned m_nfailed = 0, m_nsteps = 0;
  public:
    Adaptive(const std::string& tableau, const DVec2& copyable)
        : m_stepper(tableau, copyable), m_next(copyable), m_delta(copyable), m_size((double)(copyable[0].size() + copyable[1].size())) {}
    bool failed() const { return m_failed; }
    unsigned nfailed() const { return m_nfailed; }
    unsigned nsteps() const { return m_nsteps; }
    double get_error() const { return m_eps0; }
    // u1 may alias u0 (dg::AdaptiveTimeloop calls it so)
    template <class RHS, class Control, class Norm>
    void step(RHS& rhs, double t0, const DVec2& u0, double& t1, DVec2& u1, double& dt, Control control, Norm norm, double rtol,
              double atol, double reject_limit = 2) {
        m_stepper.step(rhs, t0, u0, m_t_next, m_next, dt, m_delta);
        m_nsteps++;
        const double rs = rtol * std::sqrt(m_size), as = atol * std::sqrt(m_size);  // detail::Tolerance (adaptive.h:123-134)
        for (int q = 0; q < 2; q++) check(dgb_adaptive_tolerance(u0[q].size(), rs, as, u0[q].data(), m_delta[q].data(), nullptr));
        m_eps0 = norm(m_delta);
        m_dt0 = dt;
        if (m_eps0 > reject_limit || std::isnan(m_eps0)) {
            dt = control(std::array<double, 3>{m_dt0, 0, m_dt2}, std::array<double, 3>{m_eps0, m_eps1, m_eps2},
                         m_stepper.embedded_order(), m_stepper.order());
            if (std::fabs(dt) > 0.9 * std::fabs(m_dt0)) dt = 0.9 * m_dt0;
            m_failed = true;
            m_nfailed++;
            if (&u0 != &u1) for (int q = 0; q < 2; q++) blas1::copy(u0[q], u1[q]);
            t1 = t0;
            return;
        }
        if (m_eps0 < 1e-30) { dt = 1e14 * m_dt0; m_eps0 = 1e-30; }
        else {
            dt = control(std::array<double, 3>{m_dt0, m_dt1, m_dt2}, std::array<double, 3>{m_eps0, m_eps1, m_eps2},
                         m_stepper.embedded_order(), m_stepper.order());
            if (std::fabs(dt) > 100 * std::fabs(m_dt0)) dt = 100 * m_dt0;
        }
        m_eps2 = m_eps1; m_eps1 = m_eps0;
        m_dt2 = m_dt1; m_dt1 = m_dt0;
        for (int q = 0; q < 2; q++) blas1::copy(m_next[q], u1[q]);
        t1 = m_t_next;
        m_failed = false;
    }
};

// dg::ShuOsher<std::array<DVec,2>> with the identity limiter (inc/dg/runge_kutta.h:840-925); tableaus tableau.h:1262-1286
class ShuOsher {
    unsigned m_s = 0;
    std::vector<std::vector<double>> m_alpha, m_beta;  // lower triangles alpha(i,k), beta(i,k), k <= i
    std::vector<DVec2> m_u, m_k;
    double m_t1 = 1e300;
  public:
    ShuOsher(const std::string& tableau, const DVec2& copyable) {
        if (tableau == "SSPRK-2-2") { m_s = 2; m_alpha = {{1.}, {0.5, 0.5}}; m_beta = {{1.}, {0., 0.5}}; }
        else if (tableau == "SSPRK-3-3") { m_s = 3; m_alpha = {{1.}, {3. / 4., 1. / 4.}, {1. / 3., 0., 2. / 3.}}; m_beta = {{1.}, {0., 1. / 4.}, {0., 0., 2. / 3.}}; }
        else throw Error(DGB_ERR_UNSUPPORTED, "dgb200: Shu-Osher tableau " + tableau + " is not wired up");
        m_u.assign(m_s, copyable);
        m_k.assign(m_s, copyable);
    }
    template <class RHS>
    void step(RHS& rhs, double t0, const DVec2& u0, double& t1, DVec2& u1, double dt) {
        const unsigned s = m_s;
        std::vector<double> ts(s + 1);
        ts[0] = t0;
        for (int q = 0; q < 2; q++) blas1::copy(u0[q], m_u[0][q]);
        if (t0 != m_t1) rhs(ts[0], m_u[0], m_k[0]);
        for (unsigned i = 1; i <= s; i++) {
            DVec2& out = i == s ? u1 : m_u[i];
            for (int q = 0; q < 2; q++) blas1::axpbypgz(m_alpha[i - 1][0], m_u[0][q], dt * m_beta[i - 1][0], m_k[0][q], 0., out[q]);
            ts[i] = std::fma(m_alpha[i - 1][0], ts[0], dt * m_beta[i - 1][0]);  // one rounding, as the reference's compiler contracts it
            for (unsigned j = 1; j < i; j++) {
                for (int q = 0; q < 2; q++) blas1::axpbypgz(m_alpha[i - 1][j], m_u[j][q], dt * m_beta[i - 1][j], m_k[j][q], 1., out[q]);
                ts[i] += std::fma(m_alpha[i - 1][j], ts[j], dt * m_beta[i - 1][j]);
            }
            if (i != s) rhs(ts[i], m_u[i], m_k[i]);
            else rhs(ts[i], u1, m_k[0]);
        }
        m_t1 = t1 = ts[s];
    }
};

// dg::ExplicitMultistep<std::array<DVec,2>> (inc/dg/multistep.h:59-100; FilteredExplicitMultistep::init/step :592-639 with the
// identity filter); tableaus multistep_tableau.h:263-300
class ExplicitMultistep {
    unsigned m_order = 0, m_counter = 0;
    std::vector<double> m_a, m_b;
    std::vector<DVec2> m_u, m_f;
    double m_tu = 0, m_dt = 0;
  public:
    ExplicitMultistep(const std::string& tableau, const DVec2& copyable) {
        if (tableau == "AB-1-1") { m_order = 1; m_a = {1.}; m_b = {1.}; }
        else if (tableau == "AB-2-2") { m_order = 2; m_a = {1., 0.}; m_b = {1.5, -0.5}; }
        else if (tableau == "AB-3-3") { m_order = 3; m_a = {1., 0., 0.}; m_b = {23. / 12., -4. / 3., 5. / 12.}; }
        else if (tableau == "TVB-2-2") { m_order = 2; m_a = {4. / 3., -1. / 3.}; m_b = {4. / 3., -2. / 3.}; }
        else if (tableau == "TVB-3-3") {
            m_order = 3;
            m_a = {1.908535476882378, -1.334951446162515, 0.426415969280137};
            m_b = {1.502575553858997, -1.654746338401493, 0.670051276940255};
        } else throw Error(DGB_ERR_UNSUPPORTED, "dgb200: multistep tableau " + tableau + " is not wired up");
        m_u.assign(m_a.size(), copyable);
        m_f.assign(m_a.size(), copyable);
    }
    template <class RHS>
    void init(RHS& rhs, double t0, const DVec2& u0, double dt) {
        m_tu = t0; m_dt = dt;
        const size_t s = m_a.size();
        for (int q = 0; q < 2; q++) blas1::copy(u0[q], m_u[s - 1][q]);
        rhs(m_tu, m_u[s - 1], m_f[s - 1]);
        m_counter = 0;
    }
    template <class RHS>
    void step(RHS& rhs, double& t, DVec2& u) {
        const size_t s = m_a.size();
        if (m_counter < s - 1) {  // start-up: a Runge-Kutta step of the same order
            ShuOsher rk(m_order <= 2 ? "SSPRK-2-2" : "SSPRK-3-3", u);
            rk.step(rhs, t, u, t, u, m_dt);
            m_counter++;
            m_tu = t;
            for (int q = 0; q < 2; q++) blas1::copy(u[q], m_u[s - 1 - m_counter][q]);
            rhs(m_tu, m_u[s - 1 - m_counter], m_f[s - 1 - m_counter]);
            return;
        }
        t = m_tu = m_tu + m_dt;
        for (int q = 0; q < 2; q++) {
            blas1::axpby(m_a[0], m_u[0][q], m_dt * m_b[0], m_f[0][q], u[q]);
            for (size_t i = 1; i < s; i++) blas1::axpbypgz(m_a[i], m_u[i][q], m_dt * m_b[i], m_f[i][q], 1., u[q]);
        }
        std::rotate(m_f.rbegin(), m_f.rbegin() + 1, m_f.rend());
        std::rotate(m_u.rbegin(), m_u.rbegin() + 1, m_u.rend());
        for (int q = 0; q < 2; q++) blas1::copy(u[q], m_u[0][q]);
        rhs(m_tu, m_u[0], m_f[0]);
    }
};

// dg::CartesianGrid3d / dg::CylindricalGrid3d (inc/dg/topology/base_geometry.h:230-350): n polynomial coefficients in x and y,
// one per cell in z
struct Grid3d {
    dgb_grid g{};
    bool cylindrical = false;
    Grid3d(double x0, double x1, double y0, double y1, double z0, double z1, unsigned n, unsigned Nx, unsigned Ny, unsigned Nz,
           bc bcx = PER, bc bcy = PER, bc bcz = PER, bool cylindrical_ = false) : cylindrical(cylindrical_) {
        g.ndim = 3;
        g.x0[0] = x0; g.x1[0] = x1; g.x0[1] = y0; g.x1[1] = y1; g.x0[2] = z0; g.x1[2] = z1;
        g.n[0] = g.n[1] = (int)n; g.n[2] = 1;
        g.N[0] = (int)Nx; g.N[1] = (int)Ny; g.N[2] = (int)Nz;
        g.bc[0] = bcx; g.bc[1] = bcy; g.bc[2] = bcz;
    }
    size_t size() const { size_t s; check(dgb_topo_size(&g, &s)); return s; }
    unsigned Nz() const { return (unsigned)g.N[2]; }
    Grid2d perp_grid() const { return Grid2d(g.x0[0], g.x1[0], g.x0[1], g.x1[1], (unsigned)g.n[0], (unsigned)g.N[0], (unsigned)g.N[1], (bc)g.bc[0], (bc)g.bc[1]); }
    HVec weights() const { HVec w(size()); check(dgb_topo_weights(&g, w.data())); return w; }
};

// dg::Elliptic3d<Geometry, DMatrix, DVec> in its compute-in-2d mode (inc/dg/elliptic.h:557-797, set_compute_in_2d(true) -- the mode
// src/feltor/feltor.h uses): the 2-d plan applied to every plane; cylindrical grids carry vol = 1/sqrt(1/R/R)
class Elliptic3d {
    Elliptic2d m_perp;
    DVec m_weights, m_precond, m_sigma, m_vol, m_vol2d;
    unsigned m_planes = 0;
  public:
    Elliptic3d(const Grid3d& g, direction dir = forward, double jfactor = 1., bool chi_weight_jump = false)
        : m_perp(g.perp_grid(), dir, jfactor, chi_weight_jump), m_planes(g.Nz()) {
        HVec w = g.weights();
        const size_t n = g.size(), n2 = n / m_planes;
        m_precond = DVec(n, 1.);
        m_sigma = DVec(n, 1.);
        if (g.cylindrical) {
            HVec R = g.perp_grid().abscissas(0), vol(n), vol2d(n2);
            for (size_t i = 0; i < n2; i++) { double r = R[i % R.size()]; vol2d[i] = 1. / std::sqrt(1. / r / r); }  // base_geometry.h:336-344, multiply.h:389
            for (size_t i = 0; i < n; i++) { vol[i] = vol2d[i % n2]; w[i] *= vol[i]; }                        // create::volume
            m_vol = DVec(vol);
            m_vol2d = DVec(vol2d);
            m_sigma = m_vol;
            check(dgb_elliptic2d_set_vol(m_perp.plan(), m_vol2d.data()));
        }
        m_weights = DVec(w);
    }
    const DVec& weights() const { return m_weights; }
    const DVec& precond() const { return m_precond; }
    void set_chi(const DVec& sigma) {  // elliptic.h:636-645
        if (m_vol.size()) blas1::pointwiseDot(sigma, m_vol, m_sigma);
        else blas1::copy(sigma, m_sigma);
        DVec one(sigma.size(), 1.);
        blas1::pointwiseDivide(one, sigma, m_precond);
    }
    void symv(const DVec& x, DVec& y) { symv(1., x, 0., y); }
    void symv(double alpha, const DVec& x, double beta, DVec& y) {
        check(dgb_elliptic2d_symv_planes(m_perp.plan(), (int)m_planes, m_sigma.data(), alpha, x.data(), beta, y.data(), nullptr));
    }
};

// device array of int (CSR row offsets / column indices of stencil and interpolation matrices)
class IVec {
    int* m_p = nullptr;
    size_t m_n = 0;
  public:
    IVec() = default;
    IVec(const std::vector<int>& h) : m_n(h.size()) {
        if (!m_n) return;
        check(dgb_malloc((void**)&m_p, m_n * sizeof(int)));
        check(dgb_memcpy_h2d(m_p, h.data(), m_n * sizeof(int), nullptr));
        check(dgb_stream_synchronize(nullptr));
    }
    IVec(const IVec&) = delete;
    IVec(IVec&& o) noexcept { std::swap(m_p, o.m_p); std::swap(m_n, o.m_n); }
    ~IVec() { if (m_p) dgb_free(m_p); }
    size_t size() const { return m_n; }
    const int* data() const { return m_p; }
};

namespace blas2 {
// dg::blas2::stencil(f, M, x, y) (blas2.h:454) for the library's CSR filters (topology/filter.h:174-266)
enum class csr_filter { median = DGB_STENCIL_MEDIAN, swm = DGB_STENCIL_SWM, average = DGB_STENCIL_AVERAGE, symv = DGB_STENCIL_SYMV };
inline void stencil(csr_filter f, const IVec& row_offsets, const IVec& cols, const DVec* vals, const DVec& x, DVec& y, double alpha = 0.) {
    check(dgb_csr_stencil((int)f, (int)row_offsets.size() - 1, row_offsets.data(), cols.data(), vals ? vals->data() : nullptr, alpha,
                          x.data(), y.data(), nullptr));
}
}  // namespace blas2

namespace tensor {
// dg::tensor::multiply3d (multiply.h:243): out_i = lambda T_ij in_j + mu out_i; t row major, null entries = identity
inline void multiply3d(double lambda, const std::array<const DVec*, 9>& t, const std::array<const DVec*, 3>& in, double mu,
                       const std::array<DVec*, 3>& out) {
    const double* tp[9];
    for (int k = 0; k < 9; k++) tp[k] = t[k] ? t[k]->data() : nullptr;
    const double* ip[3] = {in[0]->data(), in[1]->data(), in[2]->data()};
    double* op[3] = {out[0]->data(), out[1]->data(), out[2]->data()};
    check(dgb_tensor_multiply3d(in[0]->size(), nullptr, lambda, tp, ip, mu, op, nullptr));
}
}  // namespace tensor

namespace geo {
// the fields of a dg::geo::Fieldaligned the parallel-derivative formulas read (fieldaligned.h: deltaPhi, sqrtG*, bphi*)
struct FieldalignedFields {
    double delta_phi = 0;
    const DVec *sqrtGm = nullptr, *sqrtG = nullptr, *sqrtGp = nullptr, *bphiM = nullptr, *bphi = nullptr, *bphiP = nullptr;
};
namespace detail {
inline const double* p(const DVec* v) { return v ? v->data() : nullptr; }
inline void apply(int kind, const FieldalignedFields& fa, double alpha, const DVec& a, const DVec& b, const DVec* c, double beta, DVec& g) {
    if (kind < 6) check(dgb_ds_apply(kind, g.size(), alpha, a.data(), b.data(), p(c), p(fa.bphiM), p(fa.bphi), p(fa.bphiP), fa.delta_phi, beta, g.data(), nullptr));
    else check(dgb_ds_apply_vol(kind, g.size(), alpha, a.data(), b.data(), p(c), p(fa.sqrtGm), p(fa.sqrtG), p(fa.sqrtGp), p(fa.bphiM), p(fa.bphi),
                                p(fa.bphiP), fa.delta_phi, beta, g.data(), nullptr));
}
}  // namespace detail
// free functions of inc/geometries/ds.h:743-1016, same argument order
inline void ds_forward(const FieldalignedFields& fa, double alpha, const DVec& f, const DVec& fp, double beta, DVec& g) { detail::apply(0, fa, alpha, f, fp, nullptr, beta, g); }
inline void ds_backward(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& f, double beta, DVec& g) { detail::apply(1, fa, alpha, f, fm, nullptr, beta, g); }
inline void ds_centered(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& fp, double beta, DVec& g) { detail::apply(2, fa, alpha, fm, fp, nullptr, beta, g); }
inline void ds_forward2(const FieldalignedFields& fa, double alpha, const DVec& f, const DVec& fp, const DVec& fpp, double beta, DVec& g) { detail::apply(3, fa, alpha, f, fp, &fpp, beta, g); }
inline void ds_backward2(const FieldalignedFields& fa, double alpha, const DVec& fmm, const DVec& fm, const DVec& f, double beta, DVec& g) { detail::apply(4, fa, alpha, f, fm, &fmm, beta, g); }
inline void dss_centered(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& f, const DVec& fp, double beta, DVec& g) { detail::apply(5, fa, alpha, fm, f, &fp, beta, g); }
inline void dssd_centered(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& f, const DVec& fp, double beta, DVec& g) { detail::apply(6, fa, alpha, fm, f, &fp, beta, g); }
inline void ds_divBackward(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& f, double beta, DVec& g) { detail::apply(7, fa, alpha, fm, f, nullptr, beta, g); }
inline void ds_divForward(const FieldalignedFields& fa, double alpha, const DVec& f, const DVec& fp, double beta, DVec& g) { detail::apply(8, fa, alpha, f, fp, nullptr, beta, g); }
inline void ds_divCentered(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& fp, double beta, DVec& g) { detail::apply(9, fa, alpha, fm, fp, nullptr, beta, g); }
inline void ds_average(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& fp, double beta, DVec& g) { detail::apply(10, fa, alpha, fm, fp, nullptr, beta, g); }
inline void ds_slope(const FieldalignedFields& fa, double alpha, const DVec& fm, const DVec& fp, double beta, DVec& g) { ds_centered(fa, alpha, fm, fp, beta, g); }
}  // namespace geo

}  // namespace dgb200
