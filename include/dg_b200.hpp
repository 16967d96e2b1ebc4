// dg_b200.hpp -- header-only C++17 host layer over the C ABI of libdgb200.so.
// Mirrors the names, argument order and error behaviour of the reference's public interface for the hot path
// (dg::blas1 / dg::blas2 / dg::Elliptic2d / dg::PCG / dg::MultigridCG2d, inc/dg/blas1.h, blas2.h, elliptic.h, pcg.h,
// multigrid.h) on an RAII device vector.  No arithmetic lives here: every call forwards to a dgb_* entry point.
// A Feltor tree would instead plug the C ABI into its own dispatch seams; this header is the
// stand-alone way to program against the library from C++.
#pragma once
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
    dgb_elliptic2d* plan() { return m_plan; }
};

// dg::PCG<DVec> (inc/dg/pcg.h:25-199)
class PCG {
    dgb_pcg* m_pcg = nullptr;
    unsigned m_max = 0;
    bool m_throw = true;
  public:
    PCG() = default;
    PCG(const DVec& copyable, unsigned max_iterations) : m_max(max_iterations) { check(dgb_pcg_create(&m_pcg, copyable.size())); }
    PCG(const PCG&) = delete;
    PCG(PCG&& o) noexcept { std::swap(m_pcg, o.m_pcg); m_max = o.m_max; m_throw = o.m_throw; }
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
        std::vector<int> num(m_stages);
        check(dgb_multigrid2d_solve(m_mg, A.data(), P.data(), W.data(), x.data(), b.data(), eps.data(), num.data(), nullptr));
        return std::vector<unsigned>(num.begin(), num.end());
    }
    std::vector<unsigned> solve(std::vector<Elliptic2d>& ops, DVec& x, const DVec& b, double eps) {
        return solve(ops, x, b, std::vector<double>(m_stages, eps));
    }
};

}  // namespace dgb200
