// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// C-ABI wrapper around the UNMODIFIED reference application class toefl::Explicit (src/toefl/toefl.h) and the
// reference's explicit Runge-Kutta stepper (inc/dg/runge_kutta.h ERKStep), OpenMP backend, compiled from the sources
// where they lie by oracle/Makefile into oracle/_ref/libdgref_toefl.so.  Used to pin the parity of the toefl
// right-hand side / time step built on libdgb200.so (tests/test_gpu_toefl.py, tests/golden/make_golden_toefl.py) and
// as CPU baseline of tools/toefl_bench.py.  Nothing under feltor_b200/ may link or load this file.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <array>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dg/algorithm.h"
#include "dg/file/json_utilities.h"
#include "toefl.h"

using DVec = dg::DVec;
using Vec2 = std::array<DVec, 2>;
using Rhs = toefl::Explicit<dg::CartesianGrid2d, dg::DMatrix, DVec>;

struct RefToefl {
    toefl::Parameters p;
    dg::CartesianGrid2d grid;
    Rhs rhs;
    std::string flr;
    RefToefl(const toefl::Parameters& pp, const std::string& f)
        : p(pp), grid(0, pp.lx, 0., pp.ly, pp.n, pp.Nx, pp.Ny, pp.bcx, pp.bcy), rhs(grid, pp), flr(f) {}
};

// device backend: the timed regions end when the device has finished
static void device_sync() {
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    cudaDeviceSynchronize();
#endif
}
static void in2(const double* a, const double* b, size_t n, Vec2& y) {
    y[0].assign(a, a + n);
    y[1].assign(b, b + n);
}
static void out2(const Vec2& y, double* a, double* b) {
    thrust::copy(y[0].begin(), y[0].end(), a);
    thrust::copy(y[1].begin(), y[1].end(), b);
}

extern "C" {
void* ref_toefl_create(const char* json) {
    dg::file::WrappedJsonValue js(dg::file::error::is_throw);
    js = dg::file::string2Json(json);
    toefl::Parameters p(js);
    return new RefToefl(p, js["init"]["flr"].asString());
}
void ref_toefl_free(void* h) { delete (RefToefl*)h; }
int ref_toefl_size(void* h) { return (int)((RefToefl*)h)->grid.size(); }
// initial condition of src/toefl/toefl.cpp:50-72
void ref_toefl_init(void* hh, double* y0, double* y1) {
    RefToefl* h = (RefToefl*)hh;
    const toefl::Parameters& p = h->p;
    dg::Gaussian g(p.posX * p.lx, p.posY * p.ly, p.sigma, p.sigma, p.amp);
    Vec2 y({dg::evaluate(g, h->grid), dg::evaluate(g, h->grid)});
    if (p.model == "local" || p.model == "global") {
        dg::blas1::copy(y[0], y[1]);
        if (p.tau != 0 && h->flr == "gamma_inv") dg::apply(h->rhs.gamma_inv(), y[0], y[1]);
    }
    out2(y, y0, y1);
}
double ref_toefl_rhs(void* hh, double t, const double* y0, const double* y1, double* yp0, double* yp1) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    Vec2 y, yp;
    in2(y0, y1, n, y);
    yp = y;
    auto t0 = std::chrono::steady_clock::now();
    h->rhs(t, y, yp);
    device_sync();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out2(yp, yp0, yp1);
    return sec;
}
void ref_toefl_phi(void* hh, int i, double* out) {
    const DVec& v = ((RefToefl*)hh)->rhs.phi(i);
    thrust::copy(v.begin(), v.end(), out);
}
// nsteps fixed steps of size dt with dg::ERKStep (runge_kutta.h:163-400); returns the seconds spent
double ref_toefl_erk(void* hh, const char* tableau, double t0, double dt, int nsteps, double* y0, double* y1) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    Vec2 y, y1v, delta;
    in2(y0, y1, n, y);
    y1v = y; delta = y;
    dg::ERKStep<Vec2> erk(tableau, y);
    double t = t0, t1 = t0;
    auto c0 = std::chrono::steady_clock::now();
    for (int k = 0; k < nsteps; k++) {
        erk.step(h->rhs, t, y, t1, y1v, dt, delta);
        t = t1;
        y.swap(y1v);
    }
    device_sync();
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
    out2(y, y0, y1);
    return sec;
}
// nsteps calls of dg::Adaptive<dg::ERKStep>::step with dg::pid_control and dg::l2norm, u0 aliasing u1 the way
// dg::AdaptiveTimeloop::do_integrate drives it (adaptive.h:232-395,640-685; src/toefl/toefl.cpp:88-91).
// dt_io: initial / proposed next step, dts[k] = step proposed after call k, t_io: time; returns the number of failed steps
int ref_toefl_adaptive(void* hh, const char* tableau, double* t_io, double* dt_io, int nsteps, double rtol, double atol,
                       double* y0, double* y1, double* dts) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    Vec2 y;
    in2(y0, y1, n, y);
    dg::Adaptive<dg::ERKStep<Vec2>> adapt(tableau, y);
    double t = *t_io, dt = *dt_io;
    for (int k = 0; k < nsteps; k++) {
        adapt.step(h->rhs, t, y, t, y, dt, dg::pid_control, dg::l2norm, rtol, atol);
        if (dts) dts[k] = dt;
    }
    *t_io = t;
    *dt_io = dt;
    out2(y, y0, y1);
    return (int)adapt.nfailed();
}
// dg::ExplicitMultistep (multistep.h:59-100, FilteredExplicitMultistep :514-639): init at t0, then nsteps steps of dt (the
// first steps-1 are dg::ShuOsher steps, runge_kutta.h:883-910); ts[k] = time after step k
void ref_toefl_multistep(void* hh, const char* tableau, double t0, double dt, int nsteps, double* y0, double* y1, double* ts) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    Vec2 y;
    in2(y0, y1, n, y);
    dg::ExplicitMultistep<Vec2> ms(tableau, y);
    double t = t0;
    ms.init(h->rhs, t, y, dt);
    for (int k = 0; k < nsteps; k++) {
        ms.step(h->rhs, t, y);
        if (ts) ts[k] = t;
    }
    out2(y, y0, y1);
}
// device backend: switch the fused Elliptic2d / PCG hooks of the binding on or off (integration/dgb_shim/dg/backend/dgb_fused.h)
void ref_set_fusion(int on) {
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    dgb::shim::fusion_flag() = on ? 1 : 0;
#else
    (void)on;
#endif
}
// how many dispatches of this library went to libdgb200.so entry points / to generic kernel templates (device backend only)
void ref_dispatch_counters(long long* library, long long* generic) {
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    *library = dgb::shim::counters().library;
    *generic = dgb::shim::counters().generic;
#else
    *library = 0;
    *generic = 0;
#endif
}
int ref_toefl_backend_is_device() { return THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA ? 1 : 0; }
int ref_toefl_ncalls(void* hh) { return (int)((RefToefl*)hh)->rhs.ncalls(); }

// ---- the building blocks of toefl::Explicit one by one (the class keeps them private), constructed as toefl.h:60-83 does
void ref_toefl_helmholtz_solve(void* hh, double* x, const double* b, int* numbers) {
    RefToefl* h = (RefToefl*)hh;
    const toefl::Parameters& p = h->p;
    dg::MultigridCG2d<dg::CartesianGrid2d, dg::DMatrix, DVec> mg(h->grid, p.num_stages);
    std::vector<dg::Helmholtz<dg::CartesianGrid2d, dg::DMatrix, DVec>> ops;
    for (unsigned u = 0; u < p.num_stages; u++) ops.push_back({-0.5 * p.tau, {mg.grid(u), p.pol_dir}});
    const size_t n = h->grid.size();
    DVec xv(x, x + n), bv(b, b + n);
    std::vector<unsigned> num = mg.solve(ops, xv, bv, p.eps_gamma);
    for (unsigned u = 0; u < p.num_stages; u++) numbers[u] = (int)num[u];
    thrust::copy(xv.begin(), xv.end(), x);
}
void ref_toefl_pol_solve(void* hh, const double* chi, double* x, const double* b, int* numbers) {
    RefToefl* h = (RefToefl*)hh;
    const toefl::Parameters& p = h->p;
    dg::MultigridCG2d<dg::CartesianGrid2d, dg::DMatrix, DVec> mg(h->grid, p.num_stages);
    std::vector<dg::Elliptic<dg::CartesianGrid2d, dg::DMatrix, DVec>> ops;
    for (unsigned u = 0; u < p.num_stages; u++) ops.push_back({mg.grid(u), p.pol_dir, 1.});
    const size_t n = h->grid.size();
    DVec cv(chi, chi + n), xv(x, x + n), bv(b, b + n);
    std::vector<DVec> mc = mg.project(cv);
    for (unsigned u = 0; u < p.num_stages; u++) ops[u].set_chi(mc[u]);
    std::vector<unsigned> num = mg.solve(ops, xv, bv, p.eps_pol);
    for (unsigned u = 0; u < p.num_stages; u++) numbers[u] = (int)num[u];
    thrust::copy(xv.begin(), xv.end(), x);
}
void ref_toefl_upwind(void* hh, double alpha, const double* vx, const double* vy, const double* f, double beta, double* result) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    dg::Advection<dg::CartesianGrid2d, dg::DMatrix, DVec> adv(h->grid);
    DVec a(vx, vx + n), b(vy, vy + n), c(f, f + n), r(result, result + n);
    adv.upwind(alpha, a, b, c, beta, r);
    thrust::copy(r.begin(), r.end(), result);
}
void ref_toefl_arakawa(void* hh, double alpha, const double* lhs, const double* rhs, double beta, double* result) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    dg::ArakawaX<dg::CartesianGrid2d, dg::DMatrix, DVec> ar(h->grid);
    DVec a(lhs, lhs + n), b(rhs, rhs + n), r(result, result + n);
    ar(alpha, a, b, beta, r);
    thrust::copy(r.begin(), r.end(), result);
}
void ref_toefl_variation(void* hh, const double* phi, double* out) {
    RefToefl* h = (RefToefl*)hh;
    const size_t n = h->grid.size();
    dg::Elliptic<dg::CartesianGrid2d, dg::DMatrix, DVec> pol(h->grid, h->p.pol_dir, 1.);
    DVec a(phi, phi + n), r(n);
    pol.variation(a, r);
    thrust::copy(r.begin(), r.end(), out);
}
void ref_toefl_binv(void* hh, double* out) {
    RefToefl* h = (RefToefl*)hh;
    const toefl::Parameters& p = h->p;
    DVec b = dg::evaluate(dg::LinearX(p.kappa, 1. - p.kappa * p.posX * p.lx), h->grid);
    thrust::copy(b.begin(), b.end(), out);
}
}
