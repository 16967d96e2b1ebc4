// C++ host program against include/dg_b200.hpp: the toefl right-hand side (src/toefl/toefl.h, "global" model with the
// default input src/toefl/input/default.json) driven by dg::ERKStep and dg::Adaptive exactly like src/toefl/toefl.cpp:79-91.
//   toefl_demo <N> <fixed steps> <adaptive steps> [initial state: file of 2*size doubles] [multistep steps]
// prints exact-dot checksums of the state; tests/test_cpp_host.py compares them with the fixtures of the unmodified reference.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "dg_b200.hpp"
using namespace dgb200;

struct Parameters {  // src/toefl/parameters.h with the values of input/default.json
    unsigned n = 3, Nx = 24, Ny = 24, stages = 3;
    double lx = 200, ly = 200, amp = 1.0, sigma = 10, posX = 0.3, posY = 0.5, kappa = 0.00015, tau = 1, nu = 1e-6;
    std::vector<double> eps_pol{1e-6, 1e-6, 1e-6}, eps_gamma{1e-7, 1e-7, 1e-7};
};

// toefl::Explicit<CartesianGrid2d, DMatrix, DVec>, model "global" without the Boussinesq approximation
class Explicit {
    Parameters p;
    Grid2d g;
    DVec chi, omega, uE2, binv, gamma_n;
    DVec2 phi, dxphi, dyphi, ype, lapy, v;
    Elliptic2d laplaceM;
    Advection adv;
    MultigridCG2d multigrid;
    Extrapolation old_phi, old_psi, old_gammaN;
    std::vector<Elliptic2d> multi_pol;
    std::vector<Helmholtz> multi_gamma1;
    DMatrix dx, dy;
    unsigned m_ncalls = 0;
    static DVec2 two(size_t n) { return DVec2{DVec(n, 0.), DVec(n, 0.)}; }
  public:
    explicit Explicit(const Parameters& par)
        : p(par), g(0, p.lx, 0, p.ly, p.n, p.Nx, p.Ny, DIR, PER), chi(g.size(), 0.), omega(chi), uE2(chi), gamma_n(chi), phi(two(g.size())),
          dxphi(phi), dyphi(phi), ype(phi), lapy(phi), v(phi), laplaceM(g, centered), adv(g), multigrid(g, p.stages), old_phi(2, chi),
          old_psi(2, chi), old_gammaN(2, chi), dx(create::dx(g, g.bcx(), centered)), dy(create::dy(g, g.bcy(), centered)) {
        // dg::LinearX(kappa, 1 - kappa*posX*lx) evaluated on the grid (toefl.h:63): a*x + b, one rounding (the reference's host
        // compiler contracts it)
        const double a = p.kappa, b = 1. - p.kappa * p.posX * p.lx;
        binv = DVec(evaluate([a, b](double x, double) { return std::fma(a, x, b); }, g));
        for (unsigned u = 0; u < p.stages; u++) {
            multi_pol.emplace_back(multigrid.grid(u), centered, 1.);
            multi_gamma1.emplace_back(-0.5 * p.tau, Elliptic2d(multigrid.grid(u), centered));
        }
    }
    const Grid2d& grid() const { return g; }
    unsigned ncalls() const { return m_ncalls; }
    const DVec& potential(int i) const { return phi[i]; }
    DVec2 initial_condition() {  // src/toefl/toefl.cpp:50-72, flr "gamma_inv"
        const double x0 = p.posX * p.lx, y0 = p.posY * p.ly, s = p.sigma, amp = p.amp;
        DVec gauss(evaluate([=](double x, double y) { return amp * exp(-((x - x0) * (x - x0) / 2. / s / s + (y - y0) * (y - y0) / 2. / s / s)); }, g));
        DVec2 y{gauss, gauss};
        multi_gamma1[0].symv(y[0], y[1]);
        return y;
    }
    void polarisation(double t, const DVec2& y) {  // toefl.h:197-240
        blas1::copy(y[1], chi);
        blas1::plus(chi, 1.);
        blas1::pointwiseDot(chi, binv, chi);
        blas1::pointwiseDot(chi, binv, chi);
        std::vector<DVec> multi_chi = multigrid.project(chi);
        for (unsigned u = 0; u < p.stages; u++) multi_pol[u].set_chi(multi_chi[u]);
        old_gammaN.extrapolate(t, gamma_n);
        multigrid.solve(multi_gamma1, gamma_n, y[1], p.eps_gamma);
        old_gammaN.update(t, gamma_n);
        blas1::axpby(-1., y[0], 1., gamma_n, omega);
        old_phi.extrapolate(t, phi[0]);
        multigrid.solve(multi_pol, phi[0], omega, p.eps_pol);
        old_phi.update(t, phi[0]);
    }
    void compute_psi(double t) {  // toefl.h:160-195
        old_psi.extrapolate(t, phi[1]);
        multigrid.solve(multi_gamma1, phi[1], phi[0], p.eps_gamma);
        old_psi.update(t, phi[1]);
        multi_pol[0].variation(phi[0], uE2);
        blas1::pointwiseDot(1., binv, binv, uE2, 0., uE2);
        blas1::axpby(-0.5, uE2, 1., phi[1]);
    }
    void operator()(double t, const DVec2& y, DVec2& yp) {  // toefl.h:242-308
        m_ncalls++;
        polarisation(t, y);
        compute_psi(t);
        const double tau[2] = {-1., p.tau};
        for (int u = 0; u < 2; u++) {
            blas1::copy(y[u], ype[u]);
            blas1::plus(ype[u], 1.);
        }
        for (int u = 0; u < 2; u++) {
            blas2::symv(dx, phi[u], dxphi[u]);
            blas2::symv(dy, phi[u], dyphi[u]);
            blas1::pointwiseDot(-1., binv, dyphi[u], 0., v[0]);
            blas1::pointwiseDot(+1., binv, dxphi[u], 0., v[1]);
            blas1::plus(v[1], -tau[u] * p.kappa);
            adv.upwind(-1., v[0], v[1], y[u], 0., yp[u]);
            blas1::pointwiseDot(p.kappa, ype[u], dyphi[u], 1., yp[u]);
        }
        for (int u = 0; u < 2; u++) {
            laplaceM.symv(-1., y[u], 0., lapy[u]);
            blas1::axpby(p.nu, lapy[u], 1., yp[u]);
        }
    }
};

// the initial state: from a file (the test hands over the reference's, whose host exp() arguments are contracted by its
// compiler) or computed here
static DVec2 initial(Explicit& rhs, const char* file) {
    if (!file) return rhs.initial_condition();
    const size_t n = rhs.grid().size();
    HVec a(n), b(n);
    FILE* f = fopen(file, "rb");
    if (!f || fread(a.data(), sizeof(double), n, f) != n || fread(b.data(), sizeof(double), n, f) != n) { fprintf(stderr, "cannot read %s\n", file); exit(2); }
    fclose(f);
    return DVec2{DVec(a), DVec(b)};
}

int main(int argc, char** argv) {
    Parameters p;
    p.Nx = p.Ny = argc > 1 ? atoi(argv[1]) : 24;
    const int fixed = argc > 2 ? atoi(argv[2]) : 3, adaptive = argc > 3 ? atoi(argv[3]) : 8;
    const char* file = argc > 4 ? argv[4] : nullptr;
    {
        Explicit rhs(p);
        DVec2 y0 = initial(rhs, file), y1(y0), delta(y0);
        printf("init checksum: %.17g %.17g\n", blas1::dot(y0[0], y0[0]), blas1::dot(y0[1], y0[1]));
        ERKStep erk("Bogacki-Shampine-4-2-3", y0);
        double t = 0., t1 = 0.;
        for (int k = 0; k < fixed; k++) {
            erk.step(rhs, t, y0, t1, y1, 0.5, delta);
            t = t1;
            std::swap(y0, y1);
        }
        printf("erk checksum: %.17g %.17g phi %.17g %.17g calls %u\n", blas1::dot(y0[0], y0[0]), blas1::dot(y0[1], y0[1]),
               blas1::dot(rhs.potential(0), rhs.potential(0)), blas1::dot(rhs.potential(1), rhs.potential(1)), rhs.ncalls());
    }
    {
        Explicit rhs(p);
        DVec2 y = initial(rhs, file);
        Adaptive adapt("Bogacki-Shampine-4-2-3", y);
        double t = 0., dt = 1e-6;
        printf("adaptive dts:");
        for (int k = 0; k < adaptive; k++) {
            adapt.step(rhs, t, y, t, y, dt, pid_control, l2norm, 1e-5, 1e-6);
            printf(" %.17g", dt);
        }
        printf("\nadaptive checksum: %.17g %.17g t %.17g failed %u\n", blas1::dot(y[0], y[0]), blas1::dot(y[1], y[1]), t, adapt.nfailed());
    }
    const int msteps = argc > 5 ? atoi(argv[5]) : 0;
    if (msteps > 0) {  // dg::ExplicitMultistep("TVB-3-3"), dt = 0.3 (BASELINE config 3 names a multistep stepper)
        Explicit rhs(p);
        DVec2 y = initial(rhs, file);
        ExplicitMultistep ms("TVB-3-3", y);
        double t = 0.;
        ms.init(rhs, t, y, 0.3);
        printf("multistep ts:");
        for (int k = 0; k < msteps; k++) {
            ms.step(rhs, t, y);
            printf(" %.17g", t);
        }
        printf("\nmultistep checksum: %.17g %.17g calls %u\n", blas1::dot(y[0], y[0]), blas1::dot(y[1], y[1]), rhs.ncalls());
    }
    return 0;
}
