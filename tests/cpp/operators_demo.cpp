// C++ host program against include/dg_b200.hpp: the operators toefl::Explicit is built from (Helmholtz multigrid solve,
// Advection::upwind, ArakawaX, Elliptic::variation, Extrapolation, blas1::reduce) on the toefl grid; prints exact-dot
// checksums that tests/test_cpp_host.py compares with the Python harness (both call the same C ABI).
//   operators_demo <N>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "dg_b200.hpp"
using namespace dgb200;
int main(int argc, char** argv) {
    unsigned N = argc > 1 ? atoi(argv[1]) : 48;
    Grid2d grid(0, 200, 0, 200, 3, N, N, DIR, PER);
    DVec w(create::weights(grid));
    DVec f(evaluate([](double x, double y) { return sin(0.05 * x) * cos(0.03 * y); }, grid));
    DVec vx(evaluate([](double x, double y) { return cos(0.02 * x) - 0.3; }, grid)), vy(evaluate([](double x, double y) { return sin(0.04 * y) + 0.1; }, grid));
    DVec b(evaluate([](double x, double y) { return exp(-((x - 60) * (x - 60) + (y - 100) * (y - 100)) / 200.); }, grid));
    // Helmholtz multigrid solve (toefl.h:75-83,155: Gamma^{-1})
    MultigridCG2d mg(grid, 3);
    std::vector<Helmholtz> gamma;
    for (unsigned u = 0; u < 3; u++) gamma.emplace_back(-0.5, Elliptic2d(mg.grid(u), centered));
    DVec x(grid.size(), 0.);
    std::vector<unsigned> num = mg.solve(gamma, x, b, std::vector<double>{1e-7, 1e-7, 1e-7});
    printf("helmholtz iterations: %u %u %u\n", num[0], num[1], num[2]);
    printf("helmholtz checksum: %.17g\n", blas2::dot(x, w, x));
    // Advection::upwind, ArakawaX, variation
    DVec r(grid.size(), 0.25);
    Advection adv(grid);
    adv.upwind(-1., vx, vy, f, 0.5, r);
    printf("upwind checksum: %.17g\n", blas2::dot(r, w, r));
    ArakawaX arakawa(grid);
    arakawa(0.7, f, b, -0.4, r);
    printf("arakawa checksum: %.17g\n", blas2::dot(r, w, r));
    Elliptic2d pol(grid, centered, 1.);
    DVec s(grid.size(), 0.);
    pol.variation(f, s);
    printf("variation checksum: %.17g\n", blas2::dot(s, w, s));
    // Extrapolation (linear) and reduce
    Extrapolation ex(2, f);
    ex.update(0., f);
    ex.update(0.5, b);
    ex.extrapolate(1.0, s);
    printf("extrapolation checksum: %.17g\n", blas2::dot(s, w, s));
    printf("max: %.17g min: %.17g\n", blas1::reduce(s, -1e300, blas1::reduce_op::max), blas1::reduce(s, 1e300, blas1::reduce_op::min));
    // Elliptic3d in compute-in-2d mode on a cylindrical grid (R, Z, phi)
    Grid3d g3(3., 5., -1., 1., 0., 2 * M_PI, 3, 12, 10, 5, DIR, NEU, PER, true);
    Elliptic3d pol3(g3, centered, 0.7);
    HVec hx(g3.size()), hchi(g3.size());
    for (size_t i = 0; i < hx.size(); i++) { hx[i] = sin(0.37 * (double)i); hchi[i] = 1.25 + 0.5 * cos(0.11 * (double)i); }
    DVec x3(hx), y3(g3.size(), 0.25), chi3(hchi);
    pol3.set_chi(chi3);
    pol3.symv(-0.5, x3, 0.3, y3);
    printf("elliptic3d checksum: %.17g\n", blas2::dot(y3, pol3.weights(), y3));
    printf("elliptic3dprecond checksum: %.17g\n", blas2::dot(pol3.precond(), pol3.weights(), pol3.precond()));
    // blas2::stencil (median over the window i-1, i, i+1), tensor::multiply3d, two ds.h formulas
    {
        const int n = (int)grid.size();
        std::vector<int> pos(n + 1), idx;
        for (int i = 0; i < n; i++) {
            pos[i] = (int)idx.size();
            for (int k = std::max(i - 1, 0); k <= std::min(i + 1, n - 1); k++) idx.push_back(k);
        }
        pos[n] = (int)idx.size();
        IVec dpos(pos), didx(idx);
        DVec med(grid.size(), 0.);
        blas2::stencil(blas2::csr_filter::median, dpos, didx, nullptr, f, med);
        printf("median checksum: %.17g\n", blas2::dot(med, w, med));
        DVec o0(f), o1(b), o2(vx);
        tensor::multiply3d(-1.5, {&vx, nullptr, &vy, nullptr, nullptr, &b, &f, nullptr, nullptr}, {&f, &b, &vy}, 0.25, {&o0, &o1, &o2});
        printf("tensor3d checksum: %.17g\n", blas2::dot(o0, w, o1) + blas2::dot(o2, w, o2));
        geo::FieldalignedFields fa;
        DVec G0(grid.size(), 1.5), Gm(grid.size(), 1.25), Gp(grid.size(), 1.75), bphi(grid.size(), 0.5);
        fa.delta_phi = 0.1; fa.sqrtGm = &Gm; fa.sqrtG = &G0; fa.sqrtGp = &Gp; fa.bphiM = &bphi; fa.bphi = &bphi; fa.bphiP = &bphi;
        DVec g1(grid.size(), 0.5), g2(grid.size(), 0.5);
        geo::ds_divCentered(fa, 0.7, f, b, -0.3, g1);
        geo::dss_centered(fa, 0.7, f, b, vx, -0.3, g2);
        printf("dsdiv checksum: %.17g\ndss checksum: %.17g\n", blas2::dot(g1, w, g1), blas2::dot(g2, w, g2));
    }
    return 0;
}
