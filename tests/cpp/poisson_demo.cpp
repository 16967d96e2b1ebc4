// C++ host program against include/dg_b200.hpp: the elliptic2d_b.cpp problem (inc/dg/elliptic2d_b.cpp:22-38,141-149).
//   poisson_demo --host-only        : host-side part only (no device needed; used by the CPU test)
//   poisson_demo <Nx> <Ny> <stages> : nested-iteration multigrid solve and a plain PCG solve on the device
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "dg_b200.hpp"
using namespace dgb200;
const double amp = 0.9;
double pol(double x, double y) { return 1. + amp * sin(x) * sin(y); }
double rhs(double x, double y) {
    return 2. * sin(x) * sin(y) * (amp * sin(x) * sin(y) + 1) - amp * sin(x) * sin(x) * cos(y) * cos(y) - amp * cos(x) * cos(x) * sin(y) * sin(y);
}
double sol(double x, double y) { return sin(x) * sin(y); }
int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "--host-only")) {
        Grid2d g(0, M_PI, 0, 2 * M_PI, 3, 8, 6, DIR, PER);
        HVec w = create::weights(g), b = evaluate(rhs, g);
        double s = 0;
        for (size_t i = 0; i < w.size(); i++) s += w[i];
        printf("host-only size %zu sum(weights) %.15g version %d\n", g.size(), s, dgb_version());
        try { Grid2d bad(0, 1, 0, 1, 25, 4, 4); (void)create::weights(bad); return 1; } catch (Error& e) { printf("caught: %s\n", e.what()); }
        return fabs(s - 2 * M_PI * M_PI) < 1e-12 ? 0 : 1;
    }
    unsigned Nx = argc > 1 ? atoi(argv[1]) : 64, Ny = argc > 2 ? atoi(argv[2]) : 64, stages = argc > 3 ? atoi(argv[3]) : 3;
    Grid2d grid(0, M_PI, 0, 2 * M_PI, 3, Nx, Ny, DIR, PER);
    DVec w2d(create::weights(grid)), b(evaluate(rhs, grid)), chi(evaluate(pol, grid)), solution(evaluate(sol, grid));
    MultigridCG2d multigrid(grid, stages);
    std::vector<DVec> multi_chi = multigrid.project(chi);
    std::vector<Elliptic2d> multi_pol;
    for (unsigned u = 0; u < stages; u++) {
        multi_pol.emplace_back(multigrid.grid(u), forward, 1.0);
        multi_pol[u].set_chi(multi_chi[u]);
    }
    DVec x(grid.size(), 0.);
    std::vector<unsigned> num = multigrid.solve(multi_pol, x, b, 1e-6);
    printf("multigrid iterations:");
    for (unsigned u = 0; u < stages; u++) printf(" %u", num[u]);
    blas1::axpby(1., solution, -1., x);
    double err = sqrt(blas2::dot(w2d, x) / blas2::dot(w2d, solution));
    printf("\nrelative error %.3e\n", err);
    PCG pcg(x, 100000);
    DVec y(grid.size(), 0.);
    unsigned it = pcg.solve(multi_pol[0], y, b, multi_pol[0].precond(), multi_pol[0].weights(), 1e-6);
    printf("pcg iterations: %u\n", it);
    // the generic solve (any operator with symv / any callable): same iteration, unfused -- must agree bit for bit
    DVec z(grid.size(), 0.);
    Elliptic2d& A = multi_pol[0];
    unsigned itg = pcg.solve([&A](const DVec& in, DVec& out) { A.symv(in, out); }, z, b, A.precond(), A.weights(), 1e-6);
    blas1::axpby(1., y, -1., z);
    printf("generic pcg iterations: %u difference %.17g checksum %.17g\n", itg, blas2::dot(A.weights(), z), blas2::dot(A.weights(), y));
    return err < 1e-4 ? 0 : 1;
}
