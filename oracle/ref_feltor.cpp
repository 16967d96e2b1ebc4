// TEST INFRASTRUCTURE (oracle): C-ABI wrapper around the UNMODIFIED 3-d application class feltor::Explicit
// (/root/reference/src/feltor/feltor.h) so that one right-hand side can be driven without the
// application's main (which needs NetCDF).  Compiled twice from the sources where they lie:
//   oracle/Makefile      -> oracle/_ref/libdgref_feltor.so     reference OpenMP backend (the comparator)
//   integration/Makefile -> integration/_build/libdgshim_feltor.so   the same file with -DREF_FELTOR_DEVICE on the libdgb200 binding:
//                           dg::DVec / dg::DMatrix / dg::IDMatrix, i.e. what a Feltor user builds with the device backend
// The input file is this repository's own (tests/golden/feltor_input.json: circular field, small grid); the initial state is a
// smooth density blob on a background with a sheared parallel velocity, built here from analytic functions (the application's
// init.h pulls in NetCDF).
#include <cstdio>
#include <cmath>
#include <string>
#include <array>
#include <vector>
#include "dg/algorithm.h"
#include "dg/geometries/geometries.h"
#include "dg/file/json_utilities.h"
#include "feltor.h"

#ifdef REF_FELTOR_DEVICE
using FVec = dg::DVec;
using FIMat = dg::IDMatrix;
using FMat = dg::DMatrix;
#else
using FVec = dg::HVec;
using FIMat = dg::IHMatrix;
using FMat = dg::HMatrix;
#endif
using State = std::array<std::array<FVec, 2>, 2>;

namespace {
struct Blob {
    double R0, a;
    double operator()(double R, double Z, double phi) const {
        const double r2 = ((R - R0 - 0.3 * a) * (R - R0 - 0.3 * a) + (Z - 0.1 * a) * (Z - 0.1 * a)) / (0.25 * a * a);
        return 1. + 0.5 * exp(-r2) * (1. + 0.3 * cos(phi));
    }
};
struct Flow {
    double R0, a;
    double operator()(double R, double Z, double phi) const { return 0.2 * sin(M_PI * (R - R0) / a) * cos(M_PI * Z / a) * (1. + 0.2 * sin(phi)); }
};
dg::CylindricalGrid3d make_grid(const dg::file::WrappedJsonValue& js, const feltor::Parameters& p) {
    auto box = common::box(js);
    return dg::CylindricalGrid3d(box.at("Rmin"), box.at("Rmax"), box.at("Zmin"), box.at("Zmax"), 0, 2. * M_PI, p.n, p.Nx, p.Ny,
                                 p.symmetric ? 1 : p.Nz, p.bcxN, p.bcyN, dg::PER);
}
dg::geo::TokamakMagneticField make_mag(const dg::file::WrappedJsonValue& js) {
    dg::geo::TokamakMagneticField mag = dg::geo::createMagneticField(js["magnetic_field"]["params"]);
    if (js["FCI"].get("periodify", true).asBool()) {
        auto box = common::box(js);
        mag = dg::geo::periodify(mag, box.at("Rmin"), box.at("Rmax"), box.at("Zmin"), box.at("Zmax"), dg::NEU, dg::NEU);
    }
    return mag;
}
struct RefFeltor {
    dg::file::WrappedJsonValue js;
    feltor::Parameters p;
    dg::geo::TokamakMagneticField mag;
    dg::CylindricalGrid3d grid;
    feltor::Explicit<dg::CylindricalGrid3d, FIMat, FMat, FVec> rhs;
    State y, yp;
    explicit RefFeltor(const dg::file::WrappedJsonValue& j)
        : js(j), p(js), mag(make_mag(js)), grid(make_grid(js, p)), rhs(grid, p, mag, js) {
        const double R0 = mag.R0(), a = mag.params().a();
        dg::HVec n = dg::evaluate(Blob{R0, a}, grid), u = dg::evaluate(Flow{R0, a}, grid);
        for (int s = 0; s < 2; s++) {
            dg::assign(n, y[0][s]);
            dg::assign(u, y[1][s]);
        }
        dg::blas1::scal(y[1][1], 0.1);  // ions slower than electrons
        yp = y;
    }
};
void to_host(const FVec& v, double* out) {
    dg::HVec h;
    dg::assign(v, h);
    for (size_t i = 0; i < h.size(); i++) out[i] = h[i];
}
}  // namespace

extern "C" {
void* ref_feltor_create(const char* json_text) {
    try {
        dg::file::WrappedJsonValue js(dg::file::string2Json(json_text, dg::file::comments::are_discarded, dg::file::error::is_throw));
        return new RefFeltor(js);
    } catch (std::exception& e) {
        fprintf(stderr, "ref_feltor_create: %s\n", e.what());
        return nullptr;
    }
}
void ref_feltor_free(void* h) { delete (RefFeltor*)h; }
int ref_feltor_size(void* h) { return (int)((RefFeltor*)h)->grid.size(); }
int ref_feltor_is_device() {
#ifdef REF_FELTOR_DEVICE
    return 1;
#else
    return 0;
#endif
}
// state component (field 0 density, 1 velocity; species 0 electrons, 1 ions)
void ref_feltor_state(void* h, int field, int species, double* out) { to_host(((RefFeltor*)h)->y[field][species], out); }
// one evaluation of the right-hand side at time t; returns 0 or 1 on an exception
int ref_feltor_rhs(void* hh, double t, double* dn_e, double* dn_i, double* du_e, double* du_i) {
    RefFeltor* h = (RefFeltor*)hh;
    try {
        h->rhs(t, h->y, h->yp);
    } catch (std::exception& e) {
        fprintf(stderr, "ref_feltor_rhs: %s\n", e.what());
        return 1;
    }
    to_host(h->yp[0][0], dn_e); to_host(h->yp[0][1], dn_i); to_host(h->yp[1][0], du_e); to_host(h->yp[1][1], du_i);
    return 0;
}
// y += dt * yp (explicit Euler with the last right-hand side): lets a test take a few steps so that later evaluations see a
// non-trivial potential history
void ref_feltor_euler(void* hh, double dt) {
    RefFeltor* h = (RefFeltor*)hh;
    for (int f = 0; f < 2; f++)
        for (int s = 0; s < 2; s++) dg::blas1::axpby(dt, h->yp[f][s], 1., h->y[f][s]);
}
// the potentials the last evaluation solved for: 0 phi, 1 gamma phi
void ref_feltor_potential(void* h, int i, double* out) { to_host(((RefFeltor*)h)->rhs.potential(i), out); }
}
