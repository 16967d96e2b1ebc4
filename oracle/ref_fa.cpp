// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// C-ABI wrapper around the UNMODIFIED reference classes dg::geo::Fieldaligned and dg::geo::DS (inc/geometries/fieldaligned.h,
// ds.h) on the circular test field of the reference's own benchmark inc/geometries/ds_b.cpp:70-84 (config 4 of BASELINE.json):
// CylindricalGrid3d(R0-a, R0+a, -a, a, 0, 2 pi, n, Nx, Ny, Nz, NEU, NEU, PER), createCircularField(R0 = 10, I0 = 20),
// Fieldaligned(bhat, g3d, NEU, NEU, NoLimiter(), 1e-8, mx, my, -1, method).  Compiled by oracle/Makefile into
// oracle/_ref/libdgref_fa.so with the OpenMP backend.  It hands out the REAL field-line interpolation matrices I+ / I- (CSR,
// built by the reference's field-line integration + interpolation + projection) and the fields the DS formulas need, runs
// ds.centered / ds.forward / ... as the oracle and times them as the CPU baseline of `bench.py --workload ds`.
// The matrices are private members of the class: `#define private public` around the one header is what a test harness
// can do without touching the reference sources.
#include <cstring>
#include <string>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dg/algorithm.h"
#include "geometries/magnetic_field.h"
#include "geometries/toroidal.h"
#include "geometries/testfunctors.h"
#define private public
#include "geometries/fieldaligned.h"
#undef private
#include "geometries/ds.h"

// The same file is compiled a second time by integration/Makefile with nvcc on the libdgb200 binding (-DREF_FA_DEVICE): the
// reference's Fieldaligned / DS templates instantiated on DEVICE containers (dg::IDMatrix, dg::DVec), i.e. what a Feltor
// application on a GPU runs; tests/test_gpu_shim.py compares the two builds call by call.
#ifdef REF_FA_DEVICE
using FaMatrix = dg::IDMatrix;
using FaVec = dg::DVec;
#else
using FaMatrix = dg::IHMatrix;
using FaVec = dg::HVec;
#endif
using FA = dg::geo::Fieldaligned<dg::aProductGeometry3d, FaMatrix, FaVec>;
using DSop = dg::geo::DS<dg::aProductGeometry3d, FaMatrix, FaVec>;
template <class V>
static void to_host(const V& v, double* out) { thrust::copy(v.begin(), v.end(), out); }

struct RefFA {
    dg::CylindricalGrid3d g3d;
    dg::geo::TokamakMagneticField mag;
    FA fa;
    DSop ds;
    RefFA(unsigned n, unsigned Nx, unsigned Ny, unsigned Nz, unsigned mx, unsigned my, const std::string& method)
        : g3d(10. - 1., 10. + 1., -1., 1., 0, 2. * M_PI, n, Nx, Ny, Nz, dg::NEU, dg::NEU, dg::PER),
          mag(dg::geo::createCircularField(10., 20.)),
          fa(dg::geo::createBHat(mag), g3d, dg::NEU, dg::NEU, dg::geo::NoLimiter(), 1e-8, mx, my, -1, method),
          ds(fa) {}
};

extern "C" {
void* ref_fa_create(int n, int Nx, int Ny, int Nz, int mx, int my, const char* method) {
    try { return new RefFA(n, Nx, Ny, Nz, mx, my, method); } catch (std::exception& e) { fprintf(stderr, "ref_fa_create: %s\n", e.what()); return nullptr; }
}
void ref_fa_free(void* h) { delete (RefFA*)h; }
int ref_fa_plane_size(void* h) { return (int)((RefFA*)h)->fa.m_plus.num_rows(); }
int ref_fa_size(void* h) { return (int)((RefFA*)h)->g3d.size(); }
double ref_fa_delta_phi(void* h) { return ((RefFA*)h)->fa.deltaPhi(); }
// which: 0 = I+ (m_plus), 1 = I- (m_minus)
int ref_fa_nnz(void* h, int which) {
    auto& m = which == 0 ? ((RefFA*)h)->fa.m_plus : ((RefFA*)h)->fa.m_minus;
    return (int)m.num_nnz();
}
void ref_fa_csr(void* h, int which, int* pos, int* idx, double* val) {
    auto& m = which == 0 ? ((RefFA*)h)->fa.m_plus : ((RefFA*)h)->fa.m_minus;
    thrust::copy(m.row_offsets().begin(), m.row_offsets().end(), pos);
    thrust::copy(m.column_indices().begin(), m.column_indices().end(), idx);
    thrust::copy(m.values().begin(), m.values().end(), val);
}
// 3-d fields of the Fieldaligned object: 0 bphi, 1 bphiM, 2 bphiP, 3 sqrtG, 4 sqrtGm, 5 sqrtGp, 6 hbm, 7 hbp
void ref_fa_field(void* h, int which, double* out) {
    FA& fa = ((RefFA*)h)->fa;
    const FaVec* v[8] = {&fa.bphi(), &fa.bphiM(), &fa.bphiP(), &fa.sqrtG(), &fa.sqrtGm(), &fa.sqrtGp(), &fa.hbm(), &fa.hbp()};
    to_host(*v[which], out);
}
// the test function of ds_b.cpp:86-87 pulled back to the grid
void ref_fa_testfunction(void* h, double* out) {
    RefFA* r = (RefFA*)h;
    dg::HVec f = dg::pullback(dg::geo::TestFunctionDirNeu(r->mag), r->g3d);
    std::copy(f.begin(), f.end(), out);
}
// Fieldaligned::operator()(einsPlus / einsMinus, f, fe): which 0 plus, 1 minus
void ref_fa_shift(void* h, int which, const double* f, double* fe) {
    RefFA* r = (RefFA*)h;
    const size_t n = r->g3d.size();
    FaVec in(f, f + n), out(n);
    r->fa(which == 0 ? dg::geo::einsPlus : dg::geo::einsMinus, in, out);
    to_host(out, fe);
}
// kind: 0 centered(alpha, f, beta, g)  1 forward  2 backward  3 dss  4 divCentered  5 symv (-DS^dagger DS, not used yet)
// returns the seconds of ONE application (mean over reps), result of the last one in g
double ref_fa_ds(void* h, int kind, double alpha, const double* f, double beta, double* g, int reps) {
    RefFA* r = (RefFA*)h;
    const size_t n = r->g3d.size();
    FaVec in(f, f + n), out(g, g + n), out0(out);
    double sec = 0.;
    for (int k = 0; k < reps; k++) {
        out = out0;
        auto t0 = std::chrono::steady_clock::now();
        if (kind == 0) r->ds.centered(alpha, in, beta, out);
        else if (kind == 1) r->ds.forward(alpha, in, beta, out);
        else if (kind == 2) r->ds.backward(alpha, in, beta, out);
        else if (kind == 3) r->ds.dss(alpha, in, beta, out);
        else r->ds.divCentered(alpha, in, beta, out);
        sec += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
#ifdef REF_FA_DEVICE
    cudaDeviceSynchronize();
#endif
    to_host(out, g);
    return sec / reps;
}
int ref_fa_is_device() {
#ifdef REF_FA_DEVICE
    return 1;
#else
    return 0;
#endif
}
int ref_fa_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
}
