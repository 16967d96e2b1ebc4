// Internal: device-resident EllSparseBlockMat with its launch plan.
#pragma once
#include "common.cuh"
#include <vector>

namespace dgb {

constexpr int ELL_MAX_BPL = 4;
constexpr int ELL_MAX_N = 5;  // templated fast paths; larger n -> generic kernel

struct EllDev {
    int num_rows = 0, num_cols = 0, bpl = 0, n = 0, left = 1, right = 1, nblocks = 0, rr0 = 0, rr1 = 1;
    double* data = nullptr;  // device
    int* cols = nullptr;     // device
    int* didx = nullptr;     // device
    std::vector<double> h_data;
    std::vector<int> h_cols, h_didx;
    // "interior pattern": rows i in [i_lo, i_hi) satisfy cols[i][d] == i + off[d], data_idx[i][d] == did[d]
    bool has_pattern = false;
    int i_lo = 0, i_hi = 0;
    int off[ELL_MAX_BPL] = {0, 0, 0, 0};
    int did[ELL_MAX_BPL] = {0, 0, 0, 0};
    size_t total_rows() const { return (size_t)num_rows * n * left * right; }
    size_t total_cols() const { return (size_t)num_cols * n * left * right; }
};

int ell_upload(EllDev& m, const dgb_ell_host* h);
void ell_release(EllDev& m);
int ell_symv(const EllDev& m, double alpha, const double* x, double beta, double* y, cudaStream_t st, bool force_generic);

// plain-old-data view passed to kernels
struct EllArgs {
    int num_rows, num_cols, bpl, n, left, right, rr0, rr1;
    int i_lo, i_hi;
    int off[ELL_MAX_BPL];
    const double* data;
    const int* cols;
    const int* didx;
};
inline EllArgs ell_args(const EllDev& m) {
    EllArgs a;
    a.num_rows = m.num_rows; a.num_cols = m.num_cols; a.bpl = m.bpl; a.n = m.n; a.left = m.left; a.right = m.right;
    a.rr0 = m.rr0; a.rr1 = m.rr1;
    a.i_lo = m.has_pattern ? m.i_lo : 0;
    a.i_hi = m.has_pattern ? m.i_hi : 0;
    for (int d = 0; d < ELL_MAX_BPL; d++) a.off[d] = m.off[d];
    a.data = m.data; a.cols = m.cols; a.didx = m.didx;
    return a;
}

// the interior blocks in slot order, passed by value as a kernel parameter so that every coefficient is a
// constant-bank operand of the DFMA that uses it (no load instruction, no register)
template <int N, int BPL>
struct EllCoef {
    double c[BPL][N][N];
};
template <int N, int BPL>
inline EllCoef<N, BPL> ell_coef(const EllDev& m) {
    EllCoef<N, BPL> cf;
    for (int d = 0; d < BPL; d++)
        for (int k = 0; k < N; k++)
            for (int q = 0; q < N; q++)
                cf.c[d][k][q] = m.has_pattern ? m.h_data[((size_t)m.did[d] * N + k) * N + q] : 0.;
    return cf;
}

}  // namespace dgb
