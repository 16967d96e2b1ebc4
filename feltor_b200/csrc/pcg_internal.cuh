// Internal: the Pcg workspace type is opaque outside pcg.cu; multigrid.cu drives it through these.
#pragma once
#include "elliptic.cuh"
namespace dgb {
struct Pcg;
Pcg* pcg_new(size_t n, int* err);
void pcg_delete(Pcg* s);
int pcg_solve(Pcg& s, Elliptic2dPlan& A, double* x, const double* b, const double* P, const double* W, double eps,
              double nrmb_correction, int test_frequency, int max_iter, int* iterations, cudaStream_t st);
}  // namespace dgb
