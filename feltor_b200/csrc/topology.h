// Internal: host-side topology objects shared by topology.cu, elliptic.cu and multigrid.cu.
#pragma once
#include <vector>
#include <cstddef>
#include "../../include/dgb200.h"

namespace dgb {

// small dense n x n matrix (row-major)
struct Mat {
    int n;
    std::vector<double> a;
    explicit Mat(int n_ = 0) : n(n_), a((size_t)n_ * n_, 0.) {}
    double& operator()(int i, int j) { return a[(size_t)i * n + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * n + j]; }
};

// host EllSparseBlockMat owning its arrays (fields as inc/dg/backend/sparseblockmat.h:168-177)
struct EllHost {
    int num_rows = 0, num_cols = 0, bpl = 0, n = 0, left = 1, right = 1, nblocks = 0, rr0 = 0, rr1 = 1;
    std::vector<double> data;
    std::vector<int> cols_idx, data_idx;
};

std::vector<double> dlt_abscissas(int n);
std::vector<double> dlt_weights(int n);
Mat dlt_backward(int n);
Mat dlt_forward(int n);

double grid_h(const dgb_grid* g, int u);
size_t grid_shape(const dgb_grid* g, int u);
size_t grid_size(const dgb_grid* g);
std::vector<double> grid_abscissas(const dgb_grid* g, int u);
std::vector<double> grid_weights1d(const dgb_grid* g, int u);
std::vector<double> grid_weights(const dgb_grid* g);
void update_left_right(EllHost& m, const dgb_grid* g, int coord);

int topo_dx(EllHost& out, int n, int N, double h, int bc, int dir);
int topo_jump(EllHost& out, int n, int N, double h, int bc);
int topo_fast_interpolation1d(EllHost& out, int n, int N, int multiplyn, int multiplyN);
int topo_fast_projection1d(EllHost& out, int nold, int N, int dividen, int divideN);
void interpolation_row_xspace(double X, double x0, double x1, int n, int N, std::vector<int>& cols, std::vector<double>& vals);
void ell_view(const EllHost& m, dgb_ell_host* v);

}  // namespace dgb
