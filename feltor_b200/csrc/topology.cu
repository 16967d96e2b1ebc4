// Host-side topology setup (no device code): Gauss-Legendre grids, weights, and the block-ELL matrices of the
// discontinuous Galerkin derivative / jump / fast projection / fast interpolation stencils.
// Our own implementation of what the reference builds in inc/dg/topology/{dlt,grid,weights,operator,dx,
// derivatives,interpolation,projection,fast_interpolation}.h.  The ARITHMETIC (order of the floating point
// operations, fused multiply-adds in the small dense products) follows the reference so that the coefficients
// come out bit-identical to what its OpenMP build produces (checked in tests/test_topology.py).
#include <cstdint>
#include "common.cuh"
#include "topology.h"
#include <cmath>
#include <cstring>
#include <algorithm>

namespace dgb {
#include "dlt_tables.inc"

// ------------------------------------------------------------------------------------------------ DLT tables
static inline size_t off_vec(int n) { return (size_t)n * (n - 1) / 2; }
static inline size_t off_mat(int n) { return (size_t)(n - 1) * n * (2 * n - 1) / 6; }
std::vector<double> dlt_abscissas(int n) { return {DLT_ABSCISSAS + off_vec(n), DLT_ABSCISSAS + off_vec(n) + n}; }
std::vector<double> dlt_weights(int n) { return {DLT_WEIGHTS + off_vec(n), DLT_WEIGHTS + off_vec(n) + n}; }
Mat dlt_backward(int n) { Mat m(n); std::copy(DLT_BACKWARD + off_mat(n), DLT_BACKWARD + off_mat(n) + n * n, m.a.begin()); return m; }
Mat dlt_forward(int n) { Mat m(n); std::copy(DLT_FORWARD + off_mat(n), DLT_FORWARD + off_mat(n) + n * n, m.a.begin()); return m; }

// ------------------------------------------------------------------------------------------------ small dense algebra
// (operator.h:228-350).  The product accumulates temp += l*r over k ascending with separately rounded multiply and add:
// that is what the reference's OpenMP build does (verified bit for bit in tests/test_topology.py).
Mat operator*(const Mat& l, const Mat& r) {
    Mat t(l.n);
    for (int i = 0; i < l.n; i++)
        for (int j = 0; j < l.n; j++) {
            double s = 0.;
            for (int k = 0; k < l.n; k++) s += l(i, k) * r(k, j);
            t(i, j) = s;
        }
    return t;
}
Mat operator*(double v, const Mat& m) { Mat t(m); for (auto& x : t.a) x *= v; return t; }
Mat operator+(const Mat& l, const Mat& r) { Mat t(l); for (size_t i = 0; i < t.a.size(); i++) t.a[i] += r.a[i]; return t; }
Mat operator-(const Mat& l, const Mat& r) { Mat t(l); for (size_t i = 0; i < t.a.size(); i++) t.a[i] -= r.a[i]; return t; }
Mat operator-(const Mat& m) { Mat t(m); for (auto& x : t.a) x = -x; return t; }
Mat transpose(const Mat& m) { Mat t(m.n); for (int i = 0; i < m.n; i++) for (int j = 0; j < m.n; j++) t(i, j) = m(j, i); return t; }

// Legendre-space building blocks (operator.h:581-647)
static Mat pipj_inv(int n) { Mat m(n); for (int i = 0; i < n; i++) m(i, i) = (double)(2 * i + 1) / 2.; return m; }
static Mat pidxpj(int n) { Mat m(n); for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if ((i + j) % 2) m(i, j) = 2; return m; }
static Mat rirj(int n) { Mat m(n); for (auto& x : m.a) x = 1.; return m; }
static Mat rilj(int n) { Mat m(n); for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) m(i, j) = (j % 2 == 0) ? 1. : -1.; return m; }
static Mat lirj(int n) { return transpose(rilj(n)); }
static Mat lilj(int n) { Mat m(n); for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) m(i, j) = ((i + j) % 2 == 0) ? 1. : -1.; return m; }

static void set_block(EllHost& A, int b, const Mat& m) {
    for (int i = 0; i < m.n; i++)
        for (int j = 0; j < m.n; j++) A.data[((size_t)b * m.n + i) * m.n + j] = m(i, j);
}
static EllHost make_ell(int rows, int cols, int bpl, int nblocks, int n) {
    EllHost A;
    A.num_rows = rows; A.num_cols = cols; A.bpl = bpl; A.n = n; A.nblocks = nblocks;
    A.data.assign((size_t)nblocks * n * n, 0.);
    A.cols_idx.assign((size_t)rows * bpl, 0);
    A.data_idx.assign((size_t)rows * bpl, 0);
    return A;
}
static inline void slot(EllHost& A, int i, int d, int block, int col) {
    A.data_idx[(size_t)i * A.bpl + d] = block;
    A.cols_idx[(size_t)i * A.bpl + d] = col;
}
static bool dir_left(int bc) { return bc == DGB_DIR || bc == DGB_DIR_NEU; }
static bool neu_left(int bc) { return bc == DGB_NEU || bc == DGB_NEU_DIR; }
static bool dir_right(int bc) { return bc == DGB_DIR || bc == DGB_NEU_DIR; }
static bool neu_right(int bc) { return bc == DGB_NEU || bc == DGB_DIR_NEU; }

// three-slot stencils (centered derivative, jump): blocks {bp, a, b, a_left, a_right} (dx.h:71-119)
static EllHost assemble3(int n, int N, int bc, const Mat& bp, const Mat& a, const Mat& b, const Mat& al, const Mat& ar) {
    if (bc != DGB_PER) {
        EllHost A = make_ell(N, N, 3, 5, n);
        set_block(A, 0, bp); set_block(A, 1, a); set_block(A, 2, b); set_block(A, 3, al); set_block(A, 4, ar);
        slot(A, 0, 0, 3, 0); slot(A, 0, 1, 2, 1); slot(A, 0, 2, 2, -1);
        for (int i = 1; i < N - 1; i++)
            for (int d = 0; d < 3; d++) slot(A, i, d, d, i + d - 1);
        slot(A, N - 1, 0, 0, N - 2); slot(A, N - 1, 1, 4, N - 1); slot(A, N - 1, 2, 4, -1);
        return A;
    }
    EllHost A = make_ell(N, N, 3, 3, n);
    set_block(A, 0, bp); set_block(A, 1, a); set_block(A, 2, b);
    for (int i = 0; i < N; i++)
        for (int d = 0; d < 3; d++) slot(A, i, d, d, (i + d - 1 + N) % N);
    return A;
}

// dx.h:33-120
static EllHost dx_symm(int n, int N, double h, int bc) {
    Mat l = lilj(n), r = rirj(n), lr = lirj(n), rl = rilj(n), d = pidxpj(n), t = pipj_inv(n);
    t = (2. / h) * t;
    Mat a = (1. / 2.) * t * (d - transpose(d));
    Mat al(a), ar(a);
    if (dir_left(bc)) al = al + 0.5 * t * l;
    else if (neu_left(bc)) al = al - 0.5 * t * l;
    if (dir_right(bc)) ar = ar - 0.5 * t * r;
    else if (neu_right(bc)) ar = ar + 0.5 * t * r;
    if (bc == DGB_PER) al = ar = a;
    Mat b = t * ((1. / 2.) * rl);
    Mat bp = t * ((-1. / 2.) * lr);
    Mat bw = dlt_backward(n), fw = dlt_forward(n);
    a = bw * a * fw; al = bw * al * fw; b = bw * b * fw; ar = bw * ar * fw; bp = bw * bp * fw;
    return assemble3(n, N, bc, bp, a, b, al, ar);
}
// dx.h:133-204
static EllHost dx_plus(int n, int N, double h, int bc) {
    Mat l = lilj(n), rl = rilj(n), d = pidxpj(n), t = pipj_inv(n);
    t = (2. / h) * t;
    Mat a = t * (-l - transpose(d));
    Mat al(a), ar(a);
    if (bc == DGB_DIR || bc == DGB_DIR_NEU) al = t * (-transpose(d));
    if (bc == DGB_NEU || bc == DGB_DIR_NEU) ar = t * d;
    Mat b = t * rl;
    Mat bw = dlt_backward(n), fw = dlt_forward(n);
    a = bw * a * fw; al = bw * al * fw; b = bw * b * fw; ar = bw * ar * fw;
    if (bc != DGB_PER) {
        EllHost A = make_ell(N, N, 2, 4, n);
        set_block(A, 0, a); set_block(A, 1, b); set_block(A, 2, al); set_block(A, 3, ar);
        slot(A, 0, 0, 2, 0); slot(A, 0, 1, 1, 1);
        for (int i = 1; i < N - 1; i++)
            for (int dd = 0; dd < 2; dd++) slot(A, i, dd, dd, i + dd);
        slot(A, N - 1, 0, 3, N - 1); slot(A, N - 1, 1, 3, -1);
        return A;
    }
    EllHost A = make_ell(N, N, 2, 2, n);
    set_block(A, 0, a); set_block(A, 1, b);
    for (int i = 0; i < N; i++)
        for (int dd = 0; dd < 2; dd++) slot(A, i, dd, dd, (i + dd + N) % N);
    return A;
}
// dx.h:217-288
static EllHost dx_minus(int n, int N, double h, int bc) {
    Mat l = lilj(n), lr = lirj(n), d = pidxpj(n), t = pipj_inv(n);
    t = (2. / h) * t;
    Mat a = t * (l + d);
    Mat ar(a), al(a);
    if (bc == DGB_DIR || bc == DGB_NEU_DIR) ar = t * (-transpose(d));
    if (bc == DGB_NEU || bc == DGB_NEU_DIR) al = t * d;
    Mat bp = (-t) * lr;
    Mat bw = dlt_backward(n), fw = dlt_forward(n);
    a = bw * a * fw; al = bw * al * fw; bp = bw * bp * fw; ar = bw * ar * fw;
    if (bc != DGB_PER) {
        EllHost A = make_ell(N, N, 2, 4, n);
        set_block(A, 0, bp); set_block(A, 1, a); set_block(A, 2, al); set_block(A, 3, ar);
        slot(A, 0, 0, 2, 0); slot(A, 0, 1, 2, -1);
        for (int i = 1; i < N - 1; i++)
            for (int dd = 0; dd < 2; dd++) slot(A, i, dd, dd, i + dd - 1);
        slot(A, N - 1, 0, 0, N - 2); slot(A, N - 1, 1, 3, N - 1);
        return A;
    }
    EllHost A = make_ell(N, N, 2, 2, n);
    set_block(A, 0, bp); set_block(A, 1, a);
    for (int i = 0; i < N; i++)
        for (int dd = 0; dd < 2; dd++) slot(A, i, dd, dd, (i + dd - 1 + N) % N);
    return A;
}
// dx.h:301-377
static EllHost jump1d(int n, int N, double h, int bc) {
    Mat l = lilj(n), r = rirj(n), lr = lirj(n), rl = rilj(n);
    Mat a = l + r;
    Mat al(a), ar(a);
    if (neu_left(bc)) al = r;
    if (neu_right(bc)) ar = l;
    Mat b = -rl, bp = -lr;
    Mat t = pipj_inv(n);
    t = (2. / h) * t;
    Mat bw = dlt_backward(n), fw = dlt_forward(n);
    a = bw * t * a * fw; al = bw * t * al * fw; b = bw * t * b * fw; ar = bw * t * ar * fw; bp = bw * t * bp * fw;
    return assemble3(n, N, bc, bp, a, b, al, ar);
}

int topo_dx(EllHost& out, int n, int N, double h, int bc, int dir) {
    if (n < 1 || n > DLT_NMAX || N < 1) { set_error("dgb_topo_dx: n=%d N=%d unsupported (1<=n<=20)", n, N); return DGB_ERR_INVALID; }
    if (bc != DGB_PER && N < 2) { set_error("dgb_topo_dx: N>=2 needed for non-periodic boundaries"); return DGB_ERR_INVALID; }
    switch (dir) {
        case DGB_CENTERED: out = dx_symm(n, N, h, bc); return 0;
        case DGB_FORWARD: out = dx_plus(n, N, h, bc); return 0;
        case DGB_BACKWARD: out = dx_minus(n, N, h, bc); return 0;
    }
    set_error("dgb_topo_dx: unknown direction %d", dir);
    return DGB_ERR_INVALID;
}
int topo_jump(EllHost& out, int n, int N, double h, int bc) {
    if (n < 1 || n > DLT_NMAX || N < 1) { set_error("dgb_topo_jump: n=%d N=%d unsupported", n, N); return DGB_ERR_INVALID; }
    out = jump1d(n, N, h, bc);
    return 0;
}

// ------------------------------------------------------------------------------------------------ grids
static int check_grid(const dgb_grid* g) {
    if (!g || g->ndim < 1 || g->ndim > 3) { set_error("dgb_grid: ndim must be 1..3"); return DGB_ERR_INVALID; }
    for (int u = 0; u < g->ndim; u++)
        if (g->n[u] < 1 || g->n[u] > DLT_NMAX || g->N[u] < 1 || !(g->x1[u] > g->x0[u])) {
            set_error("dgb_grid: invalid axis %d", u);
            return DGB_ERR_INVALID;
        }
    return 0;
}
double grid_h(const dgb_grid* g, int u) { return (g->x1[u] - g->x0[u]) / (double)g->N[u]; }
size_t grid_shape(const dgb_grid* g, int u) { return (size_t)g->n[u] * g->N[u]; }
size_t grid_size(const dgb_grid* g) { size_t s = 1; for (int u = 0; u < g->ndim; u++) s *= grid_shape(g, u); return s; }
// grid.h:128-147
std::vector<double> grid_abscissas(const dgb_grid* g, int u) {
    int n = g->n[u], N = g->N[u];
    std::vector<double> abs((size_t)n * N), aa = dlt_abscissas(n);
    double hu = grid_h(g, u);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < n; j++) {
            double xmiddle = std::fma(hu, (double)i, g->x0[u]);
            double h2 = hu / 2.;
            double absj = 1. + aa[j];
            abs[(size_t)i * n + j] = std::fma(h2, absj, xmiddle);
        }
    return abs;
}
// grid.h:155-166
std::vector<double> grid_weights1d(const dgb_grid* g, int u) {
    int n = g->n[u], N = g->N[u];
    std::vector<double> v((size_t)n * N), ww = dlt_weights(n);
    double hu = grid_h(g, u);
    for (int i = 0; i < N; i++)
        for (int j = 0; j < n; j++) v[(size_t)i * n + j] = hu / 2. * ww[j];
    return v;
}
// weights.h:60: kronecker(Product, w_0, w_1, ...) with Product(x0,x1,x2) = x0*(x1*x2) (subroutines.h:99-117)
std::vector<double> grid_weights(const dgb_grid* g) {
    std::vector<double> w0 = grid_weights1d(g, 0);
    if (g->ndim == 1) return w0;
    std::vector<double> w1 = grid_weights1d(g, 1);
    if (g->ndim == 2) {
        std::vector<double> w(w0.size() * w1.size());
        for (size_t j = 0; j < w1.size(); j++)
            for (size_t i = 0; i < w0.size(); i++) w[j * w0.size() + i] = w0[i] * w1[j];
        return w;
    }
    std::vector<double> w2 = grid_weights1d(g, 2);
    std::vector<double> w(w0.size() * w1.size() * w2.size());
    for (size_t k = 0; k < w2.size(); k++)
        for (size_t j = 0; j < w1.size(); j++)
            for (size_t i = 0; i < w0.size(); i++) w[(k * w1.size() + j) * w0.size() + i] = w0[i] * (w1[j] * w2[k]);
    return w;
}
// derivatives.h:22-33
void update_left_right(EllHost& m, const dgb_grid* g, int coord) {
    int right = 1, left = 1;
    for (int u = 0; u < coord; u++) right *= (int)grid_shape(g, u);
    for (int u = coord + 1; u < g->ndim; u++) left *= (int)grid_shape(g, u);
    m.right = right; m.left = left; m.rr0 = 0; m.rr1 = right;
}

// ------------------------------------------------------------------------------------------------ dG interpolation
// Legendre polynomials at xn in [-1,1] (interpolation.h:69-95)
static std::vector<double> legendre(double xn, int n) {
    std::vector<double> px(n);
    if (xn == -1) { for (int u = 0; u < n; u++) px[u] = (u % 2 == 0) ? +1. : -1.; }
    else if (xn == 1) { for (int i = 0; i < n; i++) px[i] = 1.; }
    else {
        px[0] = 1.;
        if (n > 1) {
            px[1] = xn;
            // a*b - c*d: the reference's FMA-enabled host build evaluates this as fma(a, b, -(c*d)) (pinned by
            // tests/test_topology.py against the reference's own matrices)
            for (int i = 1; i < n - 1; i++)
                px[i + 1] = std::fma((double)(2 * i + 1) * xn, px[i], -((double)i * px[i - 1])) / (double)(i + 1);
        }
    }
    return px;
}
// one row of the x-space dG interpolation matrix for a point X inside [x0,x1] of a 1d grid (interpolation.h:245-300);
// bc handling (mirror/shift of outside points) is the caller's business: here X must lie inside the grid
void interpolation_row_xspace(double X, double x0, double x1, int n, int N, std::vector<int>& cols, std::vector<double>& vals) {
    double h = (x1 - x0) / (double)N;
    double xnn = (X - x0) / h;
    unsigned nn = (unsigned)std::floor(xnn);
    double xn = 2. * xnn - (double)(2 * nn + 1);
    if ((int)nn == N) { nn -= 1; xn = 1.; }
    int idx = -1;
    std::vector<double> gauss = dlt_abscissas(n);
    for (int k = 0; k < n; k++)
        if (std::fabs(xn - gauss[k]) < 1e-13) idx = (int)nn * n + k;
    if (idx < 0) {
        std::vector<double> px = legendre(xn, n), pxF(n, 0.);
        Mat fw = dlt_forward(n);
        for (int l = 0; l < n; l++)
            for (int k = 0; k < n; k++) pxF[l] += px[k] * fw(k, l);
        for (int l = 0; l < n; l++) { cols.push_back((int)nn * n + l); vals.push_back(pxF[l]); }
    } else {
        cols.push_back(idx);
        vals.push_back(1.);
    }
}

// fast_interpolation.h:182-205: refine every cell into multiplyN cells with n*multiplyn coefficients
int topo_fast_interpolation1d(EllHost& out, int n, int N, int multiplyn, int multiplyN) {
    if (n < 1 || n * multiplyn > DLT_NMAX || multiplyn < 1 || multiplyN < 1) { set_error("dgb_topo_fast_interpolation: invalid factors"); return DGB_ERR_INVALID; }
    dgb_grid gnew{};
    gnew.ndim = 1; gnew.x0[0] = -1.; gnew.x1[0] = 1.; gnew.n[0] = n * multiplyn; gnew.N[0] = multiplyN;
    std::vector<double> xs = grid_abscissas(&gnew, 0);
    int size = multiplyn * multiplyN;
    // NOTE the reference declares the block size as t.n() although each dense block has n*multiplyn rows only when
    // multiplyn == 1 (fast_interpolation.h:190); we keep its semantics: blocks are n x n, `size` of them
    out = make_ell(size * N, N, 1, size, n);
    int ncols = n;  // interpolX.num_cols() = g_old.size()
    for (size_t row = 0; row < xs.size(); row++) {
        std::vector<int> cols; std::vector<double> vals;
        interpolation_row_xspace(xs[row], -1., 1., n, 1, cols, vals);
        for (size_t l = 0; l < cols.size(); l++) out.data[row * ncols + cols[l]] = vals[l];
    }
    for (int i = 0; i < size * N; i++) { out.cols_idx[i] = i / size; out.data_idx[i] = i % size; }
    return 0;
}
// fast_interpolation.h:228-258 with projection.h:107-147: P = V_new * I^T * W_old, value v_new*(val*w_old)
int topo_fast_projection1d(EllHost& out, int nold, int N, int dividen, int divideN) {
    if (dividen < 1 || divideN < 1 || N % divideN != 0 || nold % dividen != 0) {
        set_error("dgb_topo_fast_projection: N=%d / n=%d not divisible by %d / %d", N, nold, divideN, dividen);
        return DGB_ERR_INVALID;
    }
    int n = nold / dividen;
    dgb_grid gold{}, gnew{};
    gold.ndim = 1; gold.x0[0] = -1.; gold.x1[0] = 1.; gold.n[0] = n * dividen; gold.N[0] = divideN;
    gnew.ndim = 1; gnew.x0[0] = -1.; gnew.x1[0] = 1.; gnew.n[0] = n; gnew.N[0] = 1;
    std::vector<double> w_old = grid_weights1d(&gold, 0), w_new = grid_weights1d(&gnew, 0);
    std::vector<double> v_new(w_new.size());
    for (size_t i = 0; i < w_new.size(); i++) v_new[i] = 1. / w_new[i];  // INVERT functor (functors.h)
    std::vector<double> xs = grid_abscissas(&gold, 0);
    int size = dividen * divideN;
    out = make_ell(N / divideN, N * dividen, size, size, n);
    // interpolation(g_old <- g_new): row = old point, cols = new points; transposed entry (row=new i, col=old)
    for (size_t oldp = 0; oldp < xs.size(); oldp++) {
        std::vector<int> cols; std::vector<double> vals;
        interpolation_row_xspace(xs[oldp], -1., 1., n, 1, cols, vals);
        for (size_t l = 0; l < cols.size(); l++) {
            int row = cols[l], col = (int)oldp;
            double val = v_new[row] * (vals[l] * w_old[col]);
            int k = col / (n * dividen), ll = (col / n) % dividen, i = row, j = col % n;
            out.data[(((size_t)k * dividen + ll) * n + i) * n + j] = val;
        }
    }
    for (int i = 0; i < N / divideN; i++)
        for (int d = 0; d < size; d++) { out.cols_idx[(size_t)i * size + d] = i * size + d; out.data_idx[(size_t)i * size + d] = d; }
    return 0;
}

void ell_view(const EllHost& m, dgb_ell_host* v) {
    v->num_rows = m.num_rows; v->num_cols = m.num_cols; v->blocks_per_line = m.bpl; v->n = m.n;
    v->left_size = m.left; v->right_size = m.right; v->num_blocks = m.nblocks;
    v->right_range[0] = m.rr0; v->right_range[1] = m.rr1;
    v->data = m.data.data(); v->cols_idx = m.cols_idx.data(); v->data_idx = m.data_idx.data();
}

}  // namespace dgb

using namespace dgb;

extern "C" {
int dgb_topo_dlt(int which, int n, double* out) {
    if (n < 1 || n > DLT_NMAX || which < 0 || which > 3) { set_error("dgb_topo_dlt: invalid arguments"); return DGB_ERR_INVALID; }
    std::vector<double> v = which == 0 ? dlt_abscissas(n) : which == 1 ? dlt_weights(n) : which == 2 ? dlt_backward(n).a : dlt_forward(n).a;
    std::copy(v.begin(), v.end(), out);
    return 0;
}
// dg::create::window_stencil (inc/dg/topology/stencil.h:56-88,177-237): the neighbourhood matrix blas2::stencil runs on.
// 1-d: row k lists the columns k - w/2 ... k - w/2 + w - 1, all values 1; points beyond the boundary are wrapped (PER) or
// mirrored (value -1 on a Dirichlet side); duplicates are kept and nothing is sorted, as in the reference.  2-d / 3-d: the
// Kronecker product, last axis outermost.  Caller-allocated outputs: row_offsets[size + 1], cols / vals[size * prod(window)].
int dgb_topo_window_stencil(const dgb_grid* g, const int* window, int* row_offsets, int* cols, double* vals) {
    int e = check_grid(g); if (e) return e;
    if (!window || !row_offsets || !cols || !vals) { set_error("dgb_topo_window_stencil: null argument"); return DGB_ERR_INVALID; }
    std::vector<std::vector<int>> ax_cols(g->ndim);
    std::vector<std::vector<double>> ax_vals(g->ndim);
    std::vector<int> len(g->ndim);
    size_t per_row = 1;
    for (int u = 0; u < g->ndim; u++) {
        const int w = window[u], n = g->n[u] * g->N[u], radius = w / 2, bc = g->bc[u];
        if (w < 1) { set_error("dgb_topo_window_stencil: window size must be positive"); return DGB_ERR_INVALID; }
        len[u] = n;
        per_row *= (size_t)w;
        for (int k = 0; k < n; k++)
            for (int l = 0; l < w; l++) {
                int c = k + l - radius;
                double v = 1.;
                if (c < 0) {
                    if (bc == DGB_PER) c += n;
                    else { c = -(c + 1); if (bc == DGB_DIR || bc == DGB_DIR_NEU) v = -1.; }
                } else if (c >= n) {
                    if (bc == DGB_PER) c -= n;
                    else { c = 2 * n - 1 - c; if (bc == DGB_DIR || bc == DGB_NEU_DIR) v = -1.; }
                }
                ax_cols[u].push_back(c);
                ax_vals[u].push_back(v);
            }
    }
    const size_t rows = grid_size(g);
    if (rows * per_row > (size_t)INT32_MAX) { set_error("dgb_topo_window_stencil: matrix too large for int indices"); return DGB_ERR_UNSUPPORTED; }
    // row index = ((iz) * ny + iy) * nx + ix; entries of a row: outermost axis slowest (tensorproduct(my, mx), xspacelib.h:38-70)
    size_t counter = 0;
    row_offsets[0] = 0;
    std::vector<int> idx(g->ndim, 0), ent(g->ndim, 0);
    for (size_t r = 0; r < rows; r++) {
        size_t rr = r;
        for (int u = 0; u < g->ndim; u++) { idx[u] = (int)(rr % len[u]); rr /= len[u]; }
        for (size_t q = 0; q < per_row; q++) {
            size_t qq = q;
            for (int u = 0; u < g->ndim; u++) { ent[u] = (int)(qq % window[u]); qq /= window[u]; }
            long long c = 0;
            double v = 1.;
            for (int u = g->ndim - 1; u >= 0; u--) {
                const size_t at = (size_t)idx[u] * window[u] + ent[u];
                c = c * len[u] + ax_cols[u][at];
                v = v * ax_vals[u][at];
            }
            cols[counter] = (int)c;
            vals[counter] = v;
            counter++;
        }
        row_offsets[r + 1] = (int)counter;
    }
    return 0;
}
// dg::create::limiter_stencil (inc/dg/topology/stencil.h:89-137,199-256): the matrix dg::CSRSlopeLimiter runs on.  Along the
// limited axis the first row of every cell holds 3 n entries -- columns of the left neighbour cell with forward(0, j), of the cell
// with forward(1, j), of the right neighbour with backward(j, 1) -- and the cell's other n - 1 rows are empty; columns beyond
// the boundary are wrapped or mirrored (sign flip on a Dirichlet side) as in detail::set_boundary (stencil.h:19-53).  In 2-d the
// Kronecker product with the identity of the other axis (tensorproduct, xspacelib.h:38-70).  Outputs caller-allocated:
// row_offsets[size + 1], cols / vals[3 size].
int dgb_topo_limiter_stencil(const dgb_grid* g, int direction, int bound, int* row_offsets, int* cols, double* vals) {
    int e = check_grid(g); if (e) return e;
    if (!row_offsets || !cols || !vals) { set_error("dgb_topo_limiter_stencil: null argument"); return DGB_ERR_INVALID; }
    if (g->ndim > 2 || direction < 0 || direction >= g->ndim) { set_error("dgb_topo_limiter_stencil: 1-d and 2-d grids, direction < ndim"); return DGB_ERR_INVALID; }
    const int n = g->n[direction], N = g->N[direction], len = n * N;
    if (n == 1) { set_error("Limiter stencil not possible for n==1!"); return DGB_ERR_INVALID; }
    Mat fw = dlt_forward(n), bw = dlt_backward(n);
    std::vector<int> c1((size_t)3 * n * N);
    std::vector<double> v1((size_t)3 * n * N);
    for (int k = 0; k < N; k++)
        for (int part = 0; part < 3; part++)
            for (int j = 0; j < n; j++) {
                int c = (k - 1 + part) * n + j;
                double v = part == 0 ? fw(0, j) : (part == 1 ? fw(1, j) : bw(j, 1));
                if (c < 0) {
                    if (bound == DGB_PER) c += len;
                    else { c = -(c + 1); if (bound == DGB_DIR || bound == DGB_DIR_NEU) v *= -1; }
                } else if (c >= len) {
                    if (bound == DGB_PER) c -= len;
                    else { c = 2 * len - 1 - c; if (bound == DGB_DIR || bound == DGB_NEU_DIR) v *= -1; }
                }
                c1[((size_t)k * 3 + part) * n + j] = c;
                v1[((size_t)k * 3 + part) * n + j] = v;
            }
    const int nx = g->n[0] * g->N[0], ny = g->ndim == 2 ? g->n[1] * g->N[1] : 1;
    size_t counter = 0;
    row_offsets[0] = 0;
    for (int iy = 0; iy < ny; iy++)
        for (int ix = 0; ix < nx; ix++) {
            const int along = direction == 0 ? ix : iy;
            if (along % n == 0) {
                const size_t at = (size_t)(along / n) * 3 * n;
                for (int q = 0; q < 3 * n; q++) {
                    cols[counter] = direction == 0 ? iy * nx + c1[at + q] : c1[at + q] * nx + ix;
                    vals[counter] = v1[at + q] * 1.;
                    counter++;
                }
            }
            row_offsets[(size_t)iy * nx + ix + 1] = (int)counter;
        }
    return 0;
}
int dgb_topo_size(const dgb_grid* g, size_t* size) {
    int e = check_grid(g); if (e) return e;
    *size = grid_size(g);
    return 0;
}
int dgb_topo_abscissas(const dgb_grid* g, int axis, double* out) {
    int e = check_grid(g); if (e) return e;
    if (axis < 0 || axis >= g->ndim) { set_error("dgb_topo_abscissas: axis out of range"); return DGB_ERR_INVALID; }
    auto v = grid_abscissas(g, axis);
    std::copy(v.begin(), v.end(), out);
    return 0;
}
int dgb_topo_weights1d(const dgb_grid* g, int axis, double* out) {
    int e = check_grid(g); if (e) return e;
    if (axis < 0 || axis >= g->ndim) { set_error("dgb_topo_weights1d: axis out of range"); return DGB_ERR_INVALID; }
    auto v = grid_weights1d(g, axis);
    std::copy(v.begin(), v.end(), out);
    return 0;
}
int dgb_topo_weights(const dgb_grid* g, double* out) {
    int e = check_grid(g); if (e) return e;
    auto v = grid_weights(g);
    std::copy(v.begin(), v.end(), out);
    return 0;
}
static int finish(dgb_ellh** m, EllHost* h, int e) {
    if (e) { delete h; return e; }
    *m = reinterpret_cast<dgb_ellh*>(h);
    return 0;
}
int dgb_topo_dx(dgb_ellh** m, int n, int N, double h, int bc, int dir) {
    EllHost* A = new EllHost();
    return finish(m, A, topo_dx(*A, n, N, h, bc, dir));
}
int dgb_topo_jump(dgb_ellh** m, int n, int N, double h, int bc) {
    EllHost* A = new EllHost();
    return finish(m, A, topo_jump(*A, n, N, h, bc));
}
int dgb_topo_derivative(dgb_ellh** m, const dgb_grid* g, int coord, int bc, int dir) {
    int e = check_grid(g); if (e) return e;
    if (coord < 0 || coord >= g->ndim) { set_error("dgb_topo_derivative: coord>=Nd not allowed"); return DGB_ERR_INVALID; }
    EllHost* A = new EllHost();
    e = topo_dx(*A, g->n[coord], g->N[coord], grid_h(g, coord), bc, dir);
    if (!e) update_left_right(*A, g, coord);
    return finish(m, A, e);
}
int dgb_topo_jump_nd(dgb_ellh** m, const dgb_grid* g, int coord, int bc) {
    int e = check_grid(g); if (e) return e;
    if (coord < 0 || coord >= g->ndim) { set_error("dgb_topo_jump_nd: coord>=Nd not allowed"); return DGB_ERR_INVALID; }
    EllHost* A = new EllHost();
    e = topo_jump(*A, g->n[coord], g->N[coord], grid_h(g, coord), bc);
    if (!e) update_left_right(*A, g, coord);
    return finish(m, A, e);
}
int dgb_topo_fast_projection(dgb_ellh** m, const dgb_grid* g, int coord, int dividen, int divideN) {
    int e = check_grid(g); if (e) return e;
    if (coord < 0 || coord >= g->ndim) { set_error("dgb_topo_fast_projection: coord>=Nd not allowed"); return DGB_ERR_INVALID; }
    EllHost* A = new EllHost();
    e = topo_fast_projection1d(*A, g->n[coord], g->N[coord], dividen, divideN);
    if (!e) update_left_right(*A, g, coord);
    return finish(m, A, e);
}
int dgb_topo_fast_interpolation(dgb_ellh** m, const dgb_grid* g, int coord, int multiplyn, int multiplyN) {
    int e = check_grid(g); if (e) return e;
    if (coord < 0 || coord >= g->ndim) { set_error("dgb_topo_fast_interpolation: coord>=Nd not allowed"); return DGB_ERR_INVALID; }
    EllHost* A = new EllHost();
    e = topo_fast_interpolation1d(*A, g->n[coord], g->N[coord], multiplyn, multiplyN);
    if (!e) update_left_right(*A, g, coord);
    return finish(m, A, e);
}
int dgb_ellh_view(const dgb_ellh* m, dgb_ell_host* view) {
    if (!m || !view) { set_error("dgb_ellh_view: NULL argument"); return DGB_ERR_INVALID; }
    ell_view(*reinterpret_cast<const EllHost*>(m), view);
    return 0;
}
int dgb_ellh_destroy(dgb_ellh* m) { delete reinterpret_cast<EllHost*>(m); return 0; }
}
