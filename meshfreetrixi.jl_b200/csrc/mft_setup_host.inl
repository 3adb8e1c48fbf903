// mft_setup_host.inl -- host orchestration of the device setup pipeline (row f1): argument checks, the cell grid and the
// stable counting sort of the points into cells, buffer management, chunked launches.  Written against a small backend
// concept so that one source serves the product (CUDA backend in mft_b200.cu: cudaMalloc / cudaMemcpy / <<<>>>) and
// the tests' emulation harness (tests/emu/: malloc / memcpy / a host loop over the thread bodies) -- the harness
// exercises this exact orchestration code on a box without a GPU.
//
// Backend concept:
//   int  alloc(void **p, size_t bytes);   void release(void *p);
//   int  h2d(void *dst, const void *src, size_t bytes);   int d2h(void *dst, const void *src, size_t bytes);
//   int  launch_knn(const mft_setup::KnnArgs &);   int launch_weights(const mft_setup::WeightArgs &);
//   int  sync();      size_t scratch_budget();     const char *error();
// all int results: 0 = ok.
#pragma once
#include "mft_setup_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>

namespace mft_setup {

template <class BE>
struct DevMem {
    BE &be;
    void *p = nullptr;
    explicit DevMem(BE &b) : be(b) {}
    DevMem(const DevMem &) = delete;
    DevMem &operator=(const DevMem &) = delete;
    ~DevMem()
    {
        if (p) be.release(p);
    }
    int alloc(size_t bytes) { return be.alloc(&p, bytes ? bytes : 8); }
    template <class T>
    int upload(const std::vector<T> &h)
    {
        if (int rc = alloc(sizeof(T) * h.size())) return rc;
        return h.empty() ? 0 : be.h2d(p, h.data(), sizeof(T) * h.size());
    }
    template <class T>
    T *as() const
    {
        return static_cast<T *>(p);
    }
};

inline int setup_fail(std::string &err, const char *fmt, long long a = 0, long long b = 0)
{
    char buf[256];
    snprintf(buf, sizeof buf, fmt, a, b);
    err = buf;
    return -1;
}

inline int setup_msg(std::string &err, const char *msg)
{
    err = msg ? msg : "backend error";
    return -2;
}

struct CellGrid {
    int gx = 1, gy = 1;
    double x0 = 0.0, y0 = 0.0, h = 1.0;
    std::vector<int> cell_start;  // gx*gy + 1
    std::vector<int> order;       // cell order -> caller index
    std::vector<int> cell_of;     // cell of the q-th point in cell order
};

// ~2 points per cell on average; ascending caller index inside a cell (stable counting sort)
inline int build_cell_grid(int64_t n, const double *x, const double *y, CellGrid &G, std::string &err)
{
    double xmin = x[0], xmax = x[0], ymin = y[0], ymax = y[0];
    for (int64_t i = 0; i < n; ++i) {
        if (!std::isfinite(x[i]) || !std::isfinite(y[i])) return setup_fail(err, "point %lld has a non-finite coordinate", (long long)i + 1);
        xmin = std::min(xmin, x[i]);
        xmax = std::max(xmax, x[i]);
        ymin = std::min(ymin, y[i]);
        ymax = std::max(ymax, y[i]);
    }
    const double ex = xmax - xmin, ey = ymax - ymin;
    const double target = std::max<double>(1.0, 0.5 * (double)n);
    double h;
    if (ex > 0.0 && ey > 0.0) h = std::sqrt(ex * ey / target);
    else if (ex > 0.0 || ey > 0.0) h = std::max(ex, ey) / target;
    else h = 1.0;
    if (!(h > 0.0) || !std::isfinite(h)) h = std::max(std::max(ex, ey), 1.0);
    auto cells = [&](double e) { return std::floor(e / h) + 1.0; };
    const double cell_cap = std::min(4.0 * (double)n + 16.0, 2.0e9);     // cell ids are 32-bit
    while (cells(ex) * cells(ey) > cell_cap) h *= 2.0;                   // very thin clouds: keep the grid O(n)
    G.gx = (int)cells(ex);
    G.gy = (int)cells(ey);
    G.x0 = xmin;
    G.y0 = ymin;
    G.h = h;
    const int64_t nc = (int64_t)G.gx * G.gy;
    std::vector<int> cell((size_t)n);
    G.cell_start.assign((size_t)nc + 1, 0);
    for (int64_t i = 0; i < n; ++i) {
        int cx = (int)std::floor((x[i] - xmin) / h), cy = (int)std::floor((y[i] - ymin) / h);
        cx = std::min(std::max(cx, 0), G.gx - 1);
        cy = std::min(std::max(cy, 0), G.gy - 1);
        cell[(size_t)i] = cy * G.gx + cx;
        G.cell_start[(size_t)cell[(size_t)i] + 1]++;
    }
    for (int64_t c = 0; c < nc; ++c) G.cell_start[(size_t)c + 1] += G.cell_start[(size_t)c];
    std::vector<int> cursor(G.cell_start.begin(), G.cell_start.end() - 1);
    G.order.resize((size_t)n);
    G.cell_of.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int q = cursor[(size_t)cell[(size_t)i]]++;
        G.order[(size_t)q] = (int)i;
        G.cell_of[(size_t)q] = cell[(size_t)i];
    }
    return 0;
}

// PointData(medusa_data, basis) neighbour search (geometry_primatives.jl:322-339).
// x, y: n coordinates; nbr1_out: nq x k row-major, 1-based, self first; dist_out (nullable): nq x k distances
// (column 2 holds what the reference reduces to dx_min / dx_avg).  query_idx1 == NULL: every point queries (nq = n, row i =
// point i); else the nq listed points (1-based) query against all n points -- what a rank of a partitioned cloud needs.
template <class BE>
int run_knn(BE &be, int64_t n, const double *x, const double *y, int k, int64_t nq, const int64_t *query_idx1, int64_t *nbr1_out,
            double *dist_out, std::string &err)
{
    if (n < 1 || !x || !y || !nbr1_out) return setup_fail(err, "setup_knn: bad arguments (n = %lld)", (long long)n);
    if (!query_idx1) nq = n;
    if (nq < 0) return setup_fail(err, "setup_knn: negative query count %lld", (long long)nq);
    if (nq == 0) return 0;
    if (k < 1 || k > kMaxK) return setup_fail(err, "setup_knn: k = %lld outside 1..%lld", k, kMaxK);
    if (k > n) return setup_fail(err, "setup_knn: k = %lld exceeds the number of points %lld", k, (long long)n);
    if (n > 2000000000LL) return setup_fail(err, "setup_knn: more than 2^31 points per device are not supported");
    CellGrid G;
    if (int rc = build_cell_grid(n, x, y, G, err)) return rc;
    std::vector<double> sx((size_t)n), sy((size_t)n);
    for (int64_t q = 0; q < n; ++q) {
        sx[(size_t)q] = x[G.order[(size_t)q]];
        sy[(size_t)q] = y[G.order[(size_t)q]];
    }
    // queries: threads in cell order of their points (neighbouring threads walk the same cells)
    std::vector<int> qpos, qrow;
    if (query_idx1) {
        std::vector<int> pos_of((size_t)n);
        for (int64_t q = 0; q < n; ++q) pos_of[(size_t)G.order[(size_t)q]] = (int)q;
        std::vector<std::pair<int, int>> byp((size_t)nq);
        for (int64_t s = 0; s < nq; ++s) {
            const int64_t id = query_idx1[s] - 1;
            if (id < 0 || id >= n) return setup_fail(err, "setup_knn: query index %lld out of range 1..%lld", (long long)query_idx1[s], (long long)n);
            byp[(size_t)s] = std::make_pair(pos_of[(size_t)id], (int)s);
        }
        std::sort(byp.begin(), byp.end());
        qpos.resize((size_t)nq);
        qrow.resize((size_t)nq);
        for (int64_t t = 0; t < nq; ++t) {
            qpos[(size_t)t] = byp[(size_t)t].first;
            qrow[(size_t)t] = byp[(size_t)t].second;
        }
    }
    DevMem<BE> d_sx(be), d_sy(be), d_sid(be), d_cell(be), d_start(be), d_nbr(be), d_dist(be), d_qpos(be), d_qrow(be);
    bool ok = !d_sx.upload(sx) && !d_sy.upload(sy) && !d_sid.upload(G.order) && !d_cell.upload(G.cell_of) && !d_start.upload(G.cell_start) &&
              !d_nbr.alloc(sizeof(int) * (size_t)nq * k) && !d_dist.alloc(sizeof(double) * (size_t)nq * k);
    if (ok && query_idx1) ok = !d_qpos.upload(qpos) && !d_qrow.upload(qrow);
    if (!ok) return setup_msg(err, be.error());
    KnnArgs A;
    A.n = n;
    A.nq = nq;
    A.qpos = query_idx1 ? d_qpos.template as<int>() : nullptr;
    A.qrow = query_idx1 ? d_qrow.template as<int>() : nullptr;
    A.k = k;
    A.gx = G.gx;
    A.gy = G.gy;
    A.x0 = G.x0;
    A.y0 = G.y0;
    A.h = G.h;
    A.sx = d_sx.template as<double>();
    A.sy = d_sy.template as<double>();
    A.sid = d_sid.template as<int>();
    A.scell = d_cell.template as<int>();
    A.cell_start = d_start.template as<int>();
    A.nbr = d_nbr.template as<int>();
    A.dist = d_dist.template as<double>();
    if (be.launch_knn(A) || be.sync()) return setup_msg(err, be.error());
    std::vector<int> nbr((size_t)nq * k);
    if (be.d2h(nbr.data(), d_nbr.p, sizeof(int) * nbr.size())) return setup_msg(err, be.error());
    for (size_t i = 0; i < nbr.size(); ++i) nbr1_out[i] = (int64_t)nbr[i] + 1;
    if (dist_out && be.d2h(dist_out, d_dist.p, sizeof(double) * (size_t)nq * k)) return setup_msg(err, be.error());
    return 0;
}

// compute_flux_operator (compute_operators.jl:409-453; k-th derivative :549-594) for a polyharmonic-spline basis r^p
// with monomials up to `degree`: per point the k stencil weights of d^kk/dx^kk and d^kk/dy^kk.
// nbr1: n_rows x k row-major, 1-based, self first (domain.pd.neighbors); wx_out, wy_out: n_rows x k row-major, aligned with
// nbr1.  n_rows = n for a whole cloud; a rank of a partitioned cloud passes its own rows only (indices stay global).
template <class BE>
int run_weights(BE &be, int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1, int p, int degree, int kk,
                double *wx_out, double *wy_out, std::string &err, const double *hybrid3 = nullptr /* alpha, beta, epsilon */)
{
    if (n < 1 || !x || !y || !nbr1 || !wx_out || !wy_out) return setup_fail(err, "setup_rbf_weights: bad arguments (n = %lld)", (long long)n);
    if (n_rows < 0) return setup_fail(err, "setup_rbf_weights: negative row count %lld", (long long)n_rows);
    if (n_rows == 0) return 0;
    if (k < 1 || k > kMaxK) return setup_fail(err, "setup_rbf_weights: k = %lld outside 1..%lld", k, kMaxK);
    if (degree < 0 || degree > 6) return setup_fail(err, "setup_rbf_weights: polynomial degree %lld outside 0..6", degree);
    if (kk < 1 || kk > 4) return setup_fail(err, "setup_rbf_weights: derivative order %lld outside 1..4", kk);
    if (p < 1 || (p % 2) == 0) return setup_fail(err, "setup_rbf_weights: polyharmonic spline power %lld must be odd and positive", p);
    if (hybrid3 && !(std::isfinite(hybrid3[0]) && std::isfinite(hybrid3[1]) && std::isfinite(hybrid3[2])))
        return setup_fail(err, "setup_rbf_weights: non-finite HybridGaussianPHS parameter");
    const int npoly = (degree + 1) * (degree + 2) / 2;
    if (npoly > k) return setup_fail(err, "setup_rbf_weights: %lld monomials need a stencil of at least that many points (k = %lld)", npoly, k);
    std::vector<int> nbr((size_t)n_rows * k);
    for (size_t i = 0; i < nbr.size(); ++i) {
        const int64_t j = nbr1[i] - 1;
        if (j < 0 || j >= n) return setup_fail(err, "setup_rbf_weights: neighbour index %lld out of range 1..%lld", (long long)nbr1[i], (long long)n);
        nbr[i] = (int)j;
    }
    const int m = k + npoly;
    const size_t per_thread = sizeof(double) * ((size_t)m * m + 2 * (size_t)m);
    int64_t chunk = (int64_t)(be.scratch_budget() / per_thread);
    chunk = std::max<int64_t>(256, chunk / 256 * 256);
    chunk = std::min<int64_t>(chunk, (n_rows + 255) / 256 * 256);
    DevMem<BE> d_x(be), d_y(be), d_nbr(be), d_scr(be), d_wx(be), d_wy(be), d_st(be);
    std::vector<double> hx(x, x + n), hy(y, y + n);
    bool ok = !d_x.upload(hx) && !d_y.upload(hy) && !d_nbr.upload(nbr) && !d_scr.alloc(per_thread * (size_t)chunk) &&
              !d_wx.alloc(sizeof(double) * (size_t)n_rows * k) && !d_wy.alloc(sizeof(double) * (size_t)n_rows * k) &&
              !d_st.alloc(sizeof(int) * (size_t)n_rows);
    if (!ok) return setup_msg(err, be.error());
    WeightArgs A;
    A.k = k;
    A.degree = degree;
    A.npoly = npoly;
    A.rbf.p = p;
    A.rbf.hybrid = hybrid3 ? 1 : 0;
    A.rbf.alpha = hybrid3 ? hybrid3[0] : 0.0;
    A.rbf.beta = hybrid3 ? hybrid3[1] : 1.0;
    A.rbf.eps2 = hybrid3 ? hybrid3[2] * hybrid3[2] : 0.0;
    A.kk = kk;
    A.x = d_x.template as<double>();
    A.y = d_y.template as<double>();
    A.nbr = d_nbr.template as<int>();
    A.scratch = d_scr.template as<double>();
    A.stride = chunk;
    A.wx = d_wx.template as<double>();
    A.wy = d_wy.template as<double>();
    A.status = d_st.template as<int>();
    for (int64_t e0 = 0; e0 < n_rows; e0 += chunk) {
        A.e0 = e0;
        A.nthreads = std::min<int64_t>(chunk, n_rows - e0);
        if (be.launch_weights(A)) return setup_msg(err, be.error());  // same stream: launches serialise on the scratch
    }
    if (be.sync()) return setup_msg(err, be.error());
    std::vector<int> status((size_t)n_rows);
    if (be.d2h(status.data(), d_st.p, sizeof(int) * (size_t)n_rows) || be.d2h(wx_out, d_wx.p, sizeof(double) * (size_t)n_rows * k) ||
        be.d2h(wy_out, d_wy.p, sizeof(double) * (size_t)n_rows * k))
        return setup_msg(err, be.error());
    for (int64_t i = 0; i < n_rows; ++i)
        if (status[(size_t)i]) return setup_fail(err, "setup_rbf_weights: singular local system in row %lld (degenerate stencil)", (long long)i + 1);
    return 0;
}

}  // namespace mft_setup
