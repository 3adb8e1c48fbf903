// mft_setup_kernels.cuh -- the setup pipeline on the device:
//
//   * exact k-nearest-neighbour search over a uniform cell grid, replacing the KDTree/knn calls of the PointData
//     constructor (src/domains/PointCloudDomain/geometry_primatives.jl:322-339): per point the k nearest points,
//     ascending by distance, the point itself first, exact distance ties ordered by ascending point index
//     (NearestNeighbors.jl leaves tie order implementation-defined; the host mirror and the
//     oracle canonicalise the same way).
//   * the per-point RBF-FD weight solve, replacing the body of compute_flux_operator
//     (src/solvers/pointcloudsolver/compute_operators.jl:409-453 first derivatives, :549-594 k-th derivatives):
//     shift/scale the stencil (:225-246), build the symmetric saddle-point matrix [R P; P' 0] from the polyharmonic
//     spline r^p and the monomials up to `degree` (:191-223, :265-267), right-hand side = derivatives of the basis at
//     the mirrored stencil with the centre nudged to (eps,eps) (:248-263), solve, rescale by the per-axis factor^k.
//
// One thread = one point; no shared memory, no intra-block communication.  Every thread body is a plain
// `MFT_HD` function of (arguments, thread id) so that the SAME source is (a) wrapped into the __global__ kernels below
// for the product and (b) compiled by g++ into the tests' emulation harness (tests/emu/), which runs the thread bodies
// in a host loop to check them against the oracle on a box without a GPU.  The product library never runs them on the
// host.  All arithmetic is FP64 with separate multiply/add (nvcc -fmad=false, g++ -ffp-contract=off), powers are
// formed from multiplications, sqrt and one division only -> emulation and device agree bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MFT_HD __host__ __device__ __forceinline__
#else
#define MFT_HD inline
#endif

namespace mft_setup {

constexpr int kMaxK = 64;     // widest stencil (reference defaults: 15/20/30/42, geometry_primatives.jl:197-198)
constexpr int kMaxPoly = 28;  // monomials of degree <= 6
constexpr double kEps = 2.220446049250313e-16;  // Base.eps()

// ---- kNN over a uniform cell grid ---------------------------------------------------------------------------------
// The host sorts the points by cell (stable counting sort: ascending point index inside a cell) and uploads the sorted
// coordinates.  Thread q serves the q-th point in cell order, so the threads of a warp walk the same few cells.
struct KnnArgs {
    int64_t n;        // indexed points
    int64_t nq;       // queries (= threads): n when every point queries itself
    const int *qpos;  // nullable: cell-order position of the t-th query's point (queries = a subset of the points, sorted
    const int *qrow;  //           by cell-order position); qrow[t] = output row of the t-th query
    int k;
    int gx, gy;             // cells per axis
    double x0, y0, h;       // lower-left corner of the grid, cell edge
    const double *sx, *sy;  // coordinates in cell order
    const int *sid;         // caller index (0-based) of the q-th point in cell order
    const int *scell;       // cell (cy*gx + cx) of the q-th point in cell order
    const int *cell_start;  // gx*gy + 1 offsets into the cell order
    int *nbr;               // out: nq x k row-major (all points: row = CALLER index), entries = caller indices (0-based)
    double *dist;           // out: nq x k row-major distances
};

struct KnnBest {
    double d[kMaxK];
    int i[kMaxK];
    int cnt;
};

MFT_HD void knn_scan(const KnnArgs &A, KnnBest &B, double qx, double qy, int first, int last)
{
    const int k = A.k;
    for (int j = first; j < last; ++j) {
        const double dx = A.sx[j] - qx, dy = A.sy[j] - qy;
        const double d = sqrt(dx * dx + dy * dy);
        const int id = A.sid[j];
        if (B.cnt == k && !(d < B.d[k - 1] || (d == B.d[k - 1] && id < B.i[k - 1]))) continue;
        int pos = B.cnt < k ? B.cnt : k - 1;
        while (pos > 0 && (B.d[pos - 1] > d || (B.d[pos - 1] == d && B.i[pos - 1] > id))) {
            B.d[pos] = B.d[pos - 1];
            B.i[pos] = B.i[pos - 1];
            --pos;
        }
        B.d[pos] = d;
        B.i[pos] = id;
        if (B.cnt < k) ++B.cnt;
    }
}

MFT_HD void knn_thread(const KnnArgs &A, int64_t t)
{
    const int64_t q = A.qpos ? A.qpos[t] : t;
    const double qx = A.sx[q], qy = A.sy[q];
    const int cell = A.scell[q];
    const int cx = cell % A.gx, cy = cell / A.gx;
    const int k = A.k;
    KnnBest B;
    B.cnt = 0;
    for (int r = 0;; ++r) {
        const int x_lo = cx - r, x_hi = cx + r, y_lo = cy - r, y_hi = cy + r;
        const int xa = x_lo > 0 ? x_lo : 0, xb = x_hi < A.gx - 1 ? x_hi : A.gx - 1;
        const int ya = y_lo > 0 ? y_lo : 0, yb = y_hi < A.gy - 1 ? y_hi : A.gy - 1;
        for (int yy = ya; yy <= yb; ++yy) {
            const int64_t row0 = (int64_t)yy * A.gx;
            if (yy == y_lo || yy == y_hi) {
                // top / bottom edge of the ring: a run of cells, contiguous in cell order
                knn_scan(A, B, qx, qy, A.cell_start[row0 + xa], A.cell_start[row0 + xb + 1]);
            } else {
                if (x_lo >= 0) knn_scan(A, B, qx, qy, A.cell_start[row0 + x_lo], A.cell_start[row0 + x_lo + 1]);
                if (x_hi <= A.gx - 1) knn_scan(A, B, qx, qy, A.cell_start[row0 + x_hi], A.cell_start[row0 + x_hi + 1]);
            }
        }
        const bool all_x = x_lo <= 0 && x_hi >= A.gx - 1, all_y = y_lo <= 0 && y_hi >= A.gy - 1;
        if (all_x && all_y) break;  // every cell was visited
        if (B.cnt == k) {
            // every unvisited point lies outside the (2r+1)^2 block of cells: its distance is at least the gap between
            // the query and the nearest block side that still has cells beyond it.  A small safety margin covers the
            // rounding of the cell assignment; an exact tie with the k-th distance keeps the search going (strict <).
            double gap = HUGE_VAL;
            if (x_lo > 0) gap = fmin(gap, qx - (A.x0 + (double)x_lo * A.h));
            if (x_hi < A.gx - 1) gap = fmin(gap, (A.x0 + (double)(x_hi + 1) * A.h) - qx);
            if (y_lo > 0) gap = fmin(gap, qy - (A.y0 + (double)y_lo * A.h));
            if (y_hi < A.gy - 1) gap = fmin(gap, (A.y0 + (double)(y_hi + 1) * A.h) - qy);
            if (B.d[k - 1] < gap - 1e-7 * A.h) break;
        }
    }
    const int64_t out = (int64_t)(A.qpos ? A.qrow[t] : A.sid[q]) * k;
    for (int j = 0; j < k; ++j) {
        A.nbr[out + j] = B.i[j];
        A.dist[out + j] = B.d[j];
    }
}

// ---- RBF-FD weights --------------------------------------------------------------------------------------------------
MFT_HD double ipow(double x, int e)  // x^e, e >= 0, by repeated multiplication (left to right)
{
    double r = 1.0;
    for (int i = 0; i < e; ++i) r = r * x;
    return r;
}
// s^(h2/2) for integer h2 of either sign: integer power times (for odd h2) one square root, reciprocal for negative
MFT_HD double half_pow(double s, int h2)
{
    const int odd = h2 & 1;
    const int half = (h2 - odd) / 2;  // exact: h2 - odd is even
    double r = half >= 0 ? ipow(s, half) : 1.0 / ipow(s, -half);
    if (odd) r = r * sqrt(s);
    return r;
}
// The radial basis as a function of s = r^2 (geometry_primatives.jl:211-262):
//   PolyharmonicSpline      phi = r^p                                  f(s) = s^q,  q = p/2
//   HybridGaussianPHS       phi = alpha exp(-(eps r)^2) + beta r^p     f(s) = alpha exp(-eps^2 s) + beta s^q
//   f^(m)(s) = [alpha (-eps^2)^m exp(-eps^2 s)] + beta q (q-1) ... (q-m+1) s^(q-m)
struct RbfKind {
    int p;          // odd power of the polyharmonic part
    int hybrid;     // 0: pure PHS (alpha, beta, eps2 unused)
    double alpha, beta, eps2;
};
MFT_HD double rbf_fd(const RbfKind &K, double s, int mm)
{
    const double q = 0.5 * (double)K.p;
    double c = 1.0;
    for (int i = 0; i < mm; ++i) c = c * (q - (double)i);
    const double phs = c * half_pow(s, K.p - 2 * mm);
    if (!K.hybrid) return phs;
    double g = K.alpha * exp(-(K.eps2 * s));
    for (int i = 0; i < mm; ++i) g = g * -K.eps2;
    return g + K.beta * phs;
}
// kk-th derivative along x of phi(x, y) = f(x^2 + y^2)
MFT_HD double rbf_axis_derivative(const RbfKind &K, double x, double s, int kk)
{
    switch (kk) {
    case 0: return rbf_fd(K, s, 0);
    case 1: return 2.0 * x * rbf_fd(K, s, 1);
    case 2: return 2.0 * rbf_fd(K, s, 1) + 4.0 * x * x * rbf_fd(K, s, 2);
    case 3: return 12.0 * x * rbf_fd(K, s, 2) + 8.0 * (x * x * x) * rbf_fd(K, s, 3);
    default: return 12.0 * rbf_fd(K, s, 2) + 48.0 * x * x * rbf_fd(K, s, 3) + 16.0 * ((x * x) * (x * x)) * rbf_fd(K, s, 4);
    }
}

struct WeightArgs {
    int64_t e0;        // first point of this launch
    int64_t nthreads;  // points of this launch
    int k;             // stencil width
    int degree;        // polynomial degree N: monomials x^a y^b, a + b <= N, ordered (d; a = d..0)
    int npoly;         // (N+1)(N+2)/2
    RbfKind rbf;       // radial basis (PHS r^p, or the Gaussian + PHS hybrid)
    int kk;            // derivative order 1..4 (1 = compute_flux_operator(solver, domain))
    const double *x, *y;  // caller order
    const int *nbr;       // n x k row-major, 0-based, self first
    double *scratch;      // (m*m + 2m) x stride doubles; entry e of thread t at scratch[e*stride + t]
    int64_t stride;
    double *wx, *wy;  // out: n x k row-major
    int *status;      // out: per point 0 = ok, 1 = singular / degenerate stencil
};

MFT_HD void weights_thread(const WeightArgs &A, int64_t t)
{
    const int64_t e = A.e0 + t;
    const int k = A.k, np = A.npoly, m = k + np;
    double *S = A.scratch + t;
    const int64_t st = A.stride;
#define MFT_MAT(r, c) S[((int64_t)(r) * m + (c)) * st]
#define MFT_RHS(r, d) S[((int64_t)m * m + (int64_t)(r) * 2 + (d)) * st]
    double xs[kMaxK], ys[kMaxK];
    const int *nb = A.nbr + e * k;
    // shift_stencil (compute_operators.jl:225-246): centre to the origin, each axis scaled by 1/max|.|
    const double xc = A.x[nb[0]], yc = A.y[nb[0]];
    double ax = 0.0, ay = 0.0;
    for (int j = 0; j < k; ++j) {
        xs[j] = A.x[nb[j]] - xc;
        ys[j] = A.y[nb[j]] - yc;
        ax = fmax(ax, fabs(xs[j]));
        ay = fmax(ay, fabs(ys[j]));
    }
    const double sx = 1.0 / ax, sy = 1.0 / ay;
    bool bad = !(ax > 0.0) || !(ay > 0.0);
    for (int j = 0; j < k; ++j) {
        xs[j] = xs[j] * sx;
        ys[j] = ys[j] * sy;
    }
    // [R P; P' 0]
    for (int i = 0; i < k; ++i) {
        for (int j = 0; j < k; ++j) {
            const double dx = xs[i] - xs[j], dy = ys[i] - ys[j];
            MFT_MAT(i, j) = rbf_fd(A.rbf, dx * dx + dy * dy, 0);
        }
        int c = k;
        for (int d = 0; d <= A.degree; ++d)
            for (int a = d; a >= 0; --a, ++c) {
                const double v = ipow(xs[i], a) * ipow(ys[i], d - a);
                MFT_MAT(i, c) = v;
                MFT_MAT(c, i) = v;
            }
    }
    for (int i = k; i < m; ++i)
        for (int j = k; j < m; ++j) MFT_MAT(i, j) = 0.0;
    // right-hand sides at the mirrored stencil x_c - x_j, the centre at (eps, eps)   (:248-263)
    for (int j = 0; j < k; ++j) {
        const double mx = j == 0 ? kEps : -xs[j], my = j == 0 ? kEps : -ys[j];
        const double s = mx * mx + my * my;
        MFT_RHS(j, 0) = rbf_axis_derivative(A.rbf, mx, s, A.kk);
        MFT_RHS(j, 1) = rbf_axis_derivative(A.rbf, my, s, A.kk);
    }
    {
        double fact = 1.0;
        for (int i = 2; i <= A.kk; ++i) fact = fact * (double)i;
        int c = k;
        for (int d = 0; d <= A.degree; ++d)
            for (int a = d; a >= 0; --a, ++c) {
                MFT_RHS(c, 0) = (a == A.kk && d - a == 0) ? fact : 0.0;
                MFT_RHS(c, 1) = (a == 0 && d - a == A.kk) ? fact : 0.0;
            }
    }
    // LU with partial pivoting (first largest magnitude), both right-hand sides carried along
    for (int c = 0; c < m && !bad; ++c) {
        int piv = c;
        double best = fabs(MFT_MAT(c, c));
        for (int r = c + 1; r < m; ++r) {
            const double v = fabs(MFT_MAT(r, c));
            if (v > best) {
                best = v;
                piv = r;
            }
        }
        if (!(best > 0.0) || best != best || best == HUGE_VAL) {
            bad = true;
            break;
        }
        if (piv != c) {
            for (int j = c; j < m; ++j) {
                const double tmp = MFT_MAT(c, j);
                MFT_MAT(c, j) = MFT_MAT(piv, j);
                MFT_MAT(piv, j) = tmp;
            }
            for (int d = 0; d < 2; ++d) {
                const double tmp = MFT_RHS(c, d);
                MFT_RHS(c, d) = MFT_RHS(piv, d);
                MFT_RHS(piv, d) = tmp;
            }
        }
        const double pv = MFT_MAT(c, c);
        for (int r = c + 1; r < m; ++r) {
            const double l = MFT_MAT(r, c) / pv;
            if (l == 0.0) continue;
            for (int j = c + 1; j < m; ++j) MFT_MAT(r, j) = MFT_MAT(r, j) - l * MFT_MAT(c, j);
            MFT_RHS(r, 0) = MFT_RHS(r, 0) - l * MFT_RHS(c, 0);
            MFT_RHS(r, 1) = MFT_RHS(r, 1) - l * MFT_RHS(c, 1);
        }
    }
    if (!bad) {
        for (int r = m - 1; r >= 0; --r) {
            double s0 = MFT_RHS(r, 0), s1 = MFT_RHS(r, 1);
            for (int j = r + 1; j < m; ++j) {
                const double a = MFT_MAT(r, j);
                s0 = s0 - a * MFT_RHS(j, 0);
                s1 = s1 - a * MFT_RHS(j, 1);
            }
            const double pv = MFT_MAT(r, r);
            MFT_RHS(r, 0) = s0 / pv;
            MFT_RHS(r, 1) = s1 / pv;
        }
    }
    // rescale: d^k/dx^k picks up the axis factor^k
    const double fx = ipow(sx, A.kk), fy = ipow(sy, A.kk);
    const double nan = HUGE_VAL - HUGE_VAL;
    for (int j = 0; j < k; ++j) {
        A.wx[e * k + j] = bad ? nan : fx * MFT_RHS(j, 0);
        A.wy[e * k + j] = bad ? nan : fy * MFT_RHS(j, 1);
    }
    A.status[e] = bad ? 1 : 0;
#undef MFT_MAT
#undef MFT_RHS
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(128) k_setup_knn(const KnnArgs A)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < A.nq) knn_thread(A, q);
}
__global__ void __launch_bounds__(128) k_setup_weights(const WeightArgs A)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < A.nthreads) weights_thread(A, t);
}
#endif

}  // namespace mft_setup
