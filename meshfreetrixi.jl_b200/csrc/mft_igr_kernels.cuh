// mft_igr_kernels.cuh -- information-geometric-regularisation source
//
// SourceIGR, src/sources/IGR.jl:  per rhs!
//   update_igr_rhs! (:117-158)  u_prim = cons2prim(u);  flux_x = Dx u_prim, flux_y = Dy u_prim (per field, forward CSC mul!);
//                               b = alpha * (trace^2 + trace_squared) with the reference's own indexing
//                               trace = flux_x[1] + flux_y[2],  trace_squared = flux_x[2]^2 + 2 flux_y[2] flux_x[3] + flux_y[3]^2
//   update_sigma!   (:169-191)  sigma .= 0; rho_inv = 1/rho;  linear_solver(sigma, A, b; maxiter = 20) with the composite operator
//                               A v = rho_inv .* v - alpha * (Dx (rho_inv .* Dx v) + Dy (rho_inv .* Dy v))      (mul!, :55-70)
//   apply           (:211-239)  du[2] += -Dx sigma,  du[3] += -Dy sigma   (mul_by_accum!(D, -1) on flux_igr = (0,sigma,0,0)/(0,0,sigma,0))
// linear_solver: IterativeSolvers.cg! (third party, unpinned; the struct default `cg` cannot be called with three
// positional arguments, so a working script passes cg!): r = b - A x, p = 0, rho_prev = 1, tol = sqrt(eps) * |r|, then while
// |r| > tol and it < maxiter:  beta = |r|^2 / |r_prev|^2;  p = r + beta p;  c = A p;  a = |r|^2 / (p.c);  x += a p;  r -= a c.
// IGR.jl is outside every @muladd scope: separate multiply and add throughout; sums over a row run in the reference's
// order (ascending caller column, the order the sliced-ELL rows are stored in).  Dot products and norms are BLAS calls in
// the reference (order unspecified): here a fixed tree (below), so parity of sigma is to rounding, not bit-exact.
//
// CG control lives on the device (IgrScalars): no host round trip, the whole source is a fixed launch sequence (CUDA-graph
// capturable); after convergence the remaining iterations' kernels return at once.
//
// One thread = one point.  Every thread body is an `MFT_HD` function of (arguments, row) returning the thread's
// contribution to the launch's reduction; the block/grid reduction around it (igr_reduce_finish) is the fixed tree
//   warp: shfl_down 16,8,4,2,1 -> 8 warp sums added serially -> partial[block];
//   last block: thread t adds partial[t], partial[t+256], ... serially, then the same block tree.
// The tests' host emulation (tests/emu/) runs the same bodies and reproduces this tree.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MFT_HD __host__ __device__ __forceinline__
#else
#define MFT_HD inline
#endif

namespace mft_igr {

constexpr int kBlock = 256;
constexpr double kRelTol = 1.4901161193847656e-08;  // sqrt(eps(Float64)): IterativeSolvers' default reltol

struct alignas(32) State4 {
    double a[4];
};

struct IgrScalars {
    double res;       // |r|
    double prev_res;  // |r| of the previous iteration (1 before the first)
    double tol;       // max(reltol * |r_0|, abstol = 0)
    double step;      // a = |r|^2 / (p.c)
    double res0;
    int iter, done, maxiter, pad;
};

struct IgrArgs {
    const unsigned char *blob;  // paired sliced-ELL forward operator [Dx, Dy]: slice s at blob + off[s]*640 =
    const int *off;             //   [idx w x 32 int32][wx w x 32 f64][wy w x 32 f64]
    int64_t n_rows;
    const void *u;  // State4 per point (+ the dummy record padding entries point at)
    void *du;
    double alpha;
    double *rho_inv, *b, *x, *r, *p, *c;  // x, p: n_tot + 1 (dummy slot = 0, gathered through the operator)
    double *t;                            // 2 x (n_tot + 1): rho_inv .* Dx p, rho_inv .* Dy p, interleaved (x, y)
    double *partial;                      // one per block
    unsigned int *ticket;
    IgrScalars *S;
};

struct RowView {
    const int *ip;
    const double *wx, *wy;
    int width;
};
MFT_HD RowView row_view(const IgrArgs &A, int64_t row)
{
    const int64_t slice = row >> 5;
    const int lane = (int)(row & 31);
    const int off0 = A.off[slice];
    RowView v;
    v.width = A.off[slice + 1] - off0;
    const unsigned char *base = A.blob + (size_t)off0 * 640;
    v.ip = reinterpret_cast<const int *>(base) + lane;
    v.wx = reinterpret_cast<const double *>(base + (size_t)v.width * 128) + lane;
    v.wy = reinterpret_cast<const double *>(base + (size_t)v.width * 384) + lane;
    return v;
}

// ---- launch 1: right-hand side b, rho_inv, CG start (x = 0, r = b, p = 0); reduction: |r|^2 -------------------------
MFT_HD double igr_rhs_row(const IgrArgs &A, int64_t row)
{
    const State4 *u = static_cast<const State4 *>(A.u);
    const RowView R = row_view(A, row);
    double fx1 = 0.0, fx2 = 0.0, fx3 = 0.0, fy2 = 0.0, fy3 = 0.0;
    for (int c = 0; c < R.width; ++c) {
        const State4 uj = u[R.ip[c * 32]];
        const double wx = R.wx[c * 32], wy = R.wy[c * 32];
        const double rho = uj.a[0], v1 = uj.a[1] / rho, v2 = uj.a[2] / rho;  // cons2prim: (rho, v1, v2, p); p is not used
        fx1 = fx1 + wx * rho;
        fx2 = fx2 + wx * v1;
        fx3 = fx3 + wx * v2;
        fy2 = fy2 + wy * v1;
        fy3 = fy3 + wy * v2;
    }
    const double trace = (0.0 + fx1) + fy2;
    const double trace_squared = (fx2 * fx2 + (2.0 * fy2) * fx3) + fy3 * fy3;
    const double b = A.alpha * (trace * trace + trace_squared);
    A.rho_inv[row] = 1.0 / u[row].a[0];
    A.b[row] = b;
    A.x[row] = 0.0;
    A.r[row] = b;
    A.p[row] = 0.0;
    return b * b;
}
MFT_HD void igr_rhs_final(const IgrArgs &A, double sum)
{
    IgrScalars &S = *A.S;
    S.res = sqrt(sum);
    S.prev_res = 1.0;
    S.res0 = S.res;
    S.tol = fmax(kRelTol * S.res, 0.0);
    S.step = 0.0;
    S.iter = 0;
    S.done = (S.res <= S.tol || S.iter >= S.maxiter) ? 1 : 0;
}

// ---- CG iteration, launch a: p = r + beta p ------------------------------------------------------------------------
MFT_HD void igr_dir_row(const IgrArgs &A, int64_t row)
{
    const IgrScalars &S = *A.S;
    const double beta = (S.res * S.res) / (S.prev_res * S.prev_res);
    A.p[row] = A.r[row] + beta * A.p[row];
}
// ---- launch b: t = rho_inv .* (Dx p, Dy p) --------------------------------------------------------------------------
MFT_HD void igr_grad_row(const IgrArgs &A, int64_t row)
{
    const RowView R = row_view(A, row);
    double gx = 0.0, gy = 0.0;
    for (int c = 0; c < R.width; ++c) {
        const double pj = A.p[R.ip[c * 32]];
        gx = gx + R.wx[c * 32] * pj;
        gy = gy + R.wy[c * 32] * pj;
    }
    const double ri = A.rho_inv[row];
    A.t[2 * row] = ri * gx;
    A.t[2 * row + 1] = ri * gy;
}
// ---- launch c: c = rho_inv .* p - alpha * (Dx t_x + Dy t_y); reduction: p . c --------------------------------------------
MFT_HD double igr_apply_row(const IgrArgs &A, int64_t row)
{
    const RowView R = row_view(A, row);
    double dx = 0.0, dy = 0.0;
    for (int c = 0; c < R.width; ++c) {
        const int64_t j = R.ip[c * 32];
        dx = dx + R.wx[c * 32] * A.t[2 * j];
        dy = dy + R.wy[c * 32] * A.t[2 * j + 1];
    }
    const double pi = A.p[row];
    double y = A.rho_inv[row] * pi;
    y = y - A.alpha * (dx + dy);
    A.c[row] = y;
    return pi * y;
}
MFT_HD void igr_apply_final(const IgrArgs &A, double dot)
{
    IgrScalars &S = *A.S;
    S.step = (S.res * S.res) / dot;
}
// ---- launch d: x += a p, r -= a c; reduction: |r|^2 --------------------------------------------------------------------
MFT_HD double igr_update_row(const IgrArgs &A, int64_t row)
{
    const double a = A.S->step;
    A.x[row] = A.x[row] + a * A.p[row];
    const double r = A.r[row] - a * A.c[row];
    A.r[row] = r;
    return r * r;
}
MFT_HD void igr_update_final(const IgrArgs &A, double sum)
{
    IgrScalars &S = *A.S;
    S.prev_res = S.res;
    S.res = sqrt(sum);
    S.iter = S.iter + 1;
    S.done = (S.res <= S.tol || S.iter >= S.maxiter) ? 1 : 0;
}
// ---- last launch: du[2] += -(Dx sigma), du[3] += -(Dy sigma)   (mul!(du_f, D, flux_f, -1, true): du += w * (sigma_j * -1)) ----
MFT_HD void igr_flux_row(const IgrArgs &A, int64_t row)
{
    const RowView R = row_view(A, row);
    State4 *du = static_cast<State4 *>(A.du) + row;
    double d1 = du->a[1], d2 = du->a[2];
    for (int c = 0; c < R.width; ++c) {
        const double sj = A.x[R.ip[c * 32]] * -1.0;
        d1 = d1 + R.wx[c * 32] * sj;
        d2 = d2 + R.wy[c * 32] * sj;
    }
    du->a[1] = d1;
    du->a[2] = d2;
}

#if defined(__CUDACC__)
// block tree + last-block combine; `fin` runs in thread 0 of the last block with the grid total
template <class Final>
__device__ __forceinline__ void igr_reduce_finish(const IgrArgs &A, double v, Final fin)
{
    __shared__ double sh[kBlock / 32];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (l == 0) sh[w] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kBlock / 32; ++k) t += sh[k];
        A.partial[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kBlock) s += __ldcg(&A.partial[b]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __syncthreads();
    if (l == 0) sh[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kBlock / 32; ++k) t += sh[k];
        fin(A, t);
        *A.ticket = 0;
    }
}

__global__ void __launch_bounds__(kBlock) k_igr_rhs(const IgrArgs A)
{
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const double v = row < A.n_rows ? igr_rhs_row(A, row) : 0.0;
    igr_reduce_finish(A, v, [](const IgrArgs &a, double s) { igr_rhs_final(a, s); });
}
__global__ void __launch_bounds__(kBlock) k_igr_dir(const IgrArgs A)
{
    if (A.S->done) return;
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row < A.n_rows) igr_dir_row(A, row);
}
__global__ void __launch_bounds__(kBlock) k_igr_grad(const IgrArgs A)
{
    if (A.S->done) return;
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row < A.n_rows) igr_grad_row(A, row);
}
__global__ void __launch_bounds__(kBlock) k_igr_apply(const IgrArgs A)
{
    if (A.S->done) return;
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const double v = row < A.n_rows ? igr_apply_row(A, row) : 0.0;
    igr_reduce_finish(A, v, [](const IgrArgs &a, double s) { igr_apply_final(a, s); });
}
__global__ void __launch_bounds__(kBlock) k_igr_update(const IgrArgs A)
{
    if (A.S->done) return;
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const double v = row < A.n_rows ? igr_update_row(A, row) : 0.0;
    igr_reduce_finish(A, v, [](const IgrArgs &a, double s) { igr_update_final(a, s); });
}
__global__ void __launch_bounds__(kBlock) k_igr_flux(const IgrArgs A)
{
    const int64_t row = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    if (row < A.n_rows) igr_flux_row(A, row);
}
#endif

}  // namespace mft_igr
