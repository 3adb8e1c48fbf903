// mft_limiter_kernels.cuh -- Zhang-Shu positivity limiter on the point cloud (stage callback)
//
// Trixi.limiter_zhang_shu!(u, threshold, variable, domain::PointCloudDomain{2}, ...)
// src/callbacks_stage/positivity_zhang_shu_point2d.jl:22-82.  Per point: minimum of `variable` over its kNN stencil
// (domain.pd.neighbors: distance-sorted, self first); where that is below the threshold the point is blended towards the
// stencil mean,  theta = (var(mean) - threshold) / (var(mean) - min),  u <- theta*u + (1-theta)*mean.  Jacobi style: all
// reads see the state before the pass (the reference builds local_u first and copies afterwards), hence two kernels.
// Arithmetic as the reference's: the mean adds the neighbours in list order and divides by k; the blend sits in an
// @muladd scope and StaticArrays' muladd(scalar, SVector, SVector) maps to per-component muladd -> fma(theta, u, (1-theta)*mean);
// Trixi's `pressure` has no fusable product.  (Third-party semantics, no reference test: parity unpinned.)
//
// One thread = one point, no intra-block communication: the thread bodies are `MFT_HD` functions so that the tests'
// host emulation (tests/emu/) can run the same source against the oracle on a box without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MFT_HD __host__ __device__ __forceinline__
#else
#define MFT_HD inline
#endif

namespace mft {

constexpr int ZS_VAR_DENSITY = 0;
constexpr int ZS_VAR_PRESSURE = 1;

struct alignas(32) ZsState {  // one Euler 2-D state record (same layout as Vec<4>)
    double a[4];
};

struct ZsArgs {
    const int *nbr;  // k x n_rows, column-major ([c][row]): device index of the c-th nearest neighbour
    int k;
    int64_t n_rows;
    const void *u;
    void *tmp;            // limited states
    unsigned char *flag;  // 1: row was limited
    double threshold, gamma;
    int variable;
};

MFT_HD double zs_variable(int variable, double gamma, const ZsState &u)
{
    if (variable == ZS_VAR_DENSITY) return u.a[0];
    return (gamma - 1.0) * (u.a[3] - 0.5 * (u.a[1] * u.a[1] + u.a[2] * u.a[2]) / u.a[0]);
}
// Julia min(::Float64, ::Float64): NaN-propagating
MFT_HD double jl_min(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return a < b ? a : b;
}

MFT_HD void zs_detect_row(const ZsArgs &A, int64_t row)
{
    const ZsState *u = static_cast<const ZsState *>(A.u);
    double vmin = HUGE_VAL;  // typemax(Float64)
    ZsState sum;
    for (int v = 0; v < 4; ++v) sum.a[v] = 0.0;
    constexpr int kBatch = 4;
    for (int c0 = 0; c0 < A.k; c0 += kBatch) {
        ZsState uj[kBatch];
        for (int b = 0; b < kBatch; ++b) {  // independent gathers first
            const int cc = c0 + b < A.k ? c0 + b : A.k - 1;
            uj[b] = u[A.nbr[(int64_t)cc * A.n_rows + row]];
        }
        for (int b = 0; b < kBatch; ++b) {
            if (c0 + b < A.k) {
                vmin = jl_min(vmin, zs_variable(A.variable, A.gamma, uj[b]));
                for (int v = 0; v < 4; ++v) sum.a[v] = sum.a[v] + uj[b].a[v];
            }
        }
    }
    unsigned char f = 0;
    if (vmin < A.threshold) {
        ZsState mean, lim;
        for (int v = 0; v < 4; ++v) mean.a[v] = sum.a[v] / (double)A.k;
        const double vmean = zs_variable(A.variable, A.gamma, mean);
        const double theta = (vmean - A.threshold) / (vmean - vmin);
        const ZsState ui = u[row];
        bool nonzero = false;
        for (int v = 0; v < 4; ++v) {
            lim.a[v] = fma(theta, ui.a[v], (1.0 - theta) * mean.a[v]);
            nonzero |= lim.a[v] != 0.0;  // local_u[element] != zero_el (NaN != 0 is true, as in Julia)
        }
        if (nonzero) {
            static_cast<ZsState *>(A.tmp)[row] = lim;
            f = 1;
        }
    }
    A.flag[row] = f;
}

MFT_HD void zs_apply_row(int64_t row, const unsigned char *flag, const void *tmp, void *u)
{
    if (flag[row]) static_cast<ZsState *>(u)[row] = static_cast<const ZsState *>(tmp)[row];
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(128) k_zs_detect(const ZsArgs A)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row < A.n_rows) zs_detect_row(A, row);
}

__global__ void __launch_bounds__(256) k_zs_apply(int64_t n_rows, const unsigned char *__restrict__ flag, const void *tmp, void *u)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row < n_rows) zs_apply_row(row, flag, tmp, u);
}
#endif

}  // namespace mft
