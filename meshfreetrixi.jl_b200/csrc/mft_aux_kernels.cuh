// mft_aux_kernels.cuh -- failure detection on the resident state (the reference has none; Trixi's
// ode_unstable_check is imported at src/MeshfreeTrixi.jl:33 and never used).  Counts the non-finite entries (NaN, +-Inf)
// of the owned rows of u: integer count, atomic adds commute -> deterministic.  Thread body is host/device so that the
// tests' host emulation can run it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MFT_HD __host__ __device__ __forceinline__
#else
#define MFT_HD inline
#endif

namespace mft {

MFT_HD int nonfinite_in_row(const double *u, int V, int64_t row)
{
    int cnt = 0;
    for (int v = 0; v < V; ++v) {
        const double x = u[row * V + v];
        if (!(x - x == 0.0)) ++cnt;  // NaN - NaN and Inf - Inf are NaN
    }
    return cnt;
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) k_count_nonfinite(const double *__restrict__ u, int V, int64_t n_rows, unsigned long long *out)
{
    unsigned long long cnt = 0;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += (int64_t)gridDim.x * blockDim.x)
        cnt += (unsigned long long)nonfinite_in_row(u, V, row);
    if (cnt) atomicAdd(out, cnt);
}
#endif

}  // namespace mft
