// tests/emu/emu_limiter.cpp -- TEST INFRASTRUCTURE, not product code.
// Host emulation of the Zhang-Shu limiter kernels (csrc/mft_limiter_kernels.cuh): the product's thread bodies run in a
// host loop, with the device-side data layout (AoS state records, column-major neighbour table), so the CPU test tier
// can compare them with the oracle.
#include "../../meshfreetrixi.jl_b200/csrc/mft_limiter_kernels.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int emu_limiter_zhang_shu(int64_t n, int k, const int64_t *nbr0 /* n x k row-major, 0-based */, double gamma, int npairs,
                                     const double *thresholds, const int *variables, double *u_soa /* 4 x n */)
{
    void *ubuf = nullptr, *tbuf = nullptr;
    if (posix_memalign(&ubuf, 64, sizeof(mft::ZsState) * (size_t)n) || posix_memalign(&tbuf, 64, sizeof(mft::ZsState) * (size_t)n)) return 1;
    mft::ZsState *u = static_cast<mft::ZsState *>(ubuf);
    for (int64_t i = 0; i < n; ++i)
        for (int v = 0; v < 4; ++v) u[i].a[v] = u_soa[v * n + i];
    std::vector<int> tab((size_t)n * k);
    for (int64_t d = 0; d < n; ++d)
        for (int q = 0; q < k; ++q) tab[(size_t)q * n + d] = (int)nbr0[d * k + q];
    std::vector<unsigned char> flag((size_t)n);
    for (int i = 0; i < npairs; ++i) {
        mft::ZsArgs A{tab.data(), k, n, ubuf, tbuf, flag.data(), thresholds[i], gamma, variables[i]};
        for (int64_t row = 0; row < n; ++row) mft::zs_detect_row(A, row);
        for (int64_t row = 0; row < n; ++row) mft::zs_apply_row(row, flag.data(), tbuf, ubuf);
    }
    for (int64_t i = 0; i < n; ++i)
        for (int v = 0; v < 4; ++v) u_soa[v * n + i] = u[i].a[v];
    std::free(ubuf);
    std::free(tbuf);
    return 0;
}

#include "../../meshfreetrixi.jl_b200/csrc/mft_aux_kernels.cuh"

// k_count_nonfinite's thread body over an AoS state (n x V)
extern "C" int64_t emu_count_nonfinite(int64_t n, int V, const double *u_aos)
{
    int64_t cnt = 0;
    for (int64_t row = 0; row < n; ++row) cnt += mft::nonfinite_in_row(u_aos, V, row);
    return cnt;
}
