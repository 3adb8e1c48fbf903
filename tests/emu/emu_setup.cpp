// tests/emu/emu_setup.cpp -- TEST INFRASTRUCTURE, not product code.
// Host emulation harness for the device setup pipeline: compiles the product's thread
// bodies (csrc/mft_setup_kernels.cuh) and its host orchestration (csrc/mft_setup_host.inl) with g++ and runs every
// "thread" in a host loop, so that the CPU-only test tier can check the kernels' logic and the orchestration against
// the oracle.  Nothing under meshfreetrixi.jl_b200/ links, loads or calls this file; the product runs the same thread
// bodies only as CUDA kernels (libmft_b200.so has no CPU path).
#include "../../meshfreetrixi.jl_b200/csrc/mft_setup_host.inl"

#include <cstdlib>
#include <cstring>

namespace {
struct HostEmu {
    std::string msg;
    size_t budget = (size_t)64 << 20;
    int alloc(void **p, size_t bytes)
    {
        *p = std::malloc(bytes);
        if (!*p) msg = "malloc failed";
        return *p ? 0 : 1;
    }
    void release(void *p) { std::free(p); }
    int h2d(void *d, const void *s, size_t b)
    {
        std::memcpy(d, s, b);
        return 0;
    }
    int d2h(void *d, const void *s, size_t b)
    {
        std::memcpy(d, s, b);
        return 0;
    }
    int launch_knn(const mft_setup::KnnArgs &A)
    {
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t q = 0; q < A.nq; ++q) mft_setup::knn_thread(A, q);
        return 0;
    }
    int launch_weights(const mft_setup::WeightArgs &A)
    {
#pragma omp parallel for schedule(dynamic, 64)
        for (int64_t t = 0; t < A.nthreads; ++t) mft_setup::weights_thread(A, t);
        return 0;
    }
    int sync() { return 0; }
    size_t scratch_budget() { return budget; }
    const char *error() { return msg.c_str(); }
};
thread_local std::string g_err;
}  // namespace

extern "C" {
const char *emu_last_error(void) { return g_err.c_str(); }
int emu_setup_knn(int64_t n, const double *x, const double *y, int k, int64_t nq, const int64_t *query_idx1, int64_t *nbr1_out,
                  double *dist_out)
{
    HostEmu be;
    return mft_setup::run_knn(be, n, x, y, k, nq, query_idx1, nbr1_out, dist_out, g_err);
}
int emu_setup_rbf_weights(int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1, int p, int degree, int kk,
                          double *wx, double *wy, int64_t scratch_bytes)
{
    HostEmu be;
    if (scratch_bytes > 0) be.budget = (size_t)scratch_bytes;
    return mft_setup::run_weights(be, n, x, y, n_rows, k, nbr1, p, degree, kk, wx, wy, g_err);
}
int emu_setup_rbf_weights_hybrid(int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1, int p, double alpha,
                                 double beta, double epsilon, int degree, int kk, double *wx, double *wy)
{
    HostEmu be;
    const double hyb[3] = {alpha, beta, epsilon};
    return mft_setup::run_weights(be, n, x, y, n_rows, k, nbr1, p, degree, kk, wx, wy, g_err, hyb);
}
}
