// mft_setup_cuda.inl -- C ABI of the device setup pipeline: CUDA backend of
// mft_setup_host.inl + the two entry points.  Included by mft_b200.cu (uses its fail()/CU()/grid_for()).
#include "mft_setup_host.inl"

namespace {
struct CudaSetupBackend {
    std::string msg;
    int bad(cudaError_t e, const char *what)
    {
        msg = std::string(what) + ": " + cudaGetErrorString(e);
        return 1;
    }
    int alloc(void **p, size_t bytes)
    {
        const cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess) *p = nullptr;
        return e == cudaSuccess ? 0 : bad(e, "cudaMalloc");
    }
    void release(void *p) { cudaFree(p); }
    int h2d(void *d, const void *s, size_t b)
    {
        const cudaError_t e = cudaMemcpy(d, s, b, cudaMemcpyHostToDevice);
        return e == cudaSuccess ? 0 : bad(e, "cudaMemcpy (host to device)");
    }
    int d2h(void *d, const void *s, size_t b)
    {
        const cudaError_t e = cudaMemcpy(d, s, b, cudaMemcpyDeviceToHost);
        return e == cudaSuccess ? 0 : bad(e, "cudaMemcpy (device to host)");
    }
    int launch_knn(const mft_setup::KnnArgs &A)
    {
        mft_setup::k_setup_knn<<<grid_for(A.nq, 128), 128>>>(A);
        const cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : bad(e, "k_setup_knn launch");
    }
    int launch_weights(const mft_setup::WeightArgs &A)
    {
        mft_setup::k_setup_weights<<<grid_for(A.nthreads, 128), 128>>>(A);
        const cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? 0 : bad(e, "k_setup_weights launch");
    }
    int sync()
    {
        const cudaError_t e = cudaDeviceSynchronize();
        return e == cudaSuccess ? 0 : bad(e, "setup kernels");
    }
    // scratch of the weight solve: the matrices of one launch; 2 GiB keeps > 250 k points of a 30 x 30 system in flight
    size_t scratch_budget()
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (size_t)256 << 20;
        return std::max<size_t>((size_t)64 << 20, std::min<size_t>((size_t)2 << 30, free_b / 4));
    }
    const char *error() { return msg.c_str(); }
};

int setup_select_device(const char *who, int device)
{
    int ndev = 0;
    const cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MFT_ENODEVICE, "%s: no CUDA device available (%s); this library has no CPU fallback", who,
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(MFT_EINVAL, "%s: device %d out of range [0,%d)", who, device, ndev);
    CU(cudaSetDevice(device));
    return MFT_OK;
}
}  // namespace

extern "C" int mft_setup_knn(int device, int64_t n, const double *x, const double *y, int k, int64_t *nbr1_out, double *dist_out)
{
    CHECK(setup_select_device("mft_setup_knn", device));
    CudaSetupBackend be;
    std::string err;
    const int rc = mft_setup::run_knn(be, n, x, y, k, n, nullptr, nbr1_out, dist_out, err);
    if (rc) return fail(rc == -1 ? MFT_EINVAL : MFT_ECUDA, "%s", err.c_str());
    return MFT_OK;
}

extern "C" int mft_setup_knn_queries(int device, int64_t n, const double *x, const double *y, int k, int64_t nq, const int64_t *query_idx1,
                                     int64_t *nbr1_out, double *dist_out)
{
    CHECK(setup_select_device("mft_setup_knn_queries", device));
    if (!query_idx1 && nq != 0) return fail(MFT_EINVAL, "mft_setup_knn_queries: query_idx1 is NULL");
    if (nq == 0) return MFT_OK;
    CudaSetupBackend be;
    std::string err;
    const int rc = mft_setup::run_knn(be, n, x, y, k, nq, query_idx1, nbr1_out, dist_out, err);
    if (rc) return fail(rc == -1 ? MFT_EINVAL : MFT_ECUDA, "%s", err.c_str());
    return MFT_OK;
}

extern "C" int mft_setup_rbf_weights(int device, int64_t n, const double *x, const double *y, int k, const int64_t *nbr1, int phs_power,
                                     int poly_degree, int deriv_order, double *wx_out, double *wy_out)
{
    CHECK(setup_select_device("mft_setup_rbf_weights", device));
    CudaSetupBackend be;
    std::string err;
    const int rc = mft_setup::run_weights(be, n, x, y, n, k, nbr1, phs_power, poly_degree, deriv_order, wx_out, wy_out, err);
    if (rc) return fail(rc == -1 ? MFT_EINVAL : MFT_ECUDA, "%s", err.c_str());
    return MFT_OK;
}

extern "C" int mft_setup_rbf_weights_rows(int device, int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1,
                                          int phs_power, int poly_degree, int deriv_order, double *wx_out, double *wy_out)
{
    CHECK(setup_select_device("mft_setup_rbf_weights_rows", device));
    CudaSetupBackend be;
    std::string err;
    const int rc = mft_setup::run_weights(be, n, x, y, n_rows, k, nbr1, phs_power, poly_degree, deriv_order, wx_out, wy_out, err);
    if (rc) return fail(rc == -1 ? MFT_EINVAL : MFT_ECUDA, "%s", err.c_str());
    return MFT_OK;
}

/* HybridGaussianPHS basis (geometry_primatives.jl:117-132, 238-262): phi = alpha exp(-(epsilon r)^2) + beta r^phs_power */
extern "C" int mft_setup_rbf_weights_hybrid(int device, int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1,
                                            int phs_power, double alpha, double beta, double epsilon, int poly_degree, int deriv_order,
                                            double *wx_out, double *wy_out)
{
    CHECK(setup_select_device("mft_setup_rbf_weights_hybrid", device));
    CudaSetupBackend be;
    std::string err;
    const double hyb[3] = {alpha, beta, epsilon};
    const int rc = mft_setup::run_weights(be, n, x, y, n_rows, k, nbr1, phs_power, poly_degree, deriv_order, wx_out, wy_out, err, hyb);
    if (rc) return fail(rc == -1 ? MFT_EINVAL : MFT_ECUDA, "%s", err.c_str());
    return MFT_OK;
}
