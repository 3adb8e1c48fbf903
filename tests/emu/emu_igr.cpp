// tests/emu/emu_igr.cpp -- TEST INFRASTRUCTURE, not product code.
// Host emulation of the IGR source kernels (csrc/mft_igr_kernels.cuh): the product's thread bodies run in host loops in
// the product's launch sequence, on the device data layout (AoS state, sliced-ELL operator blobs with a dummy record for
// padding entries), with the block/grid reduction tree of igr_reduce_finish reproduced value for value.
#include "../../meshfreetrixi.jl_b200/csrc/mft_igr_kernels.cuh"

#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
using namespace mft_igr;

// lane 0 of a shfl_down tree over 32 values (out-of-range source lane = own value)
double warp_tree(const double *v32)
{
    double v[32], o[32];
    for (int l = 0; l < 32; ++l) v[l] = v32[l];
    for (int off = 16; off > 0; off >>= 1) {
        for (int l = 0; l < 32; ++l) o[l] = v[l];
        for (int l = 0; l < 32; ++l) v[l] = o[l] + (l + off < 32 ? o[l + off] : o[l]);
    }
    return v[0];
}
double block_tree(const double *v256)
{
    double t = 0.0;
    for (int w = 0; w < kBlock / 32; ++w) t += warp_tree(v256 + 32 * w);
    return t;
}
// the reduction of one launch: per-thread values -> partial[block] -> last block combines
template <class Body>
double launch_reduce(const IgrArgs &A, Body body)
{
    const int64_t nblocks = (A.n_rows + kBlock - 1) / kBlock;
    std::vector<double> partial((size_t)nblocks);
    for (int64_t b = 0; b < nblocks; ++b) {
        double v[kBlock];
        for (int t = 0; t < kBlock; ++t) {
            const int64_t row = b * kBlock + t;
            v[t] = row < A.n_rows ? body(A, row) : 0.0;
        }
        partial[(size_t)b] = block_tree(v);
    }
    double s[kBlock];
    for (int t = 0; t < kBlock; ++t) {
        s[t] = 0.0;
        for (int64_t b = t; b < nblocks; b += kBlock) s[t] += partial[(size_t)b];
    }
    return block_tree(s);
}
template <class Body>
void launch_plain(const IgrArgs &A, Body body)
{
    for (int64_t row = 0; row < A.n_rows; ++row) body(A, row);
}
}  // namespace

// n points (no halo), device order = caller order.  nbr0/wx/wy: n x k row-major, entries of a row ALREADY in the
// reference's summation order (ascending column).  Builds the sliced-ELL blobs like build_ell() (padding -> dummy
// record n with weight 0), runs the product's launch sequence, returns sigma, the updated du and the CG status.
extern "C" int emu_igr_apply(int64_t n, int k, const int64_t *nbr0, const double *wx, const double *wy, double alpha, int maxiter,
                             const double *u_soa, double *du_soa, double *sigma_out, double *status3, double *b_out /* nullable */)
{
    const int64_t nsl = (n + 31) / 32;
    std::vector<int> off((size_t)nsl + 1);
    for (int64_t s = 0; s <= nsl; ++s) off[(size_t)s] = (int)(s * k);
    std::vector<unsigned char> blob((size_t)nsl * k * 640 + 128, 0);
    for (int64_t s = 0; s < nsl; ++s) {
        unsigned char *b = blob.data() + (size_t)off[(size_t)s] * 640;
        int *idx = reinterpret_cast<int *>(b);
        double *bx = reinterpret_cast<double *>(b + (size_t)k * 128);
        double *by = reinterpret_cast<double *>(b + (size_t)k * 384);
        for (int q = 0; q < k * 32; ++q) idx[q] = (int)n;
        for (int64_t d = s * 32; d < n && d < (s + 1) * 32; ++d)
            for (int c = 0; c < k; ++c) {
                const size_t at = (size_t)c * 32 + (size_t)(d - s * 32);
                idx[at] = (int)nbr0[d * k + c];
                bx[at] = wx[d * k + c];
                by[at] = wy[d * k + c];
            }
    }
    void *ubuf = nullptr, *dbuf = nullptr;
    if (posix_memalign(&ubuf, 64, sizeof(State4) * (size_t)(n + 1)) || posix_memalign(&dbuf, 64, sizeof(State4) * (size_t)n)) return 1;
    State4 *u = static_cast<State4 *>(ubuf), *du = static_cast<State4 *>(dbuf);
    for (int64_t i = 0; i < n; ++i)
        for (int v = 0; v < 4; ++v) {
            u[i].a[v] = u_soa[v * n + i];
            du[i].a[v] = du_soa[v * n + i];
        }
    const double dummy[4] = {1.0, 0.0, 0.0, 1.0};  // the finite Euler state mft_ctx_create puts behind the last point
    std::memcpy(u[n].a, dummy, sizeof dummy);
    const int64_t np1 = n + 1;
    std::vector<double> vec((size_t)(4 * n + 4 * np1), 0.0);
    IgrScalars S{};
    S.prev_res = 1.0;
    S.maxiter = maxiter;
    S.done = 1;
    IgrArgs A;
    A.blob = blob.data();
    A.off = off.data();
    A.n_rows = n;
    A.u = ubuf;
    A.du = dbuf;
    A.alpha = alpha;
    double *v = vec.data();
    A.rho_inv = v;
    A.b = v + n;
    A.r = v + 2 * n;
    A.c = v + 3 * n;
    A.x = v + 4 * n;
    A.p = v + 4 * n + np1;
    A.t = v + 4 * n + 2 * np1;
    A.partial = nullptr;
    A.ticket = nullptr;
    A.S = &S;
    // launch_igr()'s sequence
    igr_rhs_final(A, launch_reduce(A, [](const IgrArgs &a, int64_t r) { return igr_rhs_row(a, r); }));
    for (int it = 0; it < maxiter; ++it) {
        if (!S.done) launch_plain(A, [](const IgrArgs &a, int64_t r) { igr_dir_row(a, r); });
        if (!S.done) launch_plain(A, [](const IgrArgs &a, int64_t r) { igr_grad_row(a, r); });
        if (!S.done) igr_apply_final(A, launch_reduce(A, [](const IgrArgs &a, int64_t r) { return igr_apply_row(a, r); }));
        if (!S.done) igr_update_final(A, launch_reduce(A, [](const IgrArgs &a, int64_t r) { return igr_update_row(a, r); }));
    }
    launch_plain(A, [](const IgrArgs &a, int64_t r) { igr_flux_row(a, r); });
    for (int64_t i = 0; i < n; ++i) {
        for (int q = 0; q < 4; ++q) du_soa[q * n + i] = du[i].a[q];
        sigma_out[i] = A.x[i];
        if (b_out) b_out[i] = A.b[i];
    }
    status3[0] = S.iter;
    status3[1] = S.res;
    status3[2] = S.res0;
    std::free(ubuf);
    std::free(dbuf);
    return 0;
}
