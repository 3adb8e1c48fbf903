// mft_layout_device.inl -- the union-tile layout built ON THE DEVICE (rest of SURVEY row f1: replaces the host threads of
// build_tiler_host for the default R = 1 tiles).  Included by mft_b200.cu after mft_layout_host.inl.
//
// The per-tile work is mft_tile_build.cuh (one source for host and device, byte-identical to build_tiler_host).  Here:
//   pass 1  k_tile_sizes   union size of every tile, (W, L) of every slice             (one tile per warp, dynamic tile queue)
//   host    prefix sums of those few integers -> uoff, boff, second-copy generator positions   (ntiles + nslices numbers)
//   pass 2  k_tile_emit    colouring, second copy, step words, weights, union lists, written in place into the device arrays
// plus the same two passes run on the CPU (tile_build_portable_host) for the byte-for-byte comparison with build_tiler_host.
#include "mft_tile_build.cuh"

namespace {

__global__ void __launch_bounds__(32) k_tile_sizes(mft::tb::Op A, int ntiles, int kmax, unsigned char *scratch, size_t stride, int *next,
                                                   int *nu_of, int *wl)
{
    if (threadIdx.x != 0) return;
    mft::tb::Scratch S;
    mft::tb::scratch_bind(S, scratch + (size_t)blockIdx.x * stride, kmax, 0);
    for (;;) {
        const int t = atomicAdd(next, 1);
        if (t >= ntiles) break;
        nu_of[t] = mft::tb::tile_sizes(A, t, S, wl);
    }
}

__global__ void __launch_bounds__(32) k_tile_emit(mft::tb::Op A, mft::tb::Flags F, int ntiles, int kmax, int cap_nu, unsigned char *scratch,
                                                  size_t stride, int *next, const unsigned long long *lcg_skip, mft::tb::Out O, int *nslots_of)
{
    if (threadIdx.x != 0) return;
    mft::tb::Scratch S;
    mft::tb::scratch_bind(S, scratch + (size_t)blockIdx.x * stride, kmax, cap_nu);
    for (;;) {
        const int t = atomicAdd(next, 1);
        if (t >= ntiles) break;
        nslots_of[t] = mft::tb::tile_emit(A, F, t, lcg_skip[t], S, O);
    }
}

// everything between the two passes: offsets of the layout from the tile / slice sizes (the arithmetic of build_tiler_host)
struct TilePlan {
    std::vector<int> uoff;
    std::vector<unsigned long long> lcg_skip;
    std::vector<long long> boff;
    long long blob_bytes = 0, usum = 0;
    int max_nu = 0, maxW = 0, maxL = 0;
    int64_t nsteps = 0;
};

int tile_plan(const std::vector<int> &nu_of, const std::vector<int> &wl, int64_t ntl, int64_t nsl, bool two_copies, TilePlan &P)
{
    P.uoff.assign((size_t)ntl + 1, 0);
    P.lcg_skip.assign((size_t)ntl + 1, 0);
    int64_t usum = 0;
    for (int64_t t = 0; t < ntl; ++t) {
        if (nu_of[(size_t)t] < 0) return fail(MFT_EINVAL, "device tile build: a row is longer than the scratch allows");
        P.max_nu = std::max(P.max_nu, nu_of[(size_t)t]);
        usum += nu_of[(size_t)t] + 1;
        if (usum > 0x7fffffffLL) return fail(MFT_EINVAL, "union lists exceed 32-bit offsets");
        P.uoff[(size_t)t + 1] = (int)usum;
        P.lcg_skip[(size_t)t + 1] = P.lcg_skip[(size_t)t] + (two_copies ? (uint64_t)std::max(nu_of[(size_t)t] - 1, 0) : 0);
    }
    P.usum = usum;
    P.boff.assign((size_t)nsl, 0);
    long long at = 0;
    for (int64_t s = 0; s < nsl; ++s) {
        P.boff[(size_t)s] = at;
        at += (long long)wl[2 * s] * kSlice * 2 + 2LL * wl[2 * s + 1] * kSlice * 8;
        P.maxW = std::max(P.maxW, wl[2 * s]);
        P.maxL = std::max(P.maxL, wl[2 * s + 1]);
        P.nsteps += wl[2 * s];
    }
    P.blob_bytes = at + 128;
    return MFT_OK;
}

mft::tb::Op tile_op_host(const mft_ctx *c, const Csr2 &A, int64_t nrows_dev)
{
    mft::tb::Op op{};
    op.ptr = reinterpret_cast<const long long *>(A.ptr.data());
    op.col = A.col.data();
    op.wx = A.wx.data();
    op.wy = A.wy.data();
    op.perm = c->have_perm ? c->perm.data() : nullptr;
    op.iperm = c->have_perm ? c->iperm.data() : nullptr;
    op.nrows_dev = nrows_dev;
    op.n_tot = (int)c->n_tot;
    return op;
}

int tile_kmax(const mft_ctx *c, const Csr2 &A, int64_t nrows_dev)
{
    int64_t kmax = 1;
    for (int64_t d = 0; d < nrows_dev; ++d) {
        const int64_t r = c->have_perm ? c->perm[d] : d;
        kmax = std::max(kmax, A.ptr[r + 1] - A.ptr[r]);
    }
    return (int)std::min<int64_t>(kmax, 1 << 20);
}

}  // namespace

// the two passes on the CPU, one tile after the other (test harness of mft_tile_build.cuh)
static int tile_build_portable_host(const mft_ctx *c, const Csr2 &A, int64_t nrows_dev, bool colour, bool two_copies, bool tune, HostTileR &out)
{
    using namespace mft::tb;
    const int64_t nsl = (nrows_dev + kSlice - 1) / kSlice, ntl = (nrows_dev + kRows - 1) / kRows;
    const Op op = tile_op_host(c, A, nrows_dev);
    const int kmax = tile_kmax(c, A, nrows_dev);
    std::vector<int> nu_of((size_t)ntl, 0);
    out.wl.assign(2 * (size_t)nsl, 0);
    {
        std::vector<unsigned char> buf(scratch_bytes(kmax, 0));
        Scratch S;
        scratch_bind(S, buf.data(), kmax, 0);
        for (int64_t t = 0; t < ntl; ++t) nu_of[(size_t)t] = tile_sizes(op, t, S, out.wl.data());
    }
    TilePlan P;
    CHECK(tile_plan(nu_of, out.wl, ntl, nsl, two_copies, P));
    out.uoff = P.uoff;
    out.boff = P.boff;
    out.blob.assign((size_t)P.blob_bytes, 0);
    out.ulist.assign((size_t)P.usum, 0);
    out.uslot.assign((size_t)P.usum * 2, 0);
    const Flags F{colour ? 1 : 0, two_copies ? 1 : 0, tune ? 1 : 0};
    const Out O{out.blob.data(), out.boff.data(), out.wl.data(), out.uoff.data(), out.ulist.data(), out.uslot.data()};
    int max_slot = 0;
    {
        std::vector<unsigned char> buf(scratch_bytes(kmax, std::max(P.max_nu, 1)));
        Scratch S;
        scratch_bind(S, buf.data(), kmax, std::max(P.max_nu, 1));
        for (int64_t t = 0; t < ntl; ++t) {
            const int ns = tile_emit(op, F, t, P.lcg_skip[(size_t)t], S, O);
            if (ns < 0) return fail(MFT_EINVAL, "portable tile build: scratch too small for tile %lld", (long long)t);
            if (ns + 1 > 4095) return fail(MFT_ENOTSUP, "union tile: %d slots in one block exceed the 12-bit slot field", ns);
            max_slot = std::max(max_slot, ns);
        }
    }
    out.R = 1;
    out.nslices = (int)nsl;
    out.ntiles = (int)ntl;
    out.maxW = P.maxW;
    out.maxL = P.maxL;
    out.sstride = ((max_slot + 1 + 7) / 8) * 8;
    out.nsteps = P.nsteps;
    out.ncopy = two_copies ? 2 : 1;
    out.nnz = 0;
    for (int64_t d = 0; d < nrows_dev; ++d) {
        const int64_t r = c->have_perm ? c->perm[d] : d;
        out.nnz += A.ptr[r + 1] - A.ptr[r];
    }
    if (out.ulist.empty()) {   // (no rows at all: the placeholder entries of build_tiler_host; the device path is not taken then)
        out.ulist.push_back(0);
        out.uslot.push_back(0);
        out.uslot.push_back(0);
    }
    return MFT_OK;
}

// R = 1 union tiles of A for device rows [0, nrows_dev), built by the two kernels above
static int build_tiler_device(mft_ctx *c, const Csr2 &A, int64_t nrows_dev, bool colour, bool two_copies, bool tune, DevTileR &out)
{
    using namespace mft::tb;
    NvtxRange range("union-tile layout (device)");
    const auto t0 = std::chrono::steady_clock::now();
    const int64_t nsl = (nrows_dev + kSlice - 1) / kSlice, ntl = (nrows_dev + kRows - 1) / kRows;
    if (ntl <= 0 || ntl > 0x7fffffffLL) return fail(MFT_EINVAL, "device tile build: %lld tiles", (long long)ntl);
    const int kmax = tile_kmax(c, A, nrows_dev);
    // the operator rows on the device (transient)
    DevBuf<long long> d_ptr;
    DevBuf<int> d_col, d_perm, d_iperm;
    DevBuf<double> d_wx, d_wy;
    struct Release {
        DevBuf<long long> &a;
        DevBuf<int> &b, &c2, &d;
        DevBuf<double> &e, &f;
        ~Release()
        {
            a.release();
            b.release();
            c2.release();
            d.release();
            e.release();
            f.release();
        }
    } rel{d_ptr, d_col, d_perm, d_iperm, d_wx, d_wy};
    CHECK(d_ptr.upload(std::vector<long long>(A.ptr.begin(), A.ptr.end())));
    CHECK(d_col.upload(A.col));
    CHECK(d_wx.upload(A.wx));
    CHECK(d_wy.upload(A.wy));
    if (c->have_perm) {
        CHECK(d_perm.upload(c->perm));
        CHECK(d_iperm.upload(c->iperm));
    }
    Op op{};
    op.ptr = d_ptr.p;
    op.col = d_col.p;
    op.wx = d_wx.p;
    op.wy = d_wy.p;
    op.perm = c->have_perm ? d_perm.p : nullptr;
    op.iperm = c->have_perm ? d_iperm.p : nullptr;
    op.nrows_dev = nrows_dev;
    op.n_tot = (int)c->n_tot;

    int sms = 148;
    {
        int dev = 0;
        CU(cudaGetDevice(&dev));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    auto workers_for = [&](size_t per_worker) {
        const size_t budget = (size_t)4 << 30;   // scratch of all workers together
        int64_t w = std::min<int64_t>(ntl, (int64_t)sms * 32);
        w = std::min<int64_t>(w, (int64_t)std::max<size_t>(1, budget / std::max<size_t>(per_worker, 1)));
        return (int)std::max<int64_t>(1, w);
    };
    DevBuf<int> d_next, d_nu, d_nslots;
    DevBuf<unsigned char> d_scratch;
    CHECK(d_next.alloc(1));
    CHECK(d_nu.alloc(ntl));
    CHECK(out.wl.alloc(2 * nsl));
    CU(cudaMemsetAsync(out.wl.p, 0, sizeof(int) * 2 * (size_t)nsl, c->stream));
    CU(cudaMemsetAsync(d_next.p, 0, sizeof(int), c->stream));
    struct Release2 {
        DevBuf<int> &a, &b, &c2;
        DevBuf<unsigned char> &d;
        ~Release2()
        {
            a.release();
            b.release();
            c2.release();
            d.release();
        }
    } rel2{d_next, d_nu, d_nslots, d_scratch};
    // ---- pass 1 ----
    {
        const size_t per = scratch_bytes(kmax, 0);
        const int workers = workers_for(per);
        CHECK(d_scratch.alloc((int64_t)(per * (size_t)workers)));
        k_tile_sizes<<<workers, 32, 0, c->stream>>>(op, (int)ntl, kmax, d_scratch.p, per, d_next.p, d_nu.p, out.wl.p);
        CU(cudaGetLastError());
    }
    std::vector<int> nu_of((size_t)ntl), wl(2 * (size_t)nsl);
    CU(cudaMemcpyAsync(nu_of.data(), d_nu.p, sizeof(int) * (size_t)ntl, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(wl.data(), out.wl.p, sizeof(int) * 2 * (size_t)nsl, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    TilePlan P;
    CHECK(tile_plan(nu_of, wl, ntl, nsl, two_copies, P));
    // ---- pass 2 ----
    DevBuf<unsigned long long> d_skip;
    CHECK(out.uoff.upload(P.uoff));
    CHECK(out.boff.upload(P.boff));
    CHECK(d_skip.upload(P.lcg_skip));
    struct Release3 {
        DevBuf<unsigned long long> &a;
        ~Release3() { a.release(); }
    } rel3{d_skip};
    CHECK(out.blob.alloc(P.blob_bytes));
    CHECK(out.ulist.alloc(P.usum));
    CHECK(out.uslot.alloc(P.usum * 2));
    CHECK(d_nslots.alloc(ntl));
    CU(cudaMemsetAsync(out.blob.p, 0, (size_t)P.blob_bytes, c->stream));
    CU(cudaMemsetAsync(d_next.p, 0, sizeof(int), c->stream));
    {
        const int cap_nu = std::max(P.max_nu, 1);
        const size_t per = scratch_bytes(kmax, cap_nu);
        const int workers = workers_for(per);
        CHECK(d_scratch.alloc((int64_t)(per * (size_t)workers)));
        const Flags F{colour ? 1 : 0, two_copies ? 1 : 0, tune ? 1 : 0};
        const Out O{out.blob.p, out.boff.p, out.wl.p, out.uoff.p, out.ulist.p, out.uslot.p};
        k_tile_emit<<<workers, 32, 0, c->stream>>>(op, F, (int)ntl, kmax, cap_nu, d_scratch.p, per, d_next.p, d_skip.p, O, d_nslots.p);
        CU(cudaGetLastError());
    }
    std::vector<int> nslots_of((size_t)ntl);
    CU(cudaMemcpyAsync(nslots_of.data(), d_nslots.p, sizeof(int) * (size_t)ntl, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    int max_slot = 0;
    for (int64_t t = 0; t < ntl; ++t) {
        const int ns = nslots_of[(size_t)t];
        if (ns < 0) return fail(MFT_EINVAL, "device tile build: scratch too small for tile %lld", (long long)t);
        if (ns + 1 > 4095) return fail(MFT_ENOTSUP, "union tile: %d slots in one block exceed the 12-bit slot field", ns);
        max_slot = std::max(max_slot, ns);
    }
    out.R = 1;
    out.ncopy = two_copies ? 2 : 1;
    out.nslices = (int)nsl;
    out.ntiles = (int)ntl;
    out.maxW = P.maxW;
    out.maxL = P.maxL;
    out.sstride = ((max_slot + 1 + 7) / 8) * 8;
    out.nunion = P.usum;
    out.nsteps = P.nsteps;
    out.nnz = 0;
    for (int64_t d = 0; d < nrows_dev; ++d) {
        const int64_t r = c->have_perm ? c->perm[d] : d;
        out.nnz += A.ptr[r + 1] - A.ptr[r];
    }
    if (getenv("MFT_TRACE"))
        fprintf(stderr, "[mft] union-tile layout on the device: %lld tiles, kmax %d, max union %d: %.3f s\n", (long long)ntl, kmax, P.max_nu,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    return MFT_OK;
}

// FNV-1a over every array of a host layout (the checksum MFT_TRACE prints for build_tiler_host)
static uint64_t tile_layout_fnv(const HostTileR &h)
{
    uint64_t fnv = 1469598103934665603ULL;
    auto mix = [&](const void *p, size_t bytes) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < bytes; ++i) fnv = (fnv ^ b[i]) * 1099511628211ULL;
    };
    mix(h.blob.data(), h.blob.size());
    mix(h.boff.data(), h.boff.size() * sizeof(long long));
    mix(h.wl.data(), h.wl.size() * sizeof(int));
    mix(h.uoff.data(), h.uoff.size() * sizeof(int));
    mix(h.ulist.data(), h.ulist.size() * sizeof(int));
    mix(h.uslot.data(), h.uslot.size() * sizeof(unsigned short));
    const int meta[7] = {h.R, h.nslices, h.ntiles, h.maxW, h.maxL, h.sstride, h.ncopy};
    mix(meta, sizeof meta);
    return fnv;
}

// the layout a context holds (which: 0 forward, 1 transposed operator) as a host copy: checksum for tests
extern "C" int mft_debug_tiler_checksum(mft_ctx *c, int which, unsigned long long *fnv_out, long long *bytes_out)
{
    NEED_CTX(c);
    if (!fnv_out) return fail(MFT_EINVAL, "mft_debug_tiler_checksum: null output");
    CHECK(mft_finalize(c));
    const DevTileR &d = which == 0 ? c->fwd_tiler : c->tra_tiler;
    if (!d.ready()) return fail(MFT_EINVAL, "mft_debug_tiler_checksum: this context has no union-tile layout for operator %d", which);
    HostTileR h;
    h.R = d.R;
    h.nslices = d.nslices;
    h.ntiles = d.ntiles;
    h.maxW = d.maxW;
    h.maxL = d.maxL;
    h.sstride = d.sstride;
    h.ncopy = d.ncopy;
    auto down = [&](auto &vec, const auto &buf) -> int {
        vec.resize((size_t)buf.n);
        if (buf.n) CU(cudaMemcpy(vec.data(), buf.p, sizeof(vec[0]) * (size_t)buf.n, cudaMemcpyDeviceToHost));
        return MFT_OK;
    };
    CHECK(down(h.blob, d.blob));
    CHECK(down(h.boff, d.boff));
    CHECK(down(h.wl, d.wl));
    CHECK(down(h.uoff, d.uoff));
    CHECK(down(h.ulist, d.ulist));
    CHECK(down(h.uslot, d.uslot));
    *fnv_out = tile_layout_fnv(h);
    if (bytes_out)
        *bytes_out = (long long)(h.blob.size() + h.boff.size() * 8 + h.wl.size() * 4 + h.uoff.size() * 4 + h.ulist.size() * 4 + h.uslot.size() * 2);
    return MFT_OK;
}

// Host-only: the portable builder (the code the device runs) against build_tiler_host on a random ragged banded operator
// (arguments as mft_debug_tile_selftest; layout bits: 1 colour, 2 two copies, 4 tuned second copy).  0 = every array identical.
static int tile_build_compare_run(mft_ctx &ctx, const Csr2 &A, int layout)
{
    HostTileR h, p;
    CHECK(build_tiler_host(&ctx, A, ctx.n_local, 1, (layout & 1) != 0, (layout & 2) != 0, h, (layout & 4) != 0));
    CHECK(tile_build_portable_host(&ctx, A, ctx.n_local, (layout & 1) != 0, (layout & 2) != 0, (layout & 4) != 0, p));
    if (h.ulist.size() != p.ulist.size() || h.ulist != p.ulist) return fail(MFT_EINVAL, "tile build compare: union lists differ");
    if (h.uoff != p.uoff) return fail(MFT_EINVAL, "tile build compare: uoff differs");
    if (h.wl != p.wl) return fail(MFT_EINVAL, "tile build compare: wl differs");
    if (h.boff != p.boff) return fail(MFT_EINVAL, "tile build compare: boff differs");
    if (h.uslot != p.uslot) {
        size_t i = 0;
        while (i < h.uslot.size() && h.uslot[i] == p.uslot[i]) ++i;
        return fail(MFT_EINVAL, "tile build compare: slot tables differ at entry %lld copy %d (%d vs %d)", (long long)(i / 2), (int)(i & 1),
                    (int)h.uslot[i], (int)p.uslot[i]);
    }
    if (h.blob != p.blob) {
        size_t i = 0;
        while (i < h.blob.size() && h.blob[i] == p.blob[i]) ++i;
        return fail(MFT_EINVAL, "tile build compare: blobs differ at byte %lld of %lld", (long long)i, (long long)h.blob.size());
    }
    if (h.R != p.R || h.nslices != p.nslices || h.ntiles != p.ntiles || h.maxW != p.maxW || h.maxL != p.maxL || h.sstride != p.sstride ||
        h.ncopy != p.ncopy || h.nnz != p.nnz || h.nsteps != p.nsteps)
        return fail(MFT_EINVAL, "tile build compare: meta data differ (sstride %d vs %d, maxW %d vs %d, nnz %lld vs %lld, nsteps %lld vs %lld)", h.sstride,
                    p.sstride, h.maxW, p.maxW, (long long)h.nnz, (long long)p.nnz, (long long)h.nsteps, (long long)p.nsteps);
    return MFT_OK;
}

extern "C" int mft_debug_tile_build_compare(int64_t n, int k, int layout, int with_perm, unsigned seed)
{
    if (n <= 0 || k <= 0 || k > n) return fail(MFT_EINVAL, "mft_debug_tile_build_compare: bad arguments");
    mft_ctx ctx;
    Csr2 A;
    selftest_operator(ctx, A, n, k, with_perm, seed);
    return tile_build_compare_run(ctx, A, layout);
}

extern "C" int mft_debug_tile_build_compare_csr(int64_t n, int64_t n_rows, const int64_t *rowptr, const int32_t *col, int layout, unsigned seed)
{
    if (n <= 0 || n_rows < 0 || n_rows > n || !rowptr || !col) return fail(MFT_EINVAL, "mft_debug_tile_build_compare_csr: bad arguments");
    mft_ctx ctx;
    Csr2 A;
    uint64_t st = 0;
    int kmax = 1;
    CHECK(selftest_operator_csr(ctx, A, n, n_rows, rowptr, col, seed, st, kmax));
    return tile_build_compare_run(ctx, A, layout);
}

// two_choice_feasible (mft_tile_build.cuh) against the augmenting-path matcher of the host builder: every instance with up to 3
// points (and the dummy's bank group blocked or not), then `trials` random instances with up to 8 points.  Returns the mismatches.
extern "C" int mft_debug_matcher_compare(unsigned seed, int trials, long long *mismatches)
{
    if (!mismatches) return fail(MFT_EINVAL, "mft_debug_matcher_compare: null output");
    long long bad = 0;
    auto one = [&](const int *o0, const int *o1, int kk, unsigned blocked) {
        BankMatcher M;
        M.kk = kk;
        M.blocked = blocked;
        for (int i = 0; i < kk; ++i) {
            M.opt[i][0] = o0[i];
            M.opt[i][1] = o1[i];
        }
        if (M.perfect() != mft::tb::two_choice_feasible(o0, o1, kk, blocked)) ++bad;
    };
    int o0[8], o1[8];
    for (int kk = 0; kk <= 3; ++kk) {
        long long total = 1;
        for (int i = 0; i < kk; ++i) total *= 64;
        for (long long code = 0; code < total; ++code) {
            long long cc = code;
            for (int i = 0; i < kk; ++i) {
                o0[i] = (int)(cc & 7);
                o1[i] = (int)((cc >> 3) & 7);
                cc >>= 6;
            }
            one(o0, o1, kk, 0u);
            one(o0, o1, kk, 1u);
        }
    }
    uint64_t st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(st >> 33); };
    for (int tr = 0; tr < trials; ++tr) {
        const int kk = 1 + (int)(rnd() % 8);
        const int span = 2 + (int)(rnd() % 7);   // few bank groups: infeasible instances are common
        for (int i = 0; i < kk; ++i) {
            o0[i] = (int)(rnd() % span);
            o1[i] = (int)(rnd() % span);
        }
        one(o0, o1, kk, (rnd() & 1u) ? 1u : 0u);
    }
    *mismatches = bad;
    return MFT_OK;
}
