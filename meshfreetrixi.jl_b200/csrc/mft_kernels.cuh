// mft_kernels.cuh -- hand-written sm_100a kernels of the rhs! hot path (FP64, HBM-bound sparse stencil work).
//
// Compiled with -fmad=false: a fused multiply-add appears ONLY where the code says fma(), which is exactly
// where the reference's `@muladd` scope (Trixi flux/cons2prim, history.jl) forms one.  SparseArrays' mul!
// (the Dx/Dy applications) is outside @muladd -> separate multiply and add in EXACT mode.
//
// Data layout in HBM
//   state (u, du, uprev, approx_du, history slots) : AoS, one Vec<V> (V doubles, 8V-byte aligned) per point
//   g = eps .* (Dx u, Dy u)                        : AoS, one Vec<2V> per point (x-part then y-part)
//   operators                                      : sliced ELL, slice = 32 consecutive device rows (one warp);
//        entry (slice s, column c, lane l) lives at ((off[s] + c) * 32 + l) in idx[] / wx[] / wy[];
//        padding entries point at a dummy record one past the last point (a finite state / zeros) with weight 0:
//        they add an exact 0, keep the loops branch-free, and all padded lanes of a request hit ONE sector.
//        Within a row the entries are stored in the reference's summation order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace mft {

struct P2PPeers;
struct P2PLocal;
struct RowAux;

constexpr int kSlice = 32;
constexpr double kEps = 2.220446049250313e-16;  // Base.eps()

template <int V>
struct alignas(8 * V) Vec {
    double a[V];
};

// One operator = one array of per-slice blobs.  Slice s (32 rows, width w = off[s+1]-off[s] columns) occupies
// COLB*w bytes at base + off[s]*COLB:   [ idx : w x 32 x int32 ][ wx : w x 32 x double ][ wy : w x 32 x double ]
// (single-weight operators have no wy block).  A blob is contiguous and 128-byte aligned, so a warp brings it
// (or its index block) into shared memory with ONE bulk-async copy (cp.async.bulk, the 1-D TMA path) and the
// entries of lane l sit at stride 32 -> conflict-free shared-memory reads.
struct EllBlob {
    const unsigned char *base;
    const int *off;  // nslices + 1 column offsets
};
constexpr int kColBytes2 = kSlice * (4 + 8 + 8);  // 640: paired (Dx,Dy) operator
constexpr int kColBytes1 = kSlice * (4 + 8);      // 384: single-weight operator
constexpr int kColBytesPair = kSlice * (4 + 4 * 8);  // 1152: row-pair operator (idx + wxA, wyA, wxB, wyB)

// ---- mbarrier + bulk-async copy (TMA 1-D) -----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// Per-warp staging of one slice.  STAGE_W: the whole blob goes to shared memory; otherwise only the index block
// does (the address-dependency chain idx -> gather then never waits on DRAM), the weight blocks are read with
// coalesced loads and the weight blocks of the slice `pf_dist` ahead are pulled into L2 by a bulk prefetch.
struct SliceView {
    const int *ip;       // + c*32 : column c, this lane
    const double *wxp;   // + c*32
    const double *wyp;   // + c*32 (paired only)
    int width;
};
template <int COLB, bool STAGE_W>
__device__ __forceinline__ SliceView stage_slice(const EllBlob &op, int64_t slice, int64_t n_slices, unsigned char *buf,
                                                 uint64_t *bar, int lane, int pf_dist, bool two_phase = false)
{
    const int off = op.off[slice];
    const int width = op.off[slice + 1] - off;
    const unsigned char *src = op.base + (size_t)off * COLB;
    if (lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (lane == 0 && width > 0) {
        // two_phase (paired operators, exact order): stage [idx | wx] now, wy replaces wx after the first sweep
        const uint32_t bytes = (uint32_t)width * (STAGE_W ? (two_phase ? kSlice * 12 : COLB) : kSlice * 4);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(buf, src, bytes, bar);
        if (STAGE_W && two_phase) bulk_prefetch_l2(src + (size_t)width * kSlice * 12, (uint32_t)width * kSlice * 8);
        if (!STAGE_W) {
            const int64_t s2 = slice + pf_dist;
            if (s2 < n_slices) {
                const int off2 = op.off[s2];
                const int w2 = op.off[s2 + 1] - off2;
                if (w2 > 0) bulk_prefetch_l2(op.base + (size_t)off2 * COLB + (size_t)w2 * kSlice * 4, (uint32_t)w2 * (COLB - kSlice * 4));
            }
            if (slice < pf_dist) bulk_prefetch_l2(src + (size_t)width * kSlice * 4, (uint32_t)width * (COLB - kSlice * 4));
        }
    }
    SliceView v;
    v.width = width;
    v.ip = reinterpret_cast<const int *>(buf) + lane;
    const unsigned char *wbase = STAGE_W ? buf : src;
    v.wxp = reinterpret_cast<const double *>(wbase + (size_t)width * kSlice * 4) + lane;
    v.wyp = reinterpret_cast<const double *>(wbase + (size_t)width * kSlice * ((STAGE_W && two_phase) ? 4 : 12)) + lane;
    return v;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The three kernels of a fused stage are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel may start
// (and run the part of its prologue that only touches operator data) while the tail of its predecessor is still running;
// pdl_wait() returns once the predecessor grid has completed and its memory operations are visible.  Both instructions are
// no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- 256-bit / 64-bit read-only loads and stores ------------------------------------------------------
__device__ __forceinline__ Vec<4> ld_ro(const Vec<4> *p)
{
    Vec<4> v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
        : "=d"(v.a[0]), "=d"(v.a[1]), "=d"(v.a[2]), "=d"(v.a[3])
        : "l"(p));
    return v;
}
__device__ __forceinline__ Vec<1> ld_ro(const Vec<1> *p)
{
    Vec<1> v;
    v.a[0] = __ldg(&p->a[0]);
    return v;
}
__device__ __forceinline__ Vec<8> ld_ro(const Vec<8> *p)
{
    Vec<8> v;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
        : "=d"(v.a[0]), "=d"(v.a[1]), "=d"(v.a[2]), "=d"(v.a[3])
        : "l"(p));
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4+32];"
        : "=d"(v.a[4]), "=d"(v.a[5]), "=d"(v.a[6]), "=d"(v.a[7])
        : "l"(p));
    return v;
}
__device__ __forceinline__ Vec<2> ld_ro(const Vec<2> *p)
{
    Vec<2> v;
    const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
    v.a[0] = t.x;
    v.a[1] = t.y;
    return v;
}
__device__ __forceinline__ void st_vec(Vec<4> *p, const Vec<4> &v)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a[0]), "d"(v.a[1]), "d"(v.a[2]), "d"(v.a[3])
                 : "memory");
}
__device__ __forceinline__ void st_vec(Vec<8> *p, const Vec<8> &v)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a[0]), "d"(v.a[1]), "d"(v.a[2]), "d"(v.a[3])
                 : "memory");
    asm volatile("st.global.v4.f64 [%0+32], {%1,%2,%3,%4};" ::"l"(p), "d"(v.a[4]), "d"(v.a[5]), "d"(v.a[6]),
                 "d"(v.a[7])
                 : "memory");
}
__device__ __forceinline__ void st_vec(Vec<1> *p, const Vec<1> &v) { p->a[0] = v.a[0]; }
__device__ __forceinline__ void st_vec(Vec<2> *p, const Vec<2> &v)
{
    *reinterpret_cast<double2 *>(p) = make_double2(v.a[0], v.a[1]);
}

// read-once operator data (weights read straight from global memory): do not allocate in L1, leave it to the gathers
__device__ __forceinline__ double ld_stream(const double *p)
{
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Julia max(::Float64, ::Float64): NaN-propagating
__device__ __forceinline__ double jl_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}

// ---- physics (Trixi flux / cons2prim, third-party; inside @muladd) ---------------
constexpr int EQ_EULER2D = 0;
constexpr int EQ_ADVECTION2D = 1;

__device__ __forceinline__ void euler_prim(double gamma, const Vec<4> &u, double &v1, double &v2, double &p)
{
    v1 = u.a[1] / u.a[0];
    v2 = u.a[2] / u.a[0];
    const double s = fma(u.a[1], v1, u.a[2] * v2);
    const double e = fma(-0.5, s, u.a[3]);
    p = (gamma - 1.0) * e;
}

template <int EQ, int V>
struct Physics;

template <>
struct Physics<EQ_EULER2D, 4> {
    double gamma;
    double v1, v2, p;
    __device__ __forceinline__ void prepare(const Vec<4> &u) { euler_prim(gamma, u, v1, v2, p); }
    __device__ __forceinline__ Vec<4> flux_x(const Vec<4> &u) const
    {
        Vec<4> f;
        f.a[0] = u.a[1];
        f.a[1] = fma(u.a[1], v1, p);
        f.a[2] = u.a[1] * v2;
        f.a[3] = (u.a[3] + p) * v1;
        return f;
    }
    __device__ __forceinline__ Vec<4> flux_y(const Vec<4> &u) const
    {
        Vec<4> f;
        f.a[0] = u.a[2];
        f.a[1] = u.a[2] * v1;
        f.a[2] = fma(u.a[2], v2, p);
        f.a[3] = (u.a[3] + p) * v2;
        return f;
    }
};

template <>
struct Physics<EQ_ADVECTION2D, 1> {
    double a1, a2;
    __device__ __forceinline__ void prepare(const Vec<1> &) {}
    __device__ __forceinline__ Vec<1> flux_x(const Vec<1> &u) const { return Vec<1>{{a1 * u.a[0]}}; }
    __device__ __forceinline__ Vec<1> flux_y(const Vec<1> &u) const { return Vec<1>{{a2 * u.a[0]}}; }
};

template <int V>
__device__ __forceinline__ bool lex_less(const double *a, const double *b)
{
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (a[v] < b[v]) return true;
        if (a[v] > b[v]) return false;
    }
    return false;
}

// ---- pass A: fused flux evaluation + Dx/Dy stencil apply (+ D u, viscosity limiter, g = eps .* D u) --------
// replaces calc_fluxes! (rbfsolver.jl:247-265) and the forward half of the viscosity sources
// (hyperviscosity.jl:246-349, 364-373 / 393-402).  One thread per row; a warp owns one ELL slice, so the
// index / weight streams are read as fully coalesced 128-B / 256-B requests and each neighbour state is one
// 32-byte sector fetched with a single 256-bit load.
struct PassAArgs {
    EllBlob op;
    int64_t n_slices;
    int buf_bytes;  // shared-memory bytes per warp
    int pf_dist;    // L2 prefetch distance in slices
    int two_phase;  // exact order + staged weights: one weight buffer, refilled with wy after the x sweep
    int dummy;      // index of the dummy record (one past the last point): padding / batch tails gather it
    const void *u;
    void *du;
    void *g;
    const void *approx_du;
    const double *norms;  // residual mode: norm_parts x V candidates (one per rank) to be combined, or the final V norms
    int norm_parts;       // 0: norms[] is final; >0: combine norm_parts candidates (lexicographic or per component)
    int norm_lex;
    double *norms_out;    // nullable: row 0 publishes the combined norms (diagnostics)
    int64_t n_rows;
    double eqp0, eqp1;
    double c_uw, c_rv, dx_avg;
    int success_iter_zero;
    int accumulate;  // 1: start from the du in memory (calc_fluxes! semantics); 0: start from 0 (after reset_du!)
    // diagnostics (nullable)
    double *eps_uw, *eps_rv, *eps, *eps_c;
    void *residual;
    // fused step (mft_fused_kernels.cuh).  stats = sum[4] | mean[4] | norms[4] | .. | raw norms at [16]; norm_miss counts the
    // rows whose deviation from the mean exceeds the norms pass A was given (the leaf statistic missed a rounding tie)
    const double *stats;          // nullable: no check
    unsigned long long *norm_miss;
    // several GPUs, fused step: g rows on the send list go straight into the peers' halo tails (band tiles only), block 0
    // merges the ranks' norm records, epilogues wait for norm_ready
    const struct P2PPeers *P;     // nullable
    struct P2PLocal *L;
    const int *aux;
    const struct RowAux *rows;
    const int *route_peer;
    const long long *route_dst;
    double divisor;
    int norm_merge;               // 0: norms[] is final (one GPU: the stage kernel's last block wrote it); 1: block 0 merges the ranks' records
};

constexpr int VISC_NONE = 0;
constexpr int VISC_UPWIND = 1;
constexpr int VISC_RESIDUAL = 2;

// per-row epilogue of pass A: store du, evaluate the viscosity limiter (update_upwind_visc!, update_residual_visc!,
// update_visc!, hyperviscosity.jl:246-349) and store g = eps .* (Dx u, Dy u)
template <int V, int EQ, bool DO_FLUX, int VISC>
__device__ __forceinline__ void pass_a_epilogue(const PassAArgs &A, int64_t row, const Vec<V> &acc, const Vec<V> &gx,
                                                const Vec<V> &gy, const Vec<V> &ui, const Vec<V> &ad, Vec<2 * V> *g_ret = nullptr,
                                                const double *norms_reg = nullptr)
{
    if constexpr (DO_FLUX) st_vec(reinterpret_cast<Vec<V> *>(A.du) + row, acc);

    if constexpr (VISC != VISC_NONE) {
        static_assert(VISC == VISC_NONE || (EQ == EQ_EULER2D && V == 4), "viscosity sources are Euler-2D only");
        // update_upwind_visc!  hyperviscosity.jl:246-285
        double v1, v2, p;
        euler_prim(A.eqp0, ui, v1, v2, p);
        const double speed = sqrt(v1 * v1 + v2 * v2);
        double sound;
        if (p < 0.0 || ui.a[0] < 0.0) {
            sound = 0.0;
        } else {
            sound = sqrt(A.eqp0 * p / ui.a[0]);
        }
        const double e_uw = A.c_uw * 0.5 * A.dx_avg * (speed + sound);
        double e = e_uw, e_rv = 0.0, e_c = 1.0;
        if constexpr (VISC == VISC_RESIDUAL) {
            // update_residual_visc! :289-329 (pointwise part) and update_visc! :331-349
            Vec<V> dui;
            if constexpr (DO_FLUX) {
                dui = acc;
            } else {
                dui = reinterpret_cast<const Vec<V> *>(A.du)[row];
            }
            Vec<V> res;
#pragma unroll
            for (int v = 0; v < V; ++v) res.a[v] = fabs(ad.a[v] - dui.a[v]);
            double nrm[V];
            if (A.norm_parts > 0) {
                // global ode_maximum: combine the per-rank candidates in rank order (MPI.Allreduce(MAX), mpi.jl:76)
#pragma unroll
                for (int v = 0; v < V; ++v) nrm[v] = A.norms[v];
                for (int r = 1; r < A.norm_parts; ++r) {
                    double c[V];
#pragma unroll
                    for (int v = 0; v < V; ++v) c[v] = A.norms[r * V + v];
                    if (A.norm_lex) {
                        if (lex_less<V>(nrm, c)) {
#pragma unroll
                            for (int v = 0; v < V; ++v) nrm[v] = c[v];
                        }
                    } else {
#pragma unroll
                        for (int v = 0; v < V; ++v) nrm[v] = jl_max(nrm[v], c[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < V; ++v) nrm[v] = nrm[v] == 0.0 ? kEps : nrm[v];
                if (A.norms_out && row == 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) A.norms_out[v] = nrm[v];
                }
            } else {
                // (norms_reg: block 0 of this very launch wrote the norms; the caller fetched them after seeing norm_ready)
#pragma unroll
                for (int v = 0; v < V; ++v) nrm[v] = norms_reg ? norms_reg[v] : A.norms[v];
            }
            double mx = res.a[0] / nrm[0];
#pragma unroll
            for (int v = 1; v < V; ++v) mx = jl_max(mx, res.a[v] / nrm[v]);
            e_rv = 0.5 * A.c_rv * (A.dx_avg * A.dx_avg) * mx;
            if (isnan(e_rv) || isinf(e_rv) || A.success_iter_zero) {
                if (isnan(e_uw) || isinf(e_uw)) {
                    e = kEps;
                    e_c = 2.0;
                } else {
                    e = e_uw;
                    e_c = 1.0;
                }
            } else {
                e = e_rv < e_uw ? e_rv : e_uw;
                e_c = e_rv < e_uw ? 0.0 : 1.0;
            }
            if (A.residual) st_vec(reinterpret_cast<Vec<V> *>(A.residual) + row, res);
        }
        if (A.eps) {
            A.eps[row] = e;
            A.eps_uw[row] = e_uw;
            A.eps_rv[row] = e_rv;
            A.eps_c[row] = e_c;
        }
        Vec<2 * V> gout;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            gout.a[v] = e * gx.a[v];
            gout.a[V + v] = e * gy.a[v];
        }
        st_vec(reinterpret_cast<Vec<2 * V> *>(A.g) + row, gout);
        if (g_ret) *g_ret = gout;
    }
}

// KFIX > 0: every slice is exactly KFIX columns wide (the forward operator of a kNN cloud) and the kernel is the
// single-sweep exact variant: the x-chain runs in the sweep while the y-products w_y * (-G) are parked in registers
// (fully unrolled), then the y-chain is appended in order.  Same rounding sequence as the two-sweep form, but each
// neighbour state is gathered once and its flux evaluated once.
template <int V, int EQ, bool EXACT, bool DO_FLUX, int VISC, bool STAGE_W, int KFIX = 0>
__global__ void __launch_bounds__(128, (KFIX > 0 ? 2 : 4)) k_pass_a(const PassAArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * 4 + warp;
    if (slice >= A.n_slices) return;  // warps are independent: per-warp barriers only
    const int64_t row = slice * kSlice + lane;
    const bool live = row < A.n_rows;
    const bool two_phase = EXACT && STAGE_W && KFIX == 0 && A.two_phase;
    const SliceView sv = stage_slice<kColBytes2, STAGE_W>(A.op, slice, A.n_slices, smem_dyn + (size_t)warp * A.buf_bytes,
                                                         &bars[warp], lane, A.pf_dist, two_phase);
    const int width = sv.width;
    const int *ip = sv.ip;
    const double *wxp = sv.wxp;
    const double *wyp = sv.wyp;
    const Vec<V> *__restrict__ u = reinterpret_cast<const Vec<V> *>(A.u);

    Physics<EQ, V> ph;
    if constexpr (EQ == EQ_EULER2D) {
        ph.gamma = A.eqp0;
    } else {
        ph.a1 = A.eqp0;
        ph.a2 = A.eqp1;
    }

    Vec<V> acc, gx, gy;
#pragma unroll
    for (int v = 0; v < V; ++v) acc.a[v] = gx.a[v] = gy.a[v] = 0.0;
    if (DO_FLUX && A.accumulate && live) acc = reinterpret_cast<const Vec<V> *>(A.du)[row];
    // own-row operands of the viscosity limiter: issued before the wait so they overlap the bulk copy
    Vec<V> ui, ad;
#pragma unroll
    for (int v = 0; v < V; ++v) ui.a[v] = ad.a[v] = 0.0;
    if constexpr (VISC != VISC_NONE) {
        if (live) {
            ui = ld_ro(u + row);
            if constexpr (VISC == VISC_RESIDUAL) ad = ld_ro(reinterpret_cast<const Vec<V> *>(A.approx_du) + row);
        }
    }
    if (width > 0) mbar_wait(&bars[warp], 0);

    // Every loop below runs in batches of kBatch columns: all index / weight / neighbour-state loads of a batch
    // are issued before any arithmetic, so a thread keeps kBatch independent gathers in flight (the FP64 division
    // in the flux has a slow-path branch that otherwise stops the compiler from overlapping iterations).
    // Columns past `width` are clamped to the last column with weight 0 (adds an exact zero).
    constexpr int kBatch = KFIX > 0 ? 4 : 5;
    if constexpr (EXACT && KFIX > 0 && DO_FLUX) {
        static_assert(KFIX % 4 == 0, "KFIX must be a multiple of the batch size");
        double yp[KFIX][V];
#pragma unroll
        for (int c0 = 0; c0 < KFIX; c0 += kBatch) {
            Vec<V> uj[kBatch];
            double wa[kBatch], wb[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const int j = ip[(c0 + b) * kSlice];
                wa[b] = wxp[(c0 + b) * kSlice];
                wb[b] = wyp[(c0 + b) * kSlice];
                uj[b] = ld_ro(u + j);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                ph.prepare(uj[b]);
                const Vec<V> f = ph.flux_x(uj[b]);
                const Vec<V> h = ph.flux_y(uj[b]);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    acc.a[v] = acc.a[v] + wa[b] * (-f.a[v]);
                    yp[c0 + b][v] = wb[b] * (-h.a[v]);
                }
                if constexpr (VISC != VISC_NONE) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        gx.a[v] = gx.a[v] + wa[b] * uj[b].a[v];
                        gy.a[v] = gy.a[v] + wb[b] * uj[b].a[v];
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < KFIX; ++c) {
#pragma unroll
            for (int v = 0; v < V; ++v) acc.a[v] = acc.a[v] + yp[c][v];
        }
    } else if constexpr (EXACT) {
        // reference order: all Dx terms in ascending column order, then all Dy terms, separate mul and add
        // (SparseArrays mul!).  alpha = -1 is folded into the flux first: w * (-F).
        for (int c0 = 0; c0 < width; c0 += kBatch) {
            Vec<V> uj[kBatch];
            double w[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const bool ok = c0 + b < width;
                const int cc = ok ? c0 + b : width - 1;
                const int j = ok ? ip[cc * kSlice] : A.dummy;
                const double wv = wxp[cc * kSlice];
                w[b] = ok ? wv : 0.0;
                uj[b] = ld_ro(u + j);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                if constexpr (DO_FLUX) {
                    ph.prepare(uj[b]);
                    const Vec<V> f = ph.flux_x(uj[b]);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc.a[v] = acc.a[v] + w[b] * (-f.a[v]);
                }
                if constexpr (VISC != VISC_NONE) {
#pragma unroll
                    for (int v = 0; v < V; ++v) gx.a[v] = gx.a[v] + w[b] * uj[b].a[v];
                }
            }
        }
        if (two_phase && width > 0) {
            // every lane is done with wx: refill the weight buffer with wy (already pulled into L2 by the prefetch)
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                const unsigned char *src = A.op.base + (size_t)A.op.off[slice] * kColBytes2;
                mbar_expect_tx(&bars[warp], (uint32_t)width * kSlice * 8);
                bulk_g2s(smem_dyn + (size_t)warp * A.buf_bytes + (size_t)width * kSlice * 4, src + (size_t)width * kSlice * 12,
                         (uint32_t)width * kSlice * 8, &bars[warp]);
            }
            mbar_wait(&bars[warp], 1);
        }
        for (int c0 = 0; c0 < width; c0 += kBatch) {
            Vec<V> uj[kBatch];
            double w[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const bool ok = c0 + b < width;
                const int cc = ok ? c0 + b : width - 1;
                const int j = ok ? ip[cc * kSlice] : A.dummy;
                const double wv = wyp[cc * kSlice];
                w[b] = ok ? wv : 0.0;
                uj[b] = ld_ro(u + j);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                if constexpr (DO_FLUX) {
                    ph.prepare(uj[b]);
                    const Vec<V> f = ph.flux_y(uj[b]);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc.a[v] = acc.a[v] + w[b] * (-f.a[v]);
                }
                if constexpr (VISC != VISC_NONE) {
#pragma unroll
                    for (int v = 0; v < V; ++v) gy.a[v] = gy.a[v] + w[b] * uj[b].a[v];
                }
            }
        }
    } else {
        // single sweep, FMA accumulation (not the reference's rounding sequence; agrees to ~1e-13 normwise)
        for (int c0 = 0; c0 < width; c0 += kBatch) {
            Vec<V> uj[kBatch];
            double wa[kBatch], wb[kBatch];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const bool ok = c0 + b < width;
                const int cc = ok ? c0 + b : width - 1;
                const int j = ok ? ip[cc * kSlice] : A.dummy;
                const double w1 = wxp[cc * kSlice], w2 = wyp[cc * kSlice];
                wa[b] = ok ? w1 : 0.0;
                wb[b] = ok ? w2 : 0.0;
                uj[b] = ld_ro(u + j);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                if constexpr (DO_FLUX) {
                    ph.prepare(uj[b]);
                    const Vec<V> f = ph.flux_x(uj[b]);
                    const Vec<V> h = ph.flux_y(uj[b]);
#pragma unroll
                    for (int v = 0; v < V; ++v) acc.a[v] = fma(-wb[b], h.a[v], fma(-wa[b], f.a[v], acc.a[v]));
                }
                if constexpr (VISC != VISC_NONE) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        gx.a[v] = fma(wa[b], uj[b].a[v], gx.a[v]);
                        gy.a[v] = fma(wb[b], uj[b].a[v], gy.a[v]);
                    }
                }
            }
        }
    }

    if (!live) return;
    pass_a_epilogue<V, EQ, DO_FLUX, VISC>(A, row, acc, gx, gy, ui, ad);
}

// ---- pass A over ROW PAIRS (same idea as k_pass_b_pair): one thread owns rows (2l, 2l+1) and walks the union of the
// two stencils, so a neighbour state shared by both rows is gathered -- and its flux evaluated -- once per sweep.
// Exact order: sweep 1 = all Dx terms in ascending column order, sweep 2 = all Dy terms; zero weights add exact zeros.
template <int V, int EQ, bool EXACT, int VISC>
__global__ void __launch_bounds__(128, 3) k_pass_a_pair(const PassAArgs A)
{
    static_assert(V == 4 && EQ == EQ_EULER2D, "pair kernel is instantiated for Euler 2-D");
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * 4 + warp;
    if (slice >= A.n_slices) return;
    const int64_t rowA = slice * (2 * kSlice) + 2 * lane, rowB = rowA + 1;
    const bool liveA = rowA < A.n_rows, liveB = rowB < A.n_rows;
    const int off = A.op.off[slice];
    const int width = A.op.off[slice + 1] - off;
    const unsigned char *src = A.op.base + (size_t)off * kColBytesPair;
    unsigned char *buf = smem_dyn + (size_t)warp * A.buf_bytes;
    if (lane == 0) mbar_init(&bars[warp], 1);
    __syncwarp();
    if (lane == 0 && width > 0) {
        mbar_expect_tx(&bars[warp], (uint32_t)width * kSlice * 4);
        bulk_g2s(buf, src, (uint32_t)width * kSlice * 4, &bars[warp]);
    }
    const int *ip = reinterpret_cast<const int *>(buf) + lane;
    const double *wbase = reinterpret_cast<const double *>(src + (size_t)width * kSlice * 4) + lane;
    const size_t wstride = (size_t)width * kSlice;
    const Vec<V> *__restrict__ u = reinterpret_cast<const Vec<V> *>(A.u);
    Physics<EQ, V> ph;
    ph.gamma = A.eqp0;

    Vec<V> accA, accB, gxA, gxB, gyA, gyB, uiA, uiB, adA, adB;
#pragma unroll
    for (int v = 0; v < V; ++v)
        accA.a[v] = accB.a[v] = gxA.a[v] = gxB.a[v] = gyA.a[v] = gyB.a[v] = uiA.a[v] = uiB.a[v] = adA.a[v] = adB.a[v] = 0.0;
    if (A.accumulate) {
        if (liveA) accA = reinterpret_cast<const Vec<V> *>(A.du)[rowA];
        if (liveB) accB = reinterpret_cast<const Vec<V> *>(A.du)[rowB];
    }
    if constexpr (VISC != VISC_NONE) {
        if (liveA) uiA = ld_ro(u + rowA);
        if (liveB) uiB = ld_ro(u + rowB);
        if constexpr (VISC == VISC_RESIDUAL) {
            if (liveA) adA = ld_ro(reinterpret_cast<const Vec<V> *>(A.approx_du) + rowA);
            if (liveB) adB = ld_ro(reinterpret_cast<const Vec<V> *>(A.approx_du) + rowB);
        }
    }
    if (width > 0) mbar_wait(&bars[warp], 0);

    constexpr int kBatch = 4;
    if constexpr (EXACT) {
        auto sweep = [&](auto dir_tag) {
            constexpr int DIR = decltype(dir_tag)::value;
            const double *wA = wbase + (size_t)DIR * wstride;        // wxA | wyA
            const double *wB = wbase + (size_t)(2 + DIR) * wstride;  // wxB | wyB
            for (int c0 = 0; c0 < width; c0 += kBatch) {
                Vec<V> uj[kBatch];
                double wa[kBatch], wb[kBatch];
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const bool ok = c0 + b < width;
                    const int cc = ok ? c0 + b : width - 1;
                    const int j = ok ? ip[cc * kSlice] : A.dummy;
                    const double w1 = ld_stream(wA + (size_t)cc * kSlice), w2 = ld_stream(wB + (size_t)cc * kSlice);
                    wa[b] = ok ? w1 : 0.0;
                    wb[b] = ok ? w2 : 0.0;
                    uj[b] = ld_ro(u + j);
                }
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    ph.prepare(uj[b]);
                    Vec<V> f;
                    if constexpr (DIR == 0) f = ph.flux_x(uj[b]);
                    else f = ph.flux_y(uj[b]);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        accA.a[v] = accA.a[v] + wa[b] * (-f.a[v]);
                        accB.a[v] = accB.a[v] + wb[b] * (-f.a[v]);
                    }
                    if constexpr (VISC != VISC_NONE) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if constexpr (DIR == 0) {
                                gxA.a[v] = gxA.a[v] + wa[b] * uj[b].a[v];
                                gxB.a[v] = gxB.a[v] + wb[b] * uj[b].a[v];
                            } else {
                                gyA.a[v] = gyA.a[v] + wa[b] * uj[b].a[v];
                                gyB.a[v] = gyB.a[v] + wb[b] * uj[b].a[v];
                            }
                        }
                    }
                }
            }
        };
        sweep(std::integral_constant<int, 0>{});
        sweep(std::integral_constant<int, 1>{});
    } else {
        for (int c0 = 0; c0 < width; c0 += kBatch) {
            Vec<V> uj[kBatch];
            double w[kBatch][4];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const bool ok = c0 + b < width;
                const int cc = ok ? c0 + b : width - 1;
                const int j = ok ? ip[cc * kSlice] : A.dummy;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double wv = ld_stream(wbase + q * wstride + (size_t)cc * kSlice);
                    w[b][q] = ok ? wv : 0.0;
                }
                uj[b] = ld_ro(u + j);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                ph.prepare(uj[b]);
                const Vec<V> f = ph.flux_x(uj[b]);
                const Vec<V> h = ph.flux_y(uj[b]);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    accA.a[v] = fma(-w[b][1], h.a[v], fma(-w[b][0], f.a[v], accA.a[v]));
                    accB.a[v] = fma(-w[b][3], h.a[v], fma(-w[b][2], f.a[v], accB.a[v]));
                }
                if constexpr (VISC != VISC_NONE) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        gxA.a[v] = fma(w[b][0], uj[b].a[v], gxA.a[v]);
                        gyA.a[v] = fma(w[b][1], uj[b].a[v], gyA.a[v]);
                        gxB.a[v] = fma(w[b][2], uj[b].a[v], gxB.a[v]);
                        gyB.a[v] = fma(w[b][3], uj[b].a[v], gyB.a[v]);
                    }
                }
            }
        }
    }
    if (liveA) pass_a_epilogue<V, EQ, true, VISC>(A, rowA, accA, gxA, gyA, uiA, adA);
    if (liveB) pass_a_epilogue<V, EQ, true, VISC>(A, rowB, accB, gxB, gyB, uiB, adB);
}

// ---- pass B: du -= Dx' gX + Dy' gY over the transposed sliced-ELL operator --------------------------------
// adjoint mul! (call sites hyperviscosity.jl:376-377, 405-406):
//   tmp = sum_j A[j,i] * x[j] (ascending j), then C[i] += tmp * (-1); x-direction first, then y.
struct PassBArgs {
    EllBlob opT;
    const void *g;
    void *du;
    int64_t n_rows;
    int64_t n_slices;
    int buf_bytes;
    int pf_dist;
    int dummy;
};

template <int V, bool EXACT, bool STAGE_W>
__global__ void __launch_bounds__(128, 4) k_pass_b(const PassBArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * 4 + warp;
    if (slice >= A.n_slices) return;
    const int64_t row = slice * kSlice + lane;
    const bool live = row < A.n_rows;
    const SliceView sv = stage_slice<kColBytes2, STAGE_W>(A.opT, slice, A.n_slices, smem_dyn + (size_t)warp * A.buf_bytes,
                                                         &bars[warp], lane, A.pf_dist);
    const int width = sv.width;
    const int *ip = sv.ip;
    const double *wxp = sv.wxp;
    const double *wyp = sv.wyp;
    const Vec<2 * V> *__restrict__ g = reinterpret_cast<const Vec<2 * V> *>(A.g);

    Vec<V> tx, ty;
#pragma unroll
    for (int v = 0; v < V; ++v) tx.a[v] = ty.a[v] = 0.0;
    Vec<V> d;
#pragma unroll
    for (int v = 0; v < V; ++v) d.a[v] = 0.0;
    if (live) d = reinterpret_cast<const Vec<V> *>(A.du)[row];
    if (width > 0) mbar_wait(&bars[warp], 0);
    constexpr int kBatch = 4;
    for (int c0 = 0; c0 < width; c0 += kBatch) {
        Vec<2 * V> gj[kBatch];
        double wa[kBatch], wb[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const bool ok = c0 + b < width;
            const int cc = ok ? c0 + b : width - 1;
            const int j = ok ? ip[cc * kSlice] : A.dummy;
            double w1, w2;
            if constexpr (STAGE_W) {
                w1 = wxp[cc * kSlice];
                w2 = wyp[cc * kSlice];
            } else {
                w1 = ld_stream(wxp + cc * kSlice);
                w2 = ld_stream(wyp + cc * kSlice);
            }
            wa[b] = ok ? w1 : 0.0;
            wb[b] = ok ? w2 : 0.0;
            gj[b] = ld_ro(g + j);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if constexpr (EXACT) {
                    tx.a[v] = tx.a[v] + wa[b] * gj[b].a[v];
                    ty.a[v] = ty.a[v] + wb[b] * gj[b].a[V + v];
                } else {
                    tx.a[v] = fma(wa[b], gj[b].a[v], tx.a[v]);
                    ty.a[v] = fma(wb[b], gj[b].a[V + v], ty.a[v]);
                }
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int v = 0; v < V; ++v) d.a[v] = (d.a[v] + tx.a[v] * -1.0) + ty.a[v] * -1.0;
    st_vec(reinterpret_cast<Vec<V> *>(A.du) + row, d);
}

// ---- pass B over ROW PAIRS: one thread owns two consecutive rows and walks the UNION of their D' rows ----------------
// The gather is the bottleneck (about one 32-byte sector per cycle per SM, almost no sharing between the lanes of a
// request), and two rows that are neighbours along the curve share ~15 of 20 entries: walking the union (~25 entries,
// a zero weight where a row does not have the entry) fetches every shared g record once for both rows -> ~1/3 fewer
// gathered sectors.  Entries stay in the reference's order and a zero weight adds an exact zero, so the sums are
// unchanged bit for bit (for finite g).  Blob per pair-slice (32 lanes x 2 rows):
//   [ idx : w x 32 int32 ][ wxA ][ wyA ][ wxB ][ wyB ]   (each weight block w x 32 doubles)

struct PassBPairArgs {
    EllBlob opT;       // pair-slice blobs
    const void *g;
    void *du;
    int64_t n_rows;
    int64_t n_slices;  // pair-slices
    int buf_bytes;     // per-warp shared memory for the index block
    int dummy;
};

template <int V, bool EXACT>
__global__ void __launch_bounds__(128, 3) k_pass_b_pair(const PassBPairArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * 4 + warp;
    if (slice >= A.n_slices) return;
    const int64_t rowA = slice * (2 * kSlice) + 2 * lane, rowB = rowA + 1;
    const bool liveA = rowA < A.n_rows, liveB = rowB < A.n_rows;
    const int off = A.opT.off[slice];
    const int width = A.opT.off[slice + 1] - off;
    const unsigned char *src = A.opT.base + (size_t)off * kColBytesPair;
    unsigned char *buf = smem_dyn + (size_t)warp * A.buf_bytes;
    if (lane == 0) mbar_init(&bars[warp], 1);
    __syncwarp();
    if (lane == 0 && width > 0) {
        mbar_expect_tx(&bars[warp], (uint32_t)width * kSlice * 4);
        bulk_g2s(buf, src, (uint32_t)width * kSlice * 4, &bars[warp]);
    }
    const int *ip = reinterpret_cast<const int *>(buf) + lane;
    const double *wbase = reinterpret_cast<const double *>(src + (size_t)width * kSlice * 4) + lane;
    const size_t wstride = (size_t)width * kSlice;  // doubles per weight block
    const Vec<2 * V> *__restrict__ g = reinterpret_cast<const Vec<2 * V> *>(A.g);

    Vec<V> txA, tyA, txB, tyB, dA, dB;
#pragma unroll
    for (int v = 0; v < V; ++v) txA.a[v] = tyA.a[v] = txB.a[v] = tyB.a[v] = dA.a[v] = dB.a[v] = 0.0;
    if (liveA) dA = reinterpret_cast<const Vec<V> *>(A.du)[rowA];
    if (liveB) dB = reinterpret_cast<const Vec<V> *>(A.du)[rowB];
    if (width > 0) mbar_wait(&bars[warp], 0);

    constexpr int kBatch = 4;
    for (int c0 = 0; c0 < width; c0 += kBatch) {
        Vec<2 * V> gj[kBatch];
        double w[kBatch][4];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const bool ok = c0 + b < width;
            const int cc = ok ? c0 + b : width - 1;
            const int j = ok ? ip[cc * kSlice] : A.dummy;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double wv = ld_stream(wbase + q * wstride + (size_t)cc * kSlice);
                w[b][q] = ok ? wv : 0.0;
            }
            gj[b] = ld_ro(g + j);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if constexpr (EXACT) {
                    txA.a[v] = txA.a[v] + w[b][0] * gj[b].a[v];
                    tyA.a[v] = tyA.a[v] + w[b][1] * gj[b].a[V + v];
                    txB.a[v] = txB.a[v] + w[b][2] * gj[b].a[v];
                    tyB.a[v] = tyB.a[v] + w[b][3] * gj[b].a[V + v];
                } else {
                    txA.a[v] = fma(w[b][0], gj[b].a[v], txA.a[v]);
                    tyA.a[v] = fma(w[b][1], gj[b].a[V + v], tyA.a[v]);
                    txB.a[v] = fma(w[b][2], gj[b].a[v], txB.a[v]);
                    tyB.a[v] = fma(w[b][3], gj[b].a[V + v], tyB.a[v]);
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        dA.a[v] = (dA.a[v] + txA.a[v] * -1.0) + tyA.a[v] * -1.0;
        dB.a[v] = (dB.a[v] + txB.a[v] * -1.0) + tyB.a[v] * -1.0;
    }
    if (liveA) st_vec(reinterpret_cast<Vec<V> *>(A.du) + rowA, dA);
    if (liveB) st_vec(reinterpret_cast<Vec<V> *>(A.du) + rowB, dB);
}

// ---- generic single-matrix source: du += alpha * H u  (hyperviscosity, hyperviscosity.jl:52-64,121-134) ----
struct SpmvArgs {
    EllBlob op;
    const void *x;
    void *y;
    int64_t n_rows;
    int64_t n_slices;
    int buf_bytes;
    int pf_dist;
    int dummy;
    double alpha;
};

template <int V, bool EXACT, bool STAGE_W>
__global__ void __launch_bounds__(128, 4) k_spmv_accum(const SpmvArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * 4 + warp;
    if (slice >= A.n_slices) return;
    const int64_t row = slice * kSlice + lane;
    const bool live = row < A.n_rows;
    const SliceView sv = stage_slice<kColBytes1, STAGE_W>(A.op, slice, A.n_slices, smem_dyn + (size_t)warp * A.buf_bytes,
                                                         &bars[warp], lane, A.pf_dist);
    const int width = sv.width;
    const int *ip = sv.ip;
    const double *wp = sv.wxp;
    const Vec<V> *__restrict__ x = reinterpret_cast<const Vec<V> *>(A.x);
    Vec<V> *yp = reinterpret_cast<Vec<V> *>(A.y) + row;
    Vec<V> acc;
#pragma unroll
    for (int v = 0; v < V; ++v) acc.a[v] = 0.0;
    if (live) acc = *yp;
    if (width > 0) mbar_wait(&bars[warp], 0);
    constexpr int kBatch = 6;
    for (int c0 = 0; c0 < width; c0 += kBatch) {
        Vec<V> xj[kBatch];
        double w[kBatch];
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
            const bool ok = c0 + b < width;
            const int cc = ok ? c0 + b : width - 1;
            const int j = ok ? ip[cc * kSlice] : A.dummy;
            const double wv = wp[cc * kSlice];
            w[b] = ok ? wv : 0.0;
            xj[b] = ld_ro(x + j);
        }
#pragma unroll
        for (int b = 0; b < kBatch; ++b) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if constexpr (EXACT) {
                    acc.a[v] = acc.a[v] + w[b] * (xj[b].a[v] * A.alpha);
                } else {
                    acc.a[v] = fma(w[b], xj[b].a[v] * A.alpha, acc.a[v]);
                }
            }
        }
    }
    if (live) st_vec(yp, acc);
}

// ---- strong boundary conditions (calc_single_boundary_flux!, rbfsolver.jl:288-318) -------------------------
struct BcArgs {
    int kind;
    int64_t nb;
    const int *idx;         // device rows
    const double *normals;  // nb x 2
    const double *values;   // nb x V (AoS)
    void *u;
    void *du;  // nullable: skip du writes (du not formed yet)
};

template <int V>
__global__ void k_boundary(const BcArgs A)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.nb) return;
    const int b = A.idx[j];
    Vec<V> *u = reinterpret_cast<Vec<V> *>(A.u) + b;
    Vec<V> *du = A.du ? reinterpret_cast<Vec<V> *>(A.du) + b : nullptr;
    if (A.kind == 0) {  // Dirichlet  PointCloudBCs.jl:49-63
        Vec<V> val;
#pragma unroll
        for (int v = 0; v < V; ++v) val.a[v] = A.values[j * V + v];
        *u = val;
        if (du) {
#pragma unroll
            for (int v = 0; v < V; ++v) du->a[v] = 0.0;
        }
    } else if (A.kind == 1) {  // slip wall  PointCloudBCs.jl:15-21, 87-106
        if constexpr (V == 4) {
            const double nx0 = A.normals[2 * j], ny0 = A.normals[2 * j + 1];
            const double nrm = sqrt(nx0 * nx0 + ny0 * ny0);
            const double nx = nx0 / nrm, ny = ny0 / nrm;
            const double m1 = u->a[1], m2 = u->a[2];
            const double vdotn = m1 * nx + m2 * ny;
            u->a[1] = m1 - vdotn * nx;
            u->a[2] = m2 - vdotn * ny;
            if (du) {
                du->a[1] = 0.0;
                du->a[2] = 0.0;
            }
        }
    }
}

// ---- reductions for update_residual_visc! (ode_mean / ode_maximum, src/auxiliary/mpi.jl:40-81) --------------
// Single launch each: every block writes its partial, the last block to finish (atomic ticket) combines the partials
// in a fixed order -> deterministic for a fixed grid size.
template <int V>
__global__ void __launch_bounds__(256) k_sum_mean(const Vec<V> *__restrict__ u, int64_t n, double *partial,
                                                unsigned int *ticket, double divisor, double *stats)
{
    double s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) s[v] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vec<V> x = ld_ro(u + i);
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] += x.a[v];
    }
    __shared__ double sh[8][V];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int v = 0; v < V; ++v)
        for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_down_sync(0xffffffffu, s[v], o);
    if (l == 0)
        for (int v = 0; v < V; ++v) sh[w][v] = s[v];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int v = 0; v < V; ++v) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += sh[k][v];
            partial[(int64_t)blockIdx.x * V + v] = t;
        }
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int v = 0; v < V; ++v) s[v] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
        for (int v = 0; v < V; ++v) s[v] += __ldcg(&partial[(int64_t)b * V + v]);
#pragma unroll
    for (int v = 0; v < V; ++v)
        for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_down_sync(0xffffffffu, s[v], o);
    __syncthreads();
    if (l == 0)
        for (int v = 0; v < V; ++v) sh[w][v] = s[v];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int v = 0; v < V; ++v) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += sh[k][v];
            stats[v] = t;
            stats[V + v] = t / divisor;
        }
        *ticket = 0;
    }
}

template <int V, bool LEX>
__global__ void __launch_bounds__(256) k_maxdev_norms(const Vec<V> *__restrict__ u, int64_t n, const double *sums,
                                                    int nparts, double divisor, double *partial, unsigned int *ticket,
                                                    double *norms, int replace_zero, double *mean_out)
{
    // mean = (sum of the per-rank sums, in rank order) / divisor  -- ode_mean, src/auxiliary/mpi.jl:40-52
    double m[V], best[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        double t = sums[v];
        for (int r = 1; r < nparts; ++r) t += sums[r * V + v];
        m[v] = t / divisor;
        best[v] = -1.0;
    }
    if (mean_out && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) mean_out[v] = m[v];
    }
    auto combine = [&](double *a, const double *b) {
        if constexpr (LEX) {
            if (lex_less<V>(a, b)) {
#pragma unroll
                for (int v = 0; v < V; ++v) a[v] = b[v];
            }
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) a[v] = jl_max(a[v], b[v]);
        }
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vec<V> x = ld_ro(u + i);
        double c[V];
#pragma unroll
        for (int v = 0; v < V; ++v) c[v] = fabs(x.a[v] - m[v]);
        combine(best, c);
    }
    __shared__ double sh[8][V];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    auto block_reduce = [&]() {
        for (int o = 16; o > 0; o >>= 1) {
            double other[V];
#pragma unroll
            for (int v = 0; v < V; ++v) other[v] = __shfl_down_sync(0xffffffffu, best[v], o);
            combine(best, other);
        }
        if (l == 0)
            for (int v = 0; v < V; ++v) sh[w][v] = best[v];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 1; k < 8; ++k) combine(best, sh[k]);
    };
    block_reduce();
    if (threadIdx.x == 0) {
        for (int v = 0; v < V; ++v) partial[(int64_t)blockIdx.x * V + v] = best[v];
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int v = 0; v < V; ++v) best[v] = -1.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        double c[V];
        for (int v = 0; v < V; ++v) c[v] = __ldcg(&partial[(int64_t)b * V + v]);
        combine(best, c);
    }
    __syncthreads();
    block_reduce();
    if (threadIdx.x == 0) {
        for (int v = 0; v < V; ++v) norms[v] = (replace_zero && best[v] == 0.0) ? kEps : best[v];
        *ticket = 0;
    }
}

// all boundary groups in one launch (used when no point belongs to two groups, so the group order cannot matter)
struct BcMergedArgs {
    int64_t nb;
    const int *kind;
    const int *idx;
    const double *normals;
    const double *values;
    void *u;
    void *du;
};
template <int V>
__global__ void k_boundary_merged(const BcMergedArgs A)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.nb) return;
    const int kind = A.kind[j];
    const int b = A.idx[j];
    Vec<V> *u = reinterpret_cast<Vec<V> *>(A.u) + b;
    Vec<V> *du = A.du ? reinterpret_cast<Vec<V> *>(A.du) + b : nullptr;
    if (kind == 0) {
        Vec<V> val;
#pragma unroll
        for (int v = 0; v < V; ++v) val.a[v] = A.values[j * V + v];
        *u = val;
        if (du) {
#pragma unroll
            for (int v = 0; v < V; ++v) du->a[v] = 0.0;
        }
    } else if (kind == 1) {
        if constexpr (V == 4) {
            const double nx0 = A.normals[2 * j], ny0 = A.normals[2 * j + 1];
            const double nrm = sqrt(nx0 * nx0 + ny0 * ny0);
            const double nx = nx0 / nrm, ny = ny0 / nrm;
            const double m1 = u->a[1], m2 = u->a[2];
            const double vdotn = m1 * nx + m2 * ny;
            u->a[1] = m1 - vdotn * nx;
            u->a[2] = m2 - vdotn * ny;
            if (du) {
                du->a[1] = 0.0;
                du->a[2] = 0.0;
            }
        }
    }
}

// ---- SSPRK33 stage update (stage formulas: DESIGN.md "time loop") --------------------
// stage 1 also snapshots uprev = u.  Flat over all doubles of the owned points.
__global__ void __launch_bounds__(256) k_ssprk33_stage(int stage, double dt, double *__restrict__ uprev,
                                                     const double *__restrict__ k, double *__restrict__ u,
                                                     int64_t len)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double dt2 = 2.0 * dt;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        if (stage == 1) {
            const double up = u[i];
            uprev[i] = up;
            u[i] = fma(dt, k[i], up);
        } else if (stage == 2) {
            u[i] = fma(dt, k[i], fma(3.0, uprev[i], u[i])) / 4.0;
        } else {
            u[i] = fma(dt2, k[i], fma(2.0, u[i], uprev[i])) / 3.0;
        }
    }
}

// (the Zhang-Shu positivity limiter kernels live in mft_limiter_kernels.cuh)

// ---- SSPRK43 stage updates + embedded error estimate (OrdinaryDiffEq low-storage SSPRK43; DESIGN.md "time loop") --------
// stage 1 also snapshots uprev; stage 3 forms utilde = (uprev + 2 u3)/3 and u = (2 uprev + u3)/3; stage 4 turns utilde into
// the error vector 0.5*(utilde - u).
__global__ void __launch_bounds__(256) k_ssprk43_stage(int stage, double dt, double *__restrict__ uprev,
                                                     const double *__restrict__ k, double *__restrict__ u,
                                                     double *__restrict__ utilde, int64_t len)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double h = dt / 2.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
        if (stage == 1) {
            const double up = u[i];
            uprev[i] = up;
            u[i] = fma(h, k[i], up);
        } else if (stage == 2) {
            u[i] = fma(h, k[i], u[i]);
        } else if (stage == 3) {
            const double u3 = fma(h, k[i], u[i]);
            const double up = uprev[i];
            utilde[i] = fma(2.0, u3, up) / 3.0;
            u[i] = fma(2.0, up, u3) / 3.0;
        } else {
            const double un = fma(h, k[i], u[i]);
            u[i] = un;
            utilde[i] = 0.5 * (utilde[i] - un);
        }
    }
}

// sum_i (utilde_i / (abstol + max(|uprev_i|,|u_i|) reltol))^2 over the owned entries; deterministic (last block finishes)
__global__ void __launch_bounds__(256) k_error_sumsq(const double *__restrict__ utilde, const double *__restrict__ uprev,
                                                   const double *__restrict__ u, int64_t len, double abstol, double reltol,
                                                   double *partial, unsigned int *ticket, double *out)
{
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
        const double a = fabs(uprev[i]), b = fabs(u[i]);
        const double r = utilde[i] / (abstol + (a > b ? a : b) * reltol);
        s += r * r;
    }
    __shared__ double sh[8];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (l == 0) sh[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[k];
        partial[blockIdx.x] = t;
        __threadfence();
        is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(&partial[b]);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __syncthreads();
    if (l == 0) sh[w] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += sh[k];
        *out = t;
        *ticket = 0;
    }
}

// ---- time-history residual: approx_du = sum_s w[s] * hist[s]  (update_approx_du!, history.jl:113-129) --------
struct ApproxDuArgs {
    const double *hist[8];
    double w[8];
    int nterms;
    double *out;
    int64_t len;
};
__global__ void __launch_bounds__(256) k_approx_du(const ApproxDuArgs A)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < A.len; i += stride) {
        double acc = 0.0;
        for (int s = 0; s < A.nterms; ++s) acc = fma(A.w[s], A.hist[s][i], acc);
        A.out[i] = acc;
    }
}

// ---- SoA (caller order) <-> AoS (device order) -------------------------------------------------------------
template <int V>
__global__ void k_pack(const double *__restrict__ soa, int64_t ld, const int *__restrict__ perm, Vec<V> *out, int64_t n)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const int64_t src = perm ? perm[d] : d;
    Vec<V> v;
#pragma unroll
    for (int k = 0; k < V; ++k) v.a[k] = soa[(int64_t)k * ld + src];
    out[d] = v;
}
template <int V>
__global__ void k_unpack(const Vec<V> *__restrict__ in, const int *__restrict__ perm, double *soa, int64_t ld, int64_t n)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    const int64_t dst = perm ? perm[d] : d;
    const Vec<V> v = in[d];
#pragma unroll
    for (int k = 0; k < V; ++k) soa[(int64_t)k * ld + dst] = v.a[k];
}
// compact gather of a few rows (boundary + halo points) into an SoA buffer: buf[v*m + i] = field[rows[i]].a[v]
template <int V>
__global__ void k_gather_rows(const Vec<V> *__restrict__ field, const int *__restrict__ rows, double *buf, int64_t m)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const Vec<V> v = field[rows[i]];
#pragma unroll
    for (int k = 0; k < V; ++k) buf[(int64_t)k * m + i] = v.a[k];
}
__global__ void k_unpack_scalar(const double *__restrict__ in, const int *__restrict__ perm, double *out, int64_t n)
{
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    out[perm ? perm[d] : d] = in[d];
}

// halo send-buffer gather (perform_halo_update! pack loop, src/auxiliary/mpi.jl:234-236)
template <int W>
__global__ void k_halo_pack(const Vec<W> *__restrict__ src, const int *__restrict__ send_rows, Vec<W> *buf, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    buf[i] = src[send_rows[i]];
}

// =====================================================================================================================
// Peer-memory (NVLink) exchange: one process per GPU, peers' buffers mapped with CUDA IPC.  A sender writes its halo
// block straight into the receiver's halo tail and then raises an epoch flag in the receiver's window; the per-rank
// partial norms travel the same way.  Only kernels are involved, so a whole multi-GPU SSPRK step is graph-capturable,
// and the latency is one NVLink write instead of an NCCL launch.  Flow control: a rank re-uses a peer's halo tail only
// after that peer has signalled (credit flag) that it consumed the previous epoch.
// =====================================================================================================================
constexpr int kMaxRanks = 16;
constexpr int kRecDoubles = 128; // one norm record (mft_fused_kernels.cuh): sums + two levels of lexicographic leaves, 1 KB
constexpr unsigned long long kSpinLimit = 1ull << 31;  // ~ seconds: then give up and raise the error flag

struct P2PWindow {                                 // lives in every rank's memory; peers write into it
    unsigned long long data_flag[2][kMaxRanks];    // [field][src rank]: epoch whose halo block has fully arrived
    unsigned long long credit[2][kMaxRanks];       // [field][dst rank]: epoch of `field` that dst has consumed
    unsigned long long sum_flag[2][kMaxRanks];     // [parity][rank]
    unsigned long long max_flag[2][kMaxRanks];
    double sums[2][kMaxRanks][4];
    double maxs[2][kMaxRanks][4];
    // fused step (mft_fused_kernels.cuh): one norm record per rank and parity, and its epoch flag
    unsigned long long rec_flag[2][kMaxRanks];
    double rec[2][kMaxRanks][kRecDoubles];
};

struct P2PLocal {                                  // local counters (never written remotely)
    unsigned long long epoch[2];                   // halo epochs of field 0 (u) and 1 (g)
    unsigned long long epoch_n;                    // norms epoch
    unsigned int ticket[4];
    int error;
    unsigned long long norm_ready;                 // fused step: norms epoch whose merged norms are in stats[] (set by block 0 of pass A)
    unsigned int ticket_g;                         // fused step: band blocks of pass A that have finished their g puts
};

struct P2PPeers {
    P2PWindow *win[kMaxRanks];   // windows of all ranks (own included)
    void *field[2][kMaxRanks];   // peers' u and g arrays
    int nranks, rank;
    int ndst, nsrc;              // ranks I send to / receive from
    int dst[kMaxRanks], src[kMaxRanks];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// flag store AFTER one __threadfence_system(): a release store per flag would run a system-scope fence each (one NVLink round
// trip per flag, serialised: measured as a ~30 us tail on 8 GPUs); one fence orders the data before all the flags
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long want, int *error)
{
    unsigned long long n = 0;
    while (ld_acquire_sys(flag) < want) {
        if (++n > kSpinLimit) {
            *error = 1;
            return false;
        }
        __nanosleep(20);
    }
    return true;
}

// put: every send entry goes straight into the destination rank's halo tail.  F = field id (0: u, 1: g).
template <int W>
__global__ void __launch_bounds__(256) k_p2p_put(P2PPeers P, P2PLocal *L, int F, const Vec<W> *__restrict__ src_field,
                                               const int *__restrict__ send_rows, const int *__restrict__ send_peer,
                                               const long long *__restrict__ send_dst, int64_t n_send)
{
    __shared__ bool is_last;
    const unsigned long long e = L->epoch[F] + 1;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            // credits: everything stream-ordered before this kernel has consumed the halos of both fields' last epochs
            for (int f = 0; f < 2; ++f) {
                const unsigned long long ef = L->epoch[f];
                for (int i = 0; i < P.nsrc; ++i) st_release_sys(&P.win[P.src[i]]->credit[f][P.rank], ef);
            }
        }
        // wait until every destination has consumed epoch e-1 of this field
        for (int i = 0; i < P.ndst; ++i) spin_until(&P.win[P.rank]->credit[F][P.dst[i]], e - 1, &L->error);
    }
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_send; i += (int64_t)gridDim.x * blockDim.x) {
        const Vec<W> v = src_field[send_rows[i]];
        Vec<W> *dst = reinterpret_cast<Vec<W> *>(P.field[F][send_peer[i]]) + send_dst[i];
        st_vec(dst, v);
    }
    // fences are cumulative: the block barrier makes every thread's remote stores visible to thread 0, whose
    // system-scope fence then orders them before the ticket (and, in the last block, before the flags)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        is_last = atomicAdd(&L->ticket[F], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence_system();
        for (int i = 0; i < P.ndst; ++i) st_release_sys(&P.win[P.dst[i]]->data_flag[F][P.rank], e);
        L->epoch[F] = e;
        L->ticket[F] = 0;
    }
}

// wait until the halo blocks of the current epoch of field F have arrived from every source rank
__global__ void k_p2p_wait(P2PPeers P, P2PLocal *L, int F)
{
    if (threadIdx.x < P.nsrc) spin_until(&P.win[P.rank]->data_flag[F][P.src[threadIdx.x]], L->epoch[F], &L->error);
}

// publish V doubles (this rank's partial sum or max-deviation candidate) into every rank's window
__device__ __forceinline__ void p2p_publish(const P2PPeers &P, int which, int par, unsigned long long en, const double *val)
{
    for (int r = 0; r < P.nranks; ++r) {
        P2PWindow *w = P.win[r];
        double *dst = which == 0 ? w->sums[par][P.rank] : w->maxs[par][P.rank];
        for (int v = 0; v < 4; ++v) dst[v] = val[v];
    }
    __threadfence_system();
    for (int r = 0; r < P.nranks; ++r) {
        P2PWindow *w = P.win[r];
        st_release_sys(which == 0 ? &w->sum_flag[par][P.rank] : &w->max_flag[par][P.rank], en);
    }
}
// local sum of the owned points -> published to all ranks (epoch en = epoch_n + 1)
__global__ void __launch_bounds__(256) k_p2p_sum(const Vec<4> *__restrict__ u, int64_t n, double *partial, P2PPeers P, P2PLocal *L)
{
    constexpr int V = 4;
    double s[V];
#pragma unroll
    for (int v = 0; v < V; ++v) s[v] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vec<V> x = ld_ro(u + i);
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] += x.a[v];
    }
    __shared__ double sh[8][V];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int v = 0; v < V; ++v)
        for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_down_sync(0xffffffffu, s[v], o);
    if (l == 0)
        for (int v = 0; v < V; ++v) sh[w][v] = s[v];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int v = 0; v < V; ++v) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += sh[k][v];
            partial[(int64_t)blockIdx.x * V + v] = t;
        }
        __threadfence();
        is_last = atomicAdd(&L->ticket[2], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int v = 0; v < V; ++v) s[v] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
        for (int v = 0; v < V; ++v) s[v] += __ldcg(&partial[(int64_t)b * V + v]);
#pragma unroll
    for (int v = 0; v < V; ++v)
        for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_down_sync(0xffffffffu, s[v], o);
    __syncthreads();
    if (l == 0)
        for (int v = 0; v < V; ++v) sh[w][v] = s[v];
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot[V];
        for (int v = 0; v < V; ++v) {
            double t = 0.0;
            for (int k = 0; k < 8; ++k) t += sh[k][v];
            tot[v] = t;
        }
        const unsigned long long en = L->epoch_n + 1;
        p2p_publish(P, 0, (int)(en & 1), en, tot);
        L->ticket[2] = 0;
    }
}

// max deviation from the GLOBAL mean (sums of all ranks, combined in rank order) -> candidates published to all ranks
template <bool LEX>
__global__ void __launch_bounds__(256) k_p2p_maxdev(const Vec<4> *__restrict__ u, int64_t n, double divisor, double *partial,
                                                  P2PPeers P, P2PLocal *L, double *mean_out)
{
    constexpr int V = 4;
    // the sums of all ranks have arrived: k_p2p_wait_norms<<<>>> runs right before this kernel (one polling thread
    // per flag instead of a system-scope acquire in every block)
    const unsigned long long en = L->epoch_n + 1;
    const int par = (int)(en & 1);
    double m[V], best[V];
    const P2PWindow *win = P.win[P.rank];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        double t = __ldcv(&win->sums[par][0][v]);
        for (int r = 1; r < P.nranks; ++r) t += __ldcv(&win->sums[par][r][v]);
        m[v] = t / divisor;
        best[v] = -1.0;
    }
    if (mean_out && blockIdx.x == 0 && threadIdx.x == 0)
        for (int v = 0; v < V; ++v) mean_out[v] = m[v];
    auto combine = [&](double *a, const double *b) {
        if constexpr (LEX) {
            if (lex_less<V>(a, b)) {
#pragma unroll
                for (int v = 0; v < V; ++v) a[v] = b[v];
            }
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v) a[v] = jl_max(a[v], b[v]);
        }
    };
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const Vec<V> x = ld_ro(u + i);
        double c[V];
#pragma unroll
        for (int v = 0; v < V; ++v) c[v] = fabs(x.a[v] - m[v]);
        combine(best, c);
    }
    __shared__ double sh[8][V];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    auto block_reduce = [&]() {
        for (int o = 16; o > 0; o >>= 1) {
            double other[V];
#pragma unroll
            for (int v = 0; v < V; ++v) other[v] = __shfl_down_sync(0xffffffffu, best[v], o);
            combine(best, other);
        }
        if (l == 0)
            for (int v = 0; v < V; ++v) sh[w][v] = best[v];
        __syncthreads();
        if (threadIdx.x == 0)
            for (int k = 1; k < 8; ++k) combine(best, sh[k]);
    };
    block_reduce();
    if (threadIdx.x == 0) {
        for (int v = 0; v < V; ++v) partial[(int64_t)blockIdx.x * V + v] = best[v];
        __threadfence();
        is_last = atomicAdd(&L->ticket[3], 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
#pragma unroll
    for (int v = 0; v < V; ++v) best[v] = -1.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        double c[V];
        for (int v = 0; v < V; ++v) c[v] = __ldcg(&partial[(int64_t)b * V + v]);
        combine(best, c);
    }
    __syncthreads();
    block_reduce();
    if (threadIdx.x == 0) {
        p2p_publish(P, 1, par, en, best);
        L->epoch_n = en;
        L->ticket[3] = 0;
    }
}

// one polling thread per rank flag: which = 0 waits for the partial sums of norms epoch epoch_n+1,
// which = 1 waits for the max-deviation candidates of epoch epoch_n and stages them in a local buffer for pass A
__global__ void k_p2p_wait_norms(P2PPeers P, P2PLocal *L, int which, double *out /* nranks x 4, which == 1 */)
{
    const unsigned long long en = which == 0 ? L->epoch_n + 1 : L->epoch_n;
    const int par = (int)(en & 1);
    const P2PWindow *win = P.win[P.rank];
    const int r = threadIdx.x;
    if (r < P.nranks) {
        spin_until(which == 0 ? &win->sum_flag[par][r] : &win->max_flag[par][r], en, &L->error);
        if (which == 1)
            for (int v = 0; v < 4; ++v) out[r * 4 + v] = __ldcv(&win->maxs[par][r][v]);
    }
}

}  // namespace mft
