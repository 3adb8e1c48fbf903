// mft_tile_kernels.cuh -- "union tile" variants of pass A / pass B for Euler 2-D (V = 4).
//
// The thread-per-row gathers of k_pass_a / k_pass_b go through the L1 data pipe at one wavefront per distinct 32-byte
// sector (~20 per 32-lane request; profiles/README.md).  A block of consecutive device rows is a compact patch of the
// cloud (Hilbert order) and the UNION of its stencils is only 1.4-1.9 points per row (tools/tile_sim.py).  So:
//
//   phase 1  the block loads the union ONCE with near-coalesced 256-bit loads (the union list is sorted), evaluates the
//            two IEEE divisions v = m / rho ONCE per union point, and stores the records in shared memory as 16-byte
//            units  A[s] = (rho, m1)  B[s] = (m2, E)  C[s] = (v1, v2)          (pass B:  A..D = gx01 gx23 gy01 gy23)
//   phase 2  thread per R rows; every neighbour access is 3 (4) LDS.128 at a 12-bit local slot; the flux is rebuilt
//            from (u, v1, v2) with the reference's operations, so sums are bit-identical to k_pass_a.
//
// Indices cost 2 bytes per entry instead of 4; the union lists add ~6 bytes per row and union point.
#pragma once
#include "mft_kernels.cuh"
#include "mft_fused_kernels.cuh"

namespace mft {

// warps (= 32-row slices) per tile.  Experiment knob: MFT_NVCC_EXTRA="-DMFT_TILE_WARPS=8" -> 256-row tiles (smaller union per
// row, twice the shared memory per block); the host layout builder and its selftest follow the constant.
#ifndef MFT_TILE_WARPS
#define MFT_TILE_WARPS 4
#endif
constexpr int kTileWarps = MFT_TILE_WARPS;

// minimum resident blocks per SM the compiler must allow for the R = 1 tile kernels (register cap = 65536 / (128 x blocks)).
// Experiment knobs: MFT_NVCC_EXTRA="-DMFT_TILE_OCC_A=6 -DMFT_TILE_OCC_B=5" python meshfreetrixi.jl_b200/build.py --force
#ifndef MFT_TILE_OCC_A
#define MFT_TILE_OCC_A 6
#endif
#ifndef MFT_TILE_OCC_B
#define MFT_TILE_OCC_B 5
#endif

// steps per batch of the R <= 2 main loops (loads of a batch are issued together, then consumed in order) and software pipelining
// of the operator fetch (1: the step words + weights of batch i+1 are requested before batch i is consumed)
#ifndef MFT_TILE_BATCH_A
#define MFT_TILE_BATCH_A 4
#endif
#ifndef MFT_TILE_BATCH_B
#define MFT_TILE_BATCH_B 6
#endif
#ifndef MFT_TILE_PIPE_A
#define MFT_TILE_PIPE_A 0
#endif
#ifndef MFT_TILE_PIPE_B
#define MFT_TILE_PIPE_B 0
#endif

__device__ __forceinline__ double2 lds2(const unsigned char *arr, uint32_t off)
{
    return *reinterpret_cast<const double2 *>(arr + off);
}

struct PassBTileArgs {
    const void *g;
    void *du;
    int64_t n_rows;
    int64_t n_slices;
    // several GPUs, fused step: band tiles (T.order, T.n_free) wait for the peers' g halo rows themselves
    const P2PPeers *P;   // nullable
    P2PLocal *L;
};

// 256-bit loads that go to L2 (rows another GPU writes while this kernel is running: never through the read-only path)
__device__ __forceinline__ Vec<4> ld_cg(const Vec<4> *p)
{
    Vec<4> v;
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a[0]), "=d"(v.a[1]), "=d"(v.a[2]), "=d"(v.a[3])
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ Vec<8> ld_cg(const Vec<8> *p)
{
    Vec<8> v;
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.a[0]), "=d"(v.a[1]), "=d"(v.a[2]), "=d"(v.a[3])
                 : "l"(p)
                 : "memory");
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4+32];"
                 : "=d"(v.a[4]), "=d"(v.a[5]), "=d"(v.a[6]), "=d"(v.a[7])
                 : "l"(p)
                 : "memory");
    return v;
}

// band tile: wait until the halo rows of field F (0: u, 1: g) of the current epoch have arrived from every source rank
__device__ __forceinline__ void band_wait(const P2PPeers *P, P2PLocal *L, int F)
{
    if ((int)threadIdx.x < P->nsrc) {
        // relaxed polls (no L1 invalidation per poll); the halo rows are then read with ld.global.cg, i.e. at L2, where the
        // peer's rows arrived before its flag did
        const unsigned long long *flag = &P->win[P->rank]->data_flag[F][P->src[threadIdx.x]];
        const unsigned long long want = L->epoch[F];
        unsigned long long have, n = 0;
        do {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(have) : "l"(flag) : "memory");
            if (have < want) {
                if (++n > kSpinLimit) {
                    L->error = 1;
                    break;
                }
                __nanosleep(100);
            }
        } while (have < want);
    }
    __syncthreads();
}

// =====================================================================================================================
// R rows per thread over union tiles.
//
// With one row per thread the kernels are bound by the shared-memory wavefront rate (ncu, profiles/r1_tile_*: l1tex
// data pipe 71 % busy, 1076 shared wavefronts per 32 rows, 1.9-way bank conflicts on the scattered LDS.128).  Two levers:
//  * rows that are consecutive along the curve share ~3/4 of their stencil, so a thread that owns R consecutive rows and
//    walks the UNION of their stencils reads every shared neighbour record once for all R rows: 13.5 (R=2) / 8.3 (R=4)
//    union steps per row instead of 20;
//  * the slot of a union point is ours to choose: the builder colours the points of a tile with the 8 bank groups so
//    that the points requested together by the 8 lanes of an LDS.128 phase fall into different groups (1.6-way instead
//    of 1.9 / 2.1 / 2.5-way for R = 1 / 2 / 4).  uslot[] maps the sorted union list to the slots; the last entry of every
//    tile's list is the dummy record (a finite state / zeros) that padding steps point at.
// A step carries a 16-bit word  slot | mask << 12  (mask bit r: row r has this entry); the weights stay COMPACT per
// row (no explicit zeros, HBM traffic = nnz): a row advances its own cursor when its mask bit is set.  Where a row
// lacks the entry it adds w = 0 times a finite value, i.e. an exact zero, so the sums are unchanged bit for bit.
//
// Slice = 32 lanes x R rows (device rows slice*32R + lane*R + r).  Blob of a slice, 16-byte aligned:
//   [ word : W x 32 x uint16 ][ wx_0 : L x 32 x f64 ] .. [ wx_{R-1} ][ wy_0 ] .. [ wy_{R-1} ]
// W = longest lane union, L = longest row of the slice.  Weights are streamed from global memory (coalesced up to the
// cursor skew between lanes); only the word block is staged in shared memory (bulk async copy).
struct TileROp {
    const unsigned char *base;
    const long long *boff;  // n_slices: byte offset of the slice blob
    const int *wl;          // n_slices x 2: W, L
    const int *uoff;        // n_tiles + 1 offsets into ulist / uslot
    const int *ulist;       // device indices of the tile's union, ascending; last entry of a tile = the dummy record
    const unsigned short *uslot;  // shared-memory slots of each entry: (copy 0, copy 1)
    int sstride;            // 16-byte units per shared array copy (> largest slot)
    int ncopy;              // 1, or 2: every record is stored twice under different bank assignments; bit 14 of a step
                            // word selects the copy (R <= 2), chosen by the builder to avoid bank conflicts in the phase
    int buf_bytes;          // per-warp staging buffer (word block)
    int pf_tiles;           // > 0: pull the operator data of tile blockIdx.x + pf_tiles into L2 (it is first touched about one
                            // block lifetime later, by then an L2 hit instead of a DRAM round trip on the critical path)
    // several GPUs, fused step: block b works on tile order[b]; the first n_free blocks touch neither halo columns nor send rows
    // and start at once, the others ("band" tiles, scheduled last) wait for the peers' halo rows while the interior runs
    const int *order;       // nullable: identity
    int n_free;
};

// STAGE_W (exact order only): the warp also stages its compact weight blocks in shared memory -- wx_0..wx_{R-1} with the
// step words (one bulk copy), wy_* over them after the x sweep.  A row's cursor then indexes a lane-private column of
// shared memory (bank = lane, conflict-free) instead of issuing a global load whose lanes sit at different cursors.
// L2 prefetch of a later tile's operator data: the offsets are loaded at kernel entry (tile_pf_begin) and used after the
// block barrier (tile_pf_issue), when they have long arrived -- no stall on the way.
struct TilePf {
    int u0, u1, W, L;
    long long boff;
    bool tile_ok, slice_ok;
};
__device__ __forceinline__ TilePf tile_pf_begin(const TileROp &T, int64_t n_slices, int warp)
{
    TilePf p{0, 0, 0, 0, 0, false, false};
    if (T.pf_tiles <= 0) return p;
    const int64_t bn = (int64_t)blockIdx.x + T.pf_tiles;
    p.tile_ok = bn < (int64_t)gridDim.x;
    if (p.tile_ok) {
        const int64_t tn = T.order ? (int64_t)__ldg(T.order + bn) : bn;
        p.u0 = __ldg(T.uoff + tn);
        p.u1 = __ldg(T.uoff + tn + 1);
        const int64_t sn = tn * kTileWarps + warp;
        p.slice_ok = sn < n_slices;
        if (p.slice_ok) {
            p.boff = __ldg(T.boff + sn);
            p.W = __ldg(T.wl + 2 * sn);
            p.L = __ldg(T.wl + 2 * sn + 1);
        }
    }
    return p;
}
template <int R>
__device__ __forceinline__ void tile_pf_issue(const TileROp &T, const TilePf &p, int warp, int lane)
{
    if (!p.tile_ok) return;
    if (warp == 0) {  // union list + slot table: 4 bytes per entry each, one 128-byte line per prefetch
        for (int i = lane * 32; i < p.u1 - p.u0; i += 32 * 32) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(T.ulist + p.u0 + i));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const unsigned int *>(T.uslot) + p.u0 + i));
        }
    }
    if (lane == 0 && p.slice_ok) {
        const uint32_t bytes = (uint32_t)p.W * kSlice * 2 + 2u * R * (uint32_t)p.L * kSlice * 8;
        if (bytes > 0) bulk_prefetch_l2(T.base + p.boff, bytes);
    }
}

template <int R, bool EXACT, bool DO_FLUX, int VISC, bool STAGE_W>
__global__ void __launch_bounds__(kTileWarps * 32, (R == 1 ? MFT_TILE_OCC_A : R == 2 ? 3 : 2)) k_pass_a_tiler(const PassAArgs A, const TileROp T)
{
    static_assert(EXACT || !STAGE_W, "staged weights: exact-order (two sweep) mode only");
    constexpr int V = 4;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[kTileWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = T.order ? __ldg(T.order + blockIdx.x) : (int)blockIdx.x;
    const bool band = T.order != nullptr && (int)blockIdx.x >= T.n_free;   // several GPUs: this tile reads halo rows / feeds peers
    const int64_t slice = (int64_t)tile * kTileWarps + warp;
    const bool has_slice = slice < A.n_slices;
    const int64_t row0 = (slice * kSlice + lane) * R;
    const uint32_t arr_bytes = (uint32_t)T.sstride * 16u, nc = (uint32_t)T.ncopy;
    unsigned char *sA = smem_dyn, *sB = smem_dyn + nc * arr_bytes, *sC = smem_dyn + 2 * nc * arr_bytes;
    unsigned char *buf = smem_dyn + 3 * nc * arr_bytes + (size_t)warp * T.buf_bytes;
    pdl_launch_dependents();   // (the successor's blocks only take SM slots this grid no longer needs)

    int W = 0, L = 0;
    const unsigned char *src = nullptr;
    if (has_slice) {
        W = T.wl[2 * slice];
        L = T.wl[2 * slice + 1];
        src = T.base + T.boff[slice];
        if (lane == 0) mbar_init(&bars[warp], 1);
        __syncwarp();
        if (lane == 0 && W > 0) {
            const uint32_t wbytes = (uint32_t)L * kSlice * 8 * R;  // one direction, all R rows
            const uint32_t bytes = (uint32_t)W * kSlice * 2 + (STAGE_W ? wbytes : 0u);
            mbar_expect_tx(&bars[warp], bytes);
            bulk_g2s(buf, src, bytes, &bars[warp]);
            if (wbytes > 0) bulk_prefetch_l2(src + (size_t)W * kSlice * 2 + (STAGE_W ? wbytes : 0u), wbytes);
        }
    }
    const TilePf pf = tile_pf_begin(T, A.n_slices, warp);
    const unsigned short *ip = reinterpret_cast<const unsigned short *>(buf) + lane;
    const double *wbase = reinterpret_cast<const double *>((STAGE_W ? buf : src) + (size_t)W * kSlice * 2) + lane;  // wx_0
    const size_t rstride = (size_t)L * kSlice;                                                    // doubles per weight block
    const Vec<V> *__restrict__ u = reinterpret_cast<const Vec<V> *>(A.u);

    const int u0 = T.uoff[tile];
    const int nu = T.uoff[tile + 1] - u0;
    // everything above touched operator data only (staged step words, prefetches, tile tables); from here on the kernel reads
    // what its predecessor wrote (u, the norm records)
    pdl_wait();
    if constexpr (VISC == VISC_RESIDUAL) {
        // fused step on several GPUs: the first block turns the ranks' norm records into the norms (they are needed in the
        // epilogues only, by then the records have long arrived)
        if (A.norm_merge == 1 && blockIdx.x == 0 && warp == 0) p2p_norms_merge(*A.P, A.L, A.divisor, A.norm_lex, const_cast<double *>(A.stats), lane);
        // (One GPU: the stage kernel's last block finishes the norms itself.  Moving that combine here -- block 0, while the other
        // blocks run their main loops -- was measured three ways in round 2 and lost every time: block 0 shares L2 with 887 block
        // prologues, its round trips take microseconds, and the whole first wave then waits in its epilogue: pass A +11 ... +29 us
        // per launch against 9 us saved in the stage kernel; profiles/README.md.)
    }

    // Several GPUs: the norms of this stage (mean | raw maximum | norms, 12 doubles) come from block 0 of this very launch.  A
    // warp must not pay L2 round trips for them in its epilogue (three dependent ones there cost +34 us per launch, measured):
    // the readiness flag is loaded HERE, looked at after phase 1 (long since returned), and if block 0 was done by then the 12
    // values are fetched, one per lane, behind the main loop; the epilogue only shuffles.  Only warps of the first wave can find
    // the flag unset; they poll in the epilogue.
    double nval = 0.0;
    bool nhave = false;
    unsigned long long nflag = 0, nwant = 0;
    if constexpr (VISC == VISC_RESIDUAL) {
        if (A.norm_merge) {
            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(nflag) : "l"(&A.L->norm_ready) : "memory");
            nwant = A.L->epoch_n;
        }
    }

    if (band) band_wait(A.P, A.L, 0);   // the peers' u rows of this stage are in the halo tail from here on
    for (int t = threadIdx.x; t < nu; t += kTileWarps * 32) {
        const int j = __ldg(T.ulist + u0 + t);
        const unsigned int sl2 = __ldg(reinterpret_cast<const unsigned int *>(T.uslot) + u0 + t);
        const int sl = (int)(sl2 & 0xffffu);
        const Vec<V> x = band ? ld_cg(u + j) : ld_ro(u + j);
        const double v1 = x.a[1] / x.a[0];
        const double v2 = x.a[2] / x.a[0];
        reinterpret_cast<double2 *>(sA)[sl] = make_double2(x.a[0], x.a[1]);
        reinterpret_cast<double2 *>(sB)[sl] = make_double2(x.a[2], x.a[3]);
        reinterpret_cast<double2 *>(sC)[sl] = make_double2(v1, v2);
        if (nc == 2) {
            const int s1 = (int)(sl2 >> 16) + T.sstride;
            reinterpret_cast<double2 *>(sA)[s1] = make_double2(x.a[0], x.a[1]);
            reinterpret_cast<double2 *>(sB)[s1] = make_double2(x.a[2], x.a[3]);
            reinterpret_cast<double2 *>(sC)[s1] = make_double2(v1, v2);
        }
    }
    const uint32_t dword = nu > 0 ? (uint32_t)__ldg(T.uslot + 2 * (u0 + nu - 1)) : 0u;  // dummy slot (copy 0), empty mask

    if constexpr (VISC != VISC_NONE) {  // the epilogue's own-row operands: pull them towards the SM now
        if (has_slice && row0 < A.n_rows) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(u + row0));
            if constexpr (R == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(u + row0 + 2));
            if constexpr (VISC == VISC_RESIDUAL) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const Vec<V> *>(A.approx_du) + row0));
                if constexpr (R == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const Vec<V> *>(A.approx_du) + row0 + 2));
            }
        }
    }
    Vec<V> acc[R], gx[R], gy[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int v = 0; v < V; ++v) acc[r].a[v] = gx[r].a[v] = gy[r].a[v] = 0.0;
        if (DO_FLUX && A.accumulate && has_slice && row0 + r < A.n_rows) acc[r] = reinterpret_cast<const Vec<V> *>(A.du)[row0 + r];
    }
    __syncthreads();
    tile_pf_issue<R>(T, pf, warp, lane);
    if constexpr (VISC == VISC_RESIDUAL) {
        if (A.norm_merge) {
            nhave = nflag >= nwant;   // (the flag was read before phase 1: the values below were written before it was set)
            if (nhave && lane < 12) nval = __ldcg(A.stats + (lane < 4 ? 4 + lane : lane < 8 ? kStatsRaw + lane - 4 : lane));
        }
    }
    if (has_slice) {
    if (W > 0) mbar_wait(&bars[warp], 0);

    const double gm1 = A.eqp0 - 1.0;
    constexpr int kBatch = R <= 2 ? MFT_TILE_BATCH_A : 2;
    auto pressure = [&](const double2 &qa, const double2 &qb, const double2 &qc) {
        const double s = fma(qa.y, qc.x, qb.x * qc.y);
        const double e = fma(-0.5, s, qb.y);
        return gm1 * e;
    };
    if constexpr (EXACT) {
        auto sweep = [&](auto dir_tag) {
            constexpr int DIR = decltype(dir_tag)::value;
            if (STAGE_W && DIR == 1 && W > 0 && L > 0) {
                // every lane is done with wx: the weight buffer is refilled with the wy blocks (prefetched into L2)
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const uint32_t wbytes = (uint32_t)L * kSlice * 8 * R;
                    mbar_expect_tx(&bars[warp], wbytes);
                    bulk_g2s(buf + (size_t)W * kSlice * 2, src + (size_t)W * kSlice * 2 + wbytes, wbytes, &bars[warp]);
                }
                mbar_wait(&bars[warp], 1);
            }
            const double *wd = wbase + (STAGE_W ? (size_t)0 : (size_t)DIR * R * rstride);
            int pos[R];
#pragma unroll
            for (int r = 0; r < R; ++r) pos[r] = 0;
            // step words + weights of the batch starting at column c0 (cursors advance in step order)
            auto fetch = [&](int c0, uint32_t (&wordv)[kBatch], double (&wv)[kBatch][R]) {
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const bool ok = c0 + b < W;
                    const int cc = ok ? c0 + b : W - 1;
                    const uint32_t word = ok ? (uint32_t)ip[cc * kSlice] : dword;
                    wordv[b] = word;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const bool m = (word >> (12 + r)) & 1u;
                        wv[b][r] = 0.0;
                        if (m) wv[b][r] = STAGE_W ? wd[r * rstride + (size_t)pos[r] * kSlice] : ld_stream(wd + r * rstride + (size_t)pos[r] * kSlice);
                        pos[r] += m ? 1 : 0;
                    }
                }
            };
            [[maybe_unused]] uint32_t wordn[kBatch];
            [[maybe_unused]] double wn[kBatch][R];
            if constexpr (MFT_TILE_PIPE_A != 0) fetch(0, wordn, wn);
            for (int c0 = 0; c0 < W; c0 += kBatch) {
                double2 qa[kBatch], qb[kBatch], qc[kBatch];
                uint32_t wordc[kBatch];
                double w[kBatch][R];
                if constexpr (MFT_TILE_PIPE_A != 0) {
#pragma unroll
                    for (int b = 0; b < kBatch; ++b) {
                        wordc[b] = wordn[b];
#pragma unroll
                        for (int r = 0; r < R; ++r) w[b][r] = wn[b][r];
                    }
                    if (c0 + kBatch < W) fetch(c0 + kBatch, wordn, wn);
                } else {
                    fetch(c0, wordc, w);
                }
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const uint32_t o = ((wordc[b] & 0xfffu) << 4) + (R <= 2 ? ((wordc[b] >> 14) & 1u) * arr_bytes : 0u);
                    qa[b] = lds2(sA, o);
                    qb[b] = lds2(sB, o);
                    qc[b] = lds2(sC, o);
                }
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    if constexpr (DO_FLUX) {
                        const double p = pressure(qa[b], qb[b], qc[b]);
                        double f[4];
                        if constexpr (DIR == 0) {
                            f[0] = qa[b].y;
                            f[1] = fma(qa[b].y, qc[b].x, p);
                            f[2] = qa[b].y * qc[b].y;
                            f[3] = (qb[b].y + p) * qc[b].x;
                        } else {
                            f[0] = qb[b].x;
                            f[1] = qb[b].x * qc[b].x;
                            f[2] = fma(qb[b].x, qc[b].y, p);
                            f[3] = (qb[b].y + p) * qc[b].y;
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) {
#pragma unroll
                            for (int v = 0; v < V; ++v) acc[r].a[v] = acc[r].a[v] + w[b][r] * (-f[v]);
                        }
                    }
                    if constexpr (VISC != VISC_NONE) {
                        const double uu[4] = {qa[b].x, qa[b].y, qb[b].x, qb[b].y};
#pragma unroll
                        for (int r = 0; r < R; ++r) {
#pragma unroll
                            for (int v = 0; v < V; ++v) {
                                if constexpr (DIR == 0) gx[r].a[v] = gx[r].a[v] + w[b][r] * uu[v];
                                else gy[r].a[v] = gy[r].a[v] + w[b][r] * uu[v];
                            }
                        }
                    }
                }
            }
        };
        sweep(std::integral_constant<int, 0>{});
        sweep(std::integral_constant<int, 1>{});
    } else {
        const double *wdx = wbase, *wdy = wbase + (size_t)R * rstride;
        int pos[R];
#pragma unroll
        for (int r = 0; r < R; ++r) pos[r] = 0;
        for (int c0 = 0; c0 < W; c0 += kBatch) {
            double2 qa[kBatch], qb[kBatch], qc[kBatch];
            double wa[kBatch][R], wb[kBatch][R];
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const bool ok = c0 + b < W;
                const int cc = ok ? c0 + b : W - 1;
                const uint32_t word = ok ? (uint32_t)ip[cc * kSlice] : dword;
                const uint32_t o = ((word & 0xfffu) << 4) + (R <= 2 ? ((word >> 14) & 1u) * arr_bytes : 0u);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const bool m = (word >> (12 + r)) & 1u;
                    wa[b][r] = wb[b][r] = 0.0;
                    if (m) {
                        wa[b][r] = ld_stream(wdx + r * rstride + (size_t)pos[r] * kSlice);
                        wb[b][r] = ld_stream(wdy + r * rstride + (size_t)pos[r] * kSlice);
                    }
                    pos[r] += m ? 1 : 0;
                }
                qa[b] = lds2(sA, o);
                qb[b] = lds2(sB, o);
                qc[b] = lds2(sC, o);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const double uu[4] = {qa[b].x, qa[b].y, qb[b].x, qb[b].y};
                if constexpr (DO_FLUX) {
                    const double p = pressure(qa[b], qb[b], qc[b]);
                    const double f[4] = {qa[b].y, fma(qa[b].y, qc[b].x, p), qa[b].y * qc[b].y, (qb[b].y + p) * qc[b].x};
                    const double h[4] = {qb[b].x, qb[b].x * qc[b].x, fma(qb[b].x, qc[b].y, p), (qb[b].y + p) * qc[b].y};
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int v = 0; v < V; ++v) acc[r].a[v] = fma(-wb[b][r], h[v], fma(-wa[b][r], f[v], acc[r].a[v]));
                    }
                }
                if constexpr (VISC != VISC_NONE) {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            gx[r].a[v] = fma(wa[b][r], uu[v], gx[r].a[v]);
                            gy[r].a[v] = fma(wb[b][r], uu[v], gy[r].a[v]);
                        }
                    }
                }
            }
        }
    }
    double nmean[4] = {0.0, 0.0, 0.0, 0.0}, nraw[4] = {0.0, 0.0, 0.0, 0.0}, nnorm[4] = {0.0, 0.0, 0.0, 0.0};
    if constexpr (VISC == VISC_RESIDUAL) {
        if (A.norm_merge) {
            // first-wave blocks only: block 0 was not done when this warp started.  The whole warp waits here, converged (the
            // last slice of a cloud may hold fewer than 32 rows: no warp-level barrier inside the per-row branch below).
            // RELAXED polls (an acquire load would invalidate the SM's L1 on every poll); the values are then read at L2.
            if (!nhave) {
                if (lane == 0) {
                    const unsigned long long want = A.L->epoch_n;
                    unsigned long long have, n = 0;
                    do {
                        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(have) : "l"(&A.L->norm_ready) : "memory");
                        if (have < want) {
                            if (++n > kSpinLimit) {
                                A.L->error = 1;
                                break;
                            }
                            __nanosleep(200);
                        }
                    } while (have < want);
                }
                __syncwarp();
                if (lane < 12) nval = __ldcg(A.stats + (lane < 4 ? 4 + lane : lane < 8 ? kStatsRaw + lane - 4 : lane));
            }
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                nmean[v] = __shfl_sync(0xffffffffu, nval, v);
                nraw[v] = __shfl_sync(0xffffffffu, nval, 4 + v);
                nnorm[v] = __shfl_sync(0xffffffffu, nval, 8 + v);
            }
        } else if (A.stats && A.norm_miss) {   // (written by the previous kernel: plain loads)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                nmean[v] = A.stats[4 + v];
                nraw[v] = A.stats[kStatsRaw + v];
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r;
        if (row < A.n_rows) {
            Vec<V> ui, ad;
#pragma unroll
            for (int v = 0; v < V; ++v) ui.a[v] = ad.a[v] = 0.0;
            if constexpr (VISC != VISC_NONE) {
                ui = ld_ro(u + row);
                if constexpr (VISC == VISC_RESIDUAL) ad = ld_ro(reinterpret_cast<const Vec<V> *>(A.approx_du) + row);
            }
            if constexpr (VISC == VISC_RESIDUAL) {
                if (A.stats && A.norm_miss) {
                    // the norms are the lexicographic (or per-component) maximum of |u - mean| over ALL points: check this row
                    // (lexicographic order: a row can only exceed the norms if its FIRST key reaches theirs -- one subtraction and
                    // one compare for all but the handful of rows at the density extremes)
                    const double d0 = fabs(ui.a[0] - nmean[0]);
                    if (!A.norm_lex || d0 >= nraw[0]) {
                        double dv[V];
                        dv[0] = d0;
#pragma unroll
                        for (int v = 1; v < V; ++v) dv[v] = fabs(ui.a[v] - nmean[v]);
                        bool miss = false;
                        if (A.norm_lex) miss = lex_less<V>(nraw, dv);
                        else {
#pragma unroll
                            for (int v = 0; v < V; ++v) miss |= dv[v] > nraw[v];
                        }
                        if (miss) atomicAdd(A.norm_miss, 1ull);
                    }
                }
            }
            Vec<2 * V> gout;
            pass_a_epilogue<V, EQ_EULER2D, DO_FLUX, VISC>(A, row, acc[r], gx[r], gy[r], ui, ad, &gout, A.norm_merge ? nnorm : nullptr);
            if constexpr (VISC != VISC_NONE) {
                if (band && A.aux) {   // rows in a peer's halo: g goes straight into the peer's halo tail
                    const int ax = __ldg(A.aux + row);
                    if (ax >= 0) {
                        const RowAux ra = A.rows[ax];
                        for (int q = ra.sbeg; q < ra.send; ++q)
                            st_vec(reinterpret_cast<Vec<2 * V> *>(A.P->field[1][A.route_peer[q]]) + A.route_dst[q], gout);
                    }
                }
            }
        }
    }
    }  // has_slice
    if constexpr (VISC != VISC_NONE) {
        if (band) {
            // the last band block to finish raises the g flags of this epoch at every destination (cumulative fences as in k_p2p_put)
            __syncthreads();
            if (threadIdx.x == 0) {
                __threadfence_system();
                const unsigned int nband = gridDim.x - (unsigned int)T.n_free;
                if (atomicAdd(&A.L->ticket_g, 1u) == nband - 1) {
                    __threadfence_system();
                    const unsigned long long e = A.L->epoch[1] + 1;
                    for (int i = 0; i < A.P->ndst; ++i) st_relaxed_sys(&A.P->win[A.P->dst[i]]->data_flag[1][A.P->rank], e);
                    A.L->epoch[1] = e;
                    A.L->ticket_g = 0;
                }
            }
        }
    }
}

// STAGE_W: two sweeps (the x chain with wx_* over arrays A,B, then the y chain with wy_* over C,D -- the chains are
// independent, so splitting them changes no sum), each with its weight blocks staged like pass A.
template <int R, bool EXACT, bool STAGE_W>
__global__ void __launch_bounds__(kTileWarps * 32, (R == 1 ? MFT_TILE_OCC_B : R == 2 ? 3 : 2)) k_pass_b_tiler(const PassBTileArgs A, const TileROp T)
{
    constexpr int V = 4;
    extern __shared__ __align__(128) unsigned char smem_dyn[];
    __shared__ uint64_t bars[kTileWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = T.order ? __ldg(T.order + blockIdx.x) : (int)blockIdx.x;
    const bool band = T.order != nullptr && (int)blockIdx.x >= T.n_free;   // several GPUs: this tile reads g rows of the halo
    const int64_t slice = (int64_t)tile * kTileWarps + warp;
    const bool has_slice = slice < A.n_slices;
    const int64_t row0 = (slice * kSlice + lane) * R;
    const uint32_t arr_bytes = (uint32_t)T.sstride * 16u, nc = (uint32_t)T.ncopy;
    unsigned char *sA = smem_dyn, *sB = smem_dyn + nc * arr_bytes, *sC = smem_dyn + 2 * nc * arr_bytes, *sD = smem_dyn + 3 * nc * arr_bytes;
    unsigned char *buf = smem_dyn + 4 * nc * arr_bytes + (size_t)warp * T.buf_bytes;
    pdl_launch_dependents();

    int W = 0, L = 0;
    const unsigned char *src = nullptr;
    if (has_slice) {
        W = T.wl[2 * slice];
        L = T.wl[2 * slice + 1];
        src = T.base + T.boff[slice];
        if (lane == 0) mbar_init(&bars[warp], 1);
        __syncwarp();
        if (lane == 0 && W > 0) {
            const uint32_t wbytes = (uint32_t)L * kSlice * 8 * R;
            const uint32_t bytes = (uint32_t)W * kSlice * 2 + (STAGE_W ? wbytes : 0u);
            mbar_expect_tx(&bars[warp], bytes);
            bulk_g2s(buf, src, bytes, &bars[warp]);
            if (wbytes > 0) {
                if (STAGE_W) bulk_prefetch_l2(src + (size_t)W * kSlice * 2 + wbytes, wbytes);
                else bulk_prefetch_l2(src + (size_t)W * kSlice * 2, 2 * wbytes);
            }
        }
    }
    const TilePf pf = tile_pf_begin(T, A.n_slices, warp);
    const unsigned short *ip = reinterpret_cast<const unsigned short *>(buf) + lane;
    const double *wdx = reinterpret_cast<const double *>((STAGE_W ? buf : src) + (size_t)W * kSlice * 2) + lane;
    const size_t rstride = (size_t)L * kSlice;
    const double *wdy = wdx + (STAGE_W ? (size_t)0 : (size_t)R * rstride);
    const Vec<2 * V> *__restrict__ g = reinterpret_cast<const Vec<2 * V> *>(A.g);

    const int u0 = T.uoff[tile];
    const int nu = T.uoff[tile + 1] - u0;
    pdl_wait();                         // operator prologue done; g and du are the predecessor's output
    if (band) band_wait(A.P, A.L, 1);   // the peers' g rows of this stage are in the halo tail from here on
    for (int t = threadIdx.x; t < nu; t += kTileWarps * 32) {
        const int j = __ldg(T.ulist + u0 + t);
        const unsigned int sl2 = __ldg(reinterpret_cast<const unsigned int *>(T.uslot) + u0 + t);
        const int sl = (int)(sl2 & 0xffffu);
        const Vec<2 * V> x = band ? ld_cg(g + j) : ld_ro(g + j);
        reinterpret_cast<double2 *>(sA)[sl] = make_double2(x.a[0], x.a[1]);
        reinterpret_cast<double2 *>(sB)[sl] = make_double2(x.a[2], x.a[3]);
        reinterpret_cast<double2 *>(sC)[sl] = make_double2(x.a[4], x.a[5]);
        reinterpret_cast<double2 *>(sD)[sl] = make_double2(x.a[6], x.a[7]);
        if (nc == 2) {
            const int s1 = (int)(sl2 >> 16) + T.sstride;
            reinterpret_cast<double2 *>(sA)[s1] = make_double2(x.a[0], x.a[1]);
            reinterpret_cast<double2 *>(sB)[s1] = make_double2(x.a[2], x.a[3]);
            reinterpret_cast<double2 *>(sC)[s1] = make_double2(x.a[4], x.a[5]);
            reinterpret_cast<double2 *>(sD)[s1] = make_double2(x.a[6], x.a[7]);
        }
    }
    const uint32_t dword = nu > 0 ? (uint32_t)__ldg(T.uslot + 2 * (u0 + nu - 1)) : 0u;
    if (has_slice && row0 < A.n_rows) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const Vec<V> *>(A.du) + row0));
        if constexpr (R == 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const Vec<V> *>(A.du) + row0 + 2));
    }
    Vec<V> tx[R], ty[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int v = 0; v < V; ++v) tx[r].a[v] = ty[r].a[v] = 0.0;
    }
    __syncthreads();
    tile_pf_issue<R>(T, pf, warp, lane);
    if (!has_slice) return;
    if (W > 0) mbar_wait(&bars[warp], 0);

    constexpr int kBatch = R <= 2 ? MFT_TILE_BATCH_B : 2;
    if constexpr (STAGE_W) {
        auto sweep = [&](auto dir_tag) {
            constexpr int DIR = decltype(dir_tag)::value;
            if (DIR == 1 && W > 0 && L > 0) {
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const uint32_t wbytes = (uint32_t)L * kSlice * 8 * R;
                    mbar_expect_tx(&bars[warp], wbytes);
                    bulk_g2s(buf + (size_t)W * kSlice * 2, src + (size_t)W * kSlice * 2 + wbytes, wbytes, &bars[warp]);
                }
                mbar_wait(&bars[warp], 1);
            }
            const unsigned char *s0 = DIR == 0 ? sA : sC, *s1 = DIR == 0 ? sB : sD;
            int pos[R];
#pragma unroll
            for (int r = 0; r < R; ++r) pos[r] = 0;
            for (int c0 = 0; c0 < W; c0 += kBatch) {
                double2 qa[kBatch], qb[kBatch];
                double w[kBatch][R];
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const bool ok = c0 + b < W;
                    const int cc = ok ? c0 + b : W - 1;
                    const uint32_t word = ok ? (uint32_t)ip[cc * kSlice] : dword;
                    const uint32_t o = ((word & 0xfffu) << 4) + (R <= 2 ? ((word >> 14) & 1u) * arr_bytes : 0u);
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const bool m = (word >> (12 + r)) & 1u;
                        w[b][r] = 0.0;
                        if (m) w[b][r] = wdx[r * rstride + (size_t)pos[r] * kSlice];
                        pos[r] += m ? 1 : 0;
                    }
                    qa[b] = lds2(s0, o);
                    qb[b] = lds2(s1, o);
                }
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    const double gv[4] = {qa[b].x, qa[b].y, qb[b].x, qb[b].y};
#pragma unroll
                    for (int r = 0; r < R; ++r) {
#pragma unroll
                        for (int v = 0; v < V; ++v) {
                            if constexpr (DIR == 0) {
                                if constexpr (EXACT) tx[r].a[v] = tx[r].a[v] + w[b][r] * gv[v];
                                else tx[r].a[v] = fma(w[b][r], gv[v], tx[r].a[v]);
                            } else {
                                if constexpr (EXACT) ty[r].a[v] = ty[r].a[v] + w[b][r] * gv[v];
                                else ty[r].a[v] = fma(w[b][r], gv[v], ty[r].a[v]);
                            }
                        }
                    }
                }
            }
        };
        sweep(std::integral_constant<int, 0>{});
        sweep(std::integral_constant<int, 1>{});
    } else {
        int pos[R];
#pragma unroll
        for (int r = 0; r < R; ++r) pos[r] = 0;
        auto fetch = [&](int c0, uint32_t (&wordv)[kBatch], double (&wav)[kBatch][R], double (&wbv)[kBatch][R]) {
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const bool ok = c0 + b < W;
                const int cc = ok ? c0 + b : W - 1;
                const uint32_t word = ok ? (uint32_t)ip[cc * kSlice] : dword;
                wordv[b] = word;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const bool m = (word >> (12 + r)) & 1u;
                    wav[b][r] = wbv[b][r] = 0.0;
                    if (m) {
                        wav[b][r] = ld_stream(wdx + r * rstride + (size_t)pos[r] * kSlice);
                        wbv[b][r] = ld_stream(wdy + r * rstride + (size_t)pos[r] * kSlice);
                    }
                    pos[r] += m ? 1 : 0;
                }
            }
        };
        [[maybe_unused]] uint32_t wordn[kBatch];
        [[maybe_unused]] double wan[kBatch][R], wbn[kBatch][R];
        if constexpr (MFT_TILE_PIPE_B != 0) fetch(0, wordn, wan, wbn);
        for (int c0 = 0; c0 < W; c0 += kBatch) {
            double2 qa[kBatch], qb[kBatch], qc[kBatch], qd[kBatch];
            uint32_t wordc[kBatch];
            double wa[kBatch][R], wb[kBatch][R];
            if constexpr (MFT_TILE_PIPE_B != 0) {
#pragma unroll
                for (int b = 0; b < kBatch; ++b) {
                    wordc[b] = wordn[b];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        wa[b][r] = wan[b][r];
                        wb[b][r] = wbn[b][r];
                    }
                }
                if (c0 + kBatch < W) fetch(c0 + kBatch, wordn, wan, wbn);
            } else {
                fetch(c0, wordc, wa, wb);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const uint32_t o = ((wordc[b] & 0xfffu) << 4) + (R <= 2 ? ((wordc[b] >> 14) & 1u) * arr_bytes : 0u);
                qa[b] = lds2(sA, o);
                qb[b] = lds2(sB, o);
                qc[b] = lds2(sC, o);
                qd[b] = lds2(sD, o);
            }
#pragma unroll
            for (int b = 0; b < kBatch; ++b) {
                const double gxv[4] = {qa[b].x, qa[b].y, qb[b].x, qb[b].y};
                const double gyv[4] = {qc[b].x, qc[b].y, qd[b].x, qd[b].y};
#pragma unroll
                for (int r = 0; r < R; ++r) {
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if constexpr (EXACT) {
                            tx[r].a[v] = tx[r].a[v] + wa[b][r] * gxv[v];
                            ty[r].a[v] = ty[r].a[v] + wb[b][r] * gyv[v];
                        } else {
                            tx[r].a[v] = fma(wa[b][r], gxv[v], tx[r].a[v]);
                            ty[r].a[v] = fma(wb[b][r], gyv[v], ty[r].a[v]);
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r;
        if (row < A.n_rows) {
            Vec<V> d = reinterpret_cast<const Vec<V> *>(A.du)[row];
#pragma unroll
            for (int v = 0; v < V; ++v) d.a[v] = (d.a[v] + tx[r].a[v] * -1.0) + ty[r].a[v] * -1.0;
            st_vec(reinterpret_cast<Vec<V> *>(A.du) + row, d);
        }
    }
}

}  // namespace mft
