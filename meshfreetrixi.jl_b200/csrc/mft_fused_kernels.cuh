// mft_fused_kernels.cuh -- the fused SSPRK33 stage kernel of the device-resident step (round 2).
//
// One launch replaces, per Runge-Kutta stage, what used to be five (six on several GPUs):
//     BC pass 2 of the previous rhs!  (calc_boundary_flux!, rbfsolver.jl:288-318: du and u at boundary points)
//     the SSPRK33 stage update         (OrdinaryDiffEq stage formulas)
//     BC pass 1 of the next rhs!       (strong BCs written into u)
//     ode_mean(u)                      (src/auxiliary/mpi.jl:40-52)
//     ode_maximum(|u - mean|)          (src/auxiliary/mpi.jl:71-81; hyperviscosity.jl:305-311)
//     the halo put of u                (perform_halo_update!, src/auxiliary/mpi.jl:217-265)
// so u, uprev and du are read once and u is written once per stage, and on several GPUs the norms need ONE exchange of a
// small record per stage instead of two dependent ones.
//
// The norms in one pass.  ode_maximum compares SVectors with isless, i.e. LEXICOGRAPHICALLY: the result is the deviation
// vector |u_p - mean| of the point p that maximises (|rho - m0|, |m1 - m_1|, |m2 - m_2|, |E - m_3|) in lexicographic order.
// The mean is only known after the pass -- but x -> fl|x - m| is monotone on either side of m, so whatever m turns out to
// be, the maximiser of the first key has rho = max rho or rho = min rho; among the points that tie there, the maximiser of
// the second key has the largest or the smallest m1; and so on.  Hence the 16 "leaves"
//     leaf(s0,s1,s2,s3) = lexicographic maximum of (s0 rho, s1 m1, s2 m2, s3 E),   s in {+,-}^4
// are a sufficient statistic: the answer is the lexicographic maximum of the 16 leaf deviation vectors, evaluated once
// the mean is known.  Leaves merge associatively and commutatively (they are maxima), so blocks and ranks combine them in
// any order with a bit-identical result; only the sums keep a fixed order.  Exact value ties (a far field at rho = 1.0
// exactly with different momenta: 8 % of the points of the vortex workload, 1808 of them pinned by Dirichlet data) are
// what the nested leaves are for.
//
// Rounding ties on the first key.  fl|rho - m0| is monotone but not injective: with ode_mean's divisor V*N the mean sits on a
// finer binary grid than the densities, rho - m0 falls between two doubles, and two ADJACENT densities (one ulp apart) can round
// to the same deviation (ties-to-even at an exact half).  The 2-GPU bench cloud does it in its very first stage: 17 far-field
// points share the largest density, the stage update moves five of them by one ulp, and the lexicographic order between the
// two groups is then decided on |m1 - mean| (profiles/r2y_norm_miss_diagnostic_2ranks.log).  A single rounding can only merge
// neighbours, so every record keeps the leaves of the extreme density AND of the next distinct density on either side
// ("second level"); the finalize evaluates all 32 deviation vectors and takes their lexicographic maximum -- candidates whose
// rounded first key is smaller simply lose.  What still cannot be seen is a rounding tie on a LATER key (a plateau that ties
// exactly on rho and carries rounding noise in m1 around an O(1) mean: any number of values collapse) or a first key whose
// deviation is coarser than the densities (|rho - m0| > 2 rho: never for positive densities and 0 <= m0 <= max rho); pass A
// checks every row against the norms it was given and counts misses (MFT_FIELD_NORM_MISSES -> MFT_ENORMS), and
// MFT_OPT_FUSED_STEP = 0 selects the two-pass kernels.
#pragma once
#include "mft_kernels.cuh"

namespace mft {

// ---- the record ----------------------------------------------------------------------------------------------------
constexpr int kRecSum = 0;     // 4: sum of the states
constexpr int kRecExt = 4;     // 2: max rho, min rho
constexpr int kRecLeaf = 6;    // 2 sides x 8 leaves x (m1, m2, E); leaf index bit c set = component c+1 is minimised
constexpr int kRecCmax = 54;   // per-component mode (MFT_OPT_MAX_LEXICOGRAPHIC = 0): max and min of every component
constexpr int kRecCmin = 58;
constexpr int kRecExt2 = 62;   // 2: the largest density below max rho, the smallest above min rho (-inf / +inf: none)
constexpr int kRecLeaf2 = 64;  // their leaves, laid out like kRecLeaf
// kRecDoubles = 128 (mft_kernels.cuh): 112 used, padded to 1 KB
static_assert(kRecLeaf2 + 48 <= kRecDoubles, "norm record layout");

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }
__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

// is a = (a1,a2,a3) lexicographically better than b under the leaf's sign bits (bit c set: smaller is better)?
__device__ __forceinline__ bool leaf_better(double a1, double a2, double a3, double b1, double b2, double b3, int bits)
{
    const double a[3] = {a1, a2, a3}, b[3] = {b1, b2, b3};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const bool mn = (bits >> c) & 1;
        if (mn ? a[c] < b[c] : a[c] > b[c]) return true;
        if (mn ? a[c] > b[c] : a[c] < b[c]) return false;
    }
    return false;
}

// order-preserving map double -> uint64 (NaNs excluded by the callers): unsigned order of the keys = numeric order
__device__ __forceinline__ unsigned long long f64_key(double x)
{
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000LL));
}
__device__ __forceinline__ double key_f64(unsigned long long k)
{
    const long long b = (long long)k;
    return __longlong_as_double(b ^ (((~b) >> 63) | (long long)0x8000000000000000LL));
}
// The second density level only has to cover values that can share a rounded deviation with the extreme: its neighbours on the
// grid of doubles.  kNearUlps bounds the distance (in representable steps) at which a value is still kept as "second level"; 1
// would do while the deviation is no coarser than the densities (0 <= mean <= max rho, comparable minimum), a few more cover
// a minimum that is several times smaller than the mean.  near_enough(a, b): b lies at most kNearUlps steps inside of a.
constexpr unsigned long long kNearUlps = 4;
template <int SIDE>
__device__ __forceinline__ bool near_enough(double a, double b)
{
    const unsigned long long ka = f64_key(a), kb = f64_key(b);
    return SIDE == 0 ? (kb <= ka && ka - kb <= kNearUlps) : (kb >= ka && kb - ka <= kNearUlps);
}
// the value kNearUlps steps inside of a (a finite; -inf / +inf map to themselves: nothing has been seen yet)
template <int SIDE>
__device__ __forceinline__ double window_bound(double a)
{
    if (!(a > neg_inf() && a < pos_inf())) return a;
    const unsigned long long ka = f64_key(a);
    return key_f64(SIDE == 0 ? ka - kNearUlps : ka + kNearUlps);
}

// maximum (MIN: minimum) of x over the lanes with `active` set, through two 32-bit REDUX operations on the keys instead of
// five rounds of 64-bit shuffles + FP64 compares; inactive lanes contribute the neutral key.  All 32 lanes must call it.
template <bool MIN>
__device__ __forceinline__ double warp_extreme(double x, bool active)
{
    unsigned long long k = f64_key(x);
    if (MIN) k = ~k;
    if (!active) k = 0ull;
    const unsigned hi = __reduce_max_sync(kFull, (unsigned)(k >> 32));
    const unsigned lo = __reduce_max_sync(kFull, (unsigned)(k >> 32) == hi ? (unsigned)k : 0u);
    unsigned long long r = ((unsigned long long)hi << 32) | lo;
    if (MIN) r = ~r;
    return key_f64(r);
}

// The 8 leaves of a tie set: `tied` = the lanes of this chunk that hold one density value.  Lanes 0..7 return leaf `lane`
// (l1, l2, l3) = (m1, m2, E) of the point that is lexicographically extreme under the leaf's sign pattern.  Warp-uniform.
__device__ __forceinline__ void tie_set_leaves(unsigned tied, double m1, double m2, double E, int lane, double &l1, double &l2, double &l3)
{
    if (__popc(tied) == 1) {
        // one point: it is every leaf of the set
        const int src = __ffs(tied) - 1;
        l1 = __shfl_sync(kFull, m1, src);
        l2 = __shfl_sync(kFull, m2, src);
        l3 = __shfl_sync(kFull, E, src);
        return;
    }
    // Level 1 first: only the lanes that hold the largest or the smallest m1 of the tie set can be a leaf (leaves with bit
    // 0 clear maximise m1, the others minimise it).  On noisy plateaus (rho ties exactly, the momenta carry rounding noise)
    // each of the two groups is a single lane and no deeper reduction is needed.
    const bool in = (tied >> lane) & 1u;
    const double g1max = warp_extreme<false>(m1, in), g1min = warp_extreme<true>(m1, in);
    l1 = l2 = l3 = 0.0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {   // half 0: leaves 0,2,4,6 (max m1); half 1: leaves 1,3,5,7 (min m1)
        const double g1 = half == 0 ? g1max : g1min;
        const unsigned grp = __ballot_sync(kFull, in && m1 == g1);
        const bool ing = (grp >> lane) & 1u;
        const int first = __ffs(grp) - 1;
        const double r2 = __shfl_sync(kFull, m2, first), r3 = __shfl_sync(kFull, E, first);
        if (lane < 8 && (lane & 1) == half) {
            l1 = g1;
            l2 = r2;
            l3 = r3;
        }
        if (__ballot_sync(kFull, ing && (m2 != r2 || E != r3)) != 0) {
            // several different states share rho and m1: nested extremes of (m2, E) per sign pattern
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {   // leaf = half | q << 1: bit 1 = minimise m2, bit 2 = minimise E
                const bool mn2 = q & 1, mn3 = (q >> 1) & 1;
                const double c2 = mn2 ? warp_extreme<true>(m2, ing) : warp_extreme<false>(m2, ing);
                const bool in2 = ing && m2 == c2;
                const double c3 = mn3 ? warp_extreme<true>(E, in2) : warp_extreme<false>(E, in2);
                if (lane == (half | (q << 1))) {
                    l2 = c2;
                    l3 = c3;
                }
            }
        }
    }
}

// One chunk of 32 points (one per lane) against the warp's running record `wr` (shared memory), side 0 = max rho,
// side 1 = min rho.  Warp-uniform register state: run_ext = wr[kRecExt + SIDE], the extreme density so far; has2 / run_thr: is
// there a second level (the next distinct density, at most kNearUlps steps inside the extreme), and the value a row has to reach
// to matter at all -- that second density, or the window bound while there is none.  A chunk can contribute its two most
// extreme distinct values; anything else in it is third at best.  Warp-uniform control flow.
template <int SIDE>
__device__ __forceinline__ void lex_side_update(double *wr, bool valid, double rho, double m1, double m2, double E, double &run_ext,
                                                double &run_thr, bool &has2, int lane)
{
    auto reaches = [](double x, double y) { return SIDE == 0 ? x >= y : x <= y; };
    auto beyond = [](double x, double y) { return SIDE == 0 ? x > y : x < y; };
    unsigned hm = __ballot_sync(kFull, valid && reaches(rho, run_thr));   // (a NaN is never hot)
    if (hm == 0) return;   // the common case: nobody comes near the running extreme
#pragma unroll 1
    for (int level = 0; level < 2; ++level) {
        double cm;
        unsigned tied;
        if (__popc(hm) == 1) {   // one candidate: no reduction
            cm = __shfl_sync(kFull, rho, __ffs(hm) - 1);
            tied = hm;
        } else {
            const bool in = (hm >> lane) & 1u;
            cm = warp_extreme<SIDE == 1>(rho, in);
            tied = __ballot_sync(kFull, in && rho == cm);
        }
        hm &= ~tied;
        // where the value goes: 0 beyond the running extreme, 1 equal to it, 3 equal to the second density, 2 a new second density
        // (between the two, or the first one inside the window)
        const int where = beyond(cm, run_ext) ? 0 : cm == run_ext ? 1 : (has2 && cm == run_thr) ? 3 : 2;
        double *leaf = wr + (where <= 1 ? kRecLeaf : kRecLeaf2) + SIDE * 24;
        if (where == 1 || where == 3) {
            // a value seen before: a point only matters if its m1 reaches the running extremes of m1 over that tie set
            const double m1max = leaf[0], m1min = leaf[3];   // leaf 0: maximise m1; leaf 1: minimise m1
            tied = __ballot_sync(kFull, ((tied >> lane) & 1u) && (m1 >= m1max || m1 <= m1min));
        }
        if (tied != 0) {
            double l1, l2, l3;   // leaf `lane` of the tie set (lanes 0..7)
            tie_set_leaves(tied, m1, m2, E, lane, l1, l2, l3);
            __syncwarp();
            double *lf = leaf + (lane & 7) * 3;
            if (where == 0) {
                // the old extreme and its leaves become the second level if they are near enough, else there is none any more
                const bool keep = near_enough<SIDE>(cm, run_ext);
                if (lane < 8) {
                    if (keep) {
                        double *l2nd = wr + kRecLeaf2 + SIDE * 24 + lane * 3;
                        l2nd[0] = lf[0];
                        l2nd[1] = lf[1];
                        l2nd[2] = lf[2];
                    }
                    lf[0] = l1;
                    lf[1] = l2;
                    lf[2] = l3;
                }
                if (lane == 0) {
                    wr[kRecExt2 + SIDE] = keep ? run_ext : (SIDE == 0 ? neg_inf() : pos_inf());
                    wr[kRecExt + SIDE] = cm;
                }
                has2 = keep;
                run_thr = keep ? run_ext : window_bound<SIDE>(cm);
                run_ext = cm;
            } else if (where == 2) {
                if (lane < 8) {
                    lf[0] = l1;
                    lf[1] = l2;
                    lf[2] = l3;
                }
                if (lane == 0) wr[kRecExt2 + SIDE] = cm;
                has2 = true;
                run_thr = cm;
            } else if (lane < 8 && leaf_better(l1, l2, l3, lf[0], lf[1], lf[2], lane)) {
                lf[0] = l1;
                lf[1] = l2;
                lf[2] = l3;
            }
            __syncwarp();
        }
        // the rest of the chunk only matters where it still reaches the (possibly new) threshold
        hm = __ballot_sync(kFull, ((hm >> lane) & 1u) && reaches(rho, run_thr));
        if (hm == 0) break;
    }
}

// ---- slots ------------------------------------------------------------------------------------------------------------
// slot = side * 8 + leaf.  A thread that owns a slot carries the leaf of the extreme density (e; a1..a3) and of the next distinct
// density (e2; c1..c3); records are folded in with slot_insert: leaves are maxima, so any order gives the same bits.
struct Slot2 {
    double e, a1, a2, a3, e2, c1, c2, c3;
};
__device__ __forceinline__ void slot_init(int slot, Slot2 &S)
{
    S.e = S.e2 = (slot >> 3) == 0 ? neg_inf() : pos_inf();
    S.a1 = S.a2 = S.a3 = S.c1 = S.c2 = S.c3 = 0.0;
}
// one density value `ext` with its leaf (l1, l2, l3) into the two-level state of the slot (second level: the next distinct density,
// kept only while it lies within kNearUlps steps of the extreme -- the same rule as in lex_side_update)
__device__ __forceinline__ void slot_insert(int slot, double ext, double l1, double l2, double l3, Slot2 &S)
{
    const int side = slot >> 3, lf = slot & 7;
    auto near = [&](double a, double b) { return side == 0 ? near_enough<0>(a, b) : near_enough<1>(a, b); };
    const bool beyond1 = side == 0 ? ext > S.e : ext < S.e;
    if (beyond1) {
        if (near(ext, S.e)) {
            S.e2 = S.e;
            S.c1 = S.a1;
            S.c2 = S.a2;
            S.c3 = S.a3;
        } else {
            S.e2 = side == 0 ? neg_inf() : pos_inf();
        }
        S.e = ext;
        S.a1 = l1;
        S.a2 = l2;
        S.a3 = l3;
    } else if (ext == S.e) {
        if (leaf_better(l1, l2, l3, S.a1, S.a2, S.a3, lf)) {
            S.a1 = l1;
            S.a2 = l2;
            S.a3 = l3;
        }
    } else if (near(S.e, ext)) {
        if (side == 0 ? ext > S.e2 : ext < S.e2) {
            S.e2 = ext;
            S.c1 = l1;
            S.c2 = l2;
            S.c3 = l3;
        } else if (ext == S.e2 && leaf_better(l1, l2, l3, S.c1, S.c2, S.c3, lf)) {
            S.c1 = l1;
            S.c2 = l2;
            S.c3 = l3;
        }
    }
}
__device__ __forceinline__ void slot_merge(int slot, const Slot2 &X, Slot2 &S)
{
    slot_insert(slot, X.e, X.a1, X.a2, X.a3, S);
    if (X.e2 > neg_inf() && X.e2 < pos_inf()) slot_insert(slot, X.e2, X.c1, X.c2, X.c3, S);   // (absent in almost every record)
}
// LD: 0 plain (shared memory / this block's data), 1 __ldcg (another block's record, at L2), 2 __ldcv (another GPU's record)
template <int LD>
__device__ __forceinline__ Slot2 slot_load(const double *R, int slot)
{
    auto ld = [](const double *p) -> double { return LD == 0 ? *p : LD == 1 ? __ldcg(p) : __ldcv(p); };
    const int side = slot >> 3, lf = slot & 7;
    const double *q = R + kRecLeaf + side * 24 + lf * 3, *q2 = R + kRecLeaf2 + side * 24 + lf * 3;
    Slot2 X;
    X.e = ld(R + kRecExt + side);
    X.a1 = ld(q);
    X.a2 = ld(q + 1);
    X.a3 = ld(q + 2);
    X.e2 = ld(R + kRecExt2 + side);
    X.c1 = X.c2 = X.c3 = 0.0;
    if (X.e2 > neg_inf() && X.e2 < pos_inf()) {
        X.c1 = ld(q2);
        X.c2 = ld(q2 + 1);
        X.c3 = ld(q2 + 2);
    }
    return X;
}
__device__ __forceinline__ void slot_store(double *R, int slot, const Slot2 &S)
{
    const int side = slot >> 3, lf = slot & 7;
    if (lf == 0) {
        R[kRecExt + side] = S.e;
        R[kRecExt2 + side] = S.e2;
    }
    double *q = R + kRecLeaf + side * 24 + lf * 3, *q2 = R + kRecLeaf2 + side * 24 + lf * 3;
    q[0] = S.a1;
    q[1] = S.a2;
    q[2] = S.a3;
    q2[0] = S.c1;
    q2[1] = S.c2;
    q2[2] = S.c3;
}
__device__ __forceinline__ Slot2 slot_shfl_xor(const Slot2 &S, int mask)
{
    Slot2 X;
    X.e = __shfl_xor_sync(kFull, S.e, mask);
    X.a1 = __shfl_xor_sync(kFull, S.a1, mask);
    X.a2 = __shfl_xor_sync(kFull, S.a2, mask);
    X.a3 = __shfl_xor_sync(kFull, S.a3, mask);
    X.e2 = __shfl_xor_sync(kFull, S.e2, mask);
    X.c1 = __shfl_xor_sync(kFull, S.c1, mask);
    X.c2 = __shfl_xor_sync(kFull, S.c2, mask);
    X.c3 = __shfl_xor_sync(kFull, S.c3, mask);
    return X;
}

// ode_mean + ode_maximum(|u - mean|) from `nrec` records (one per rank; the sums are combined in rank order), by ONE WARP:
// lanes 0..3 form the means; lane = level * 16 + slot folds its (side, leaf) slot over the ranks and evaluates the deviation vector
// of the slot's first (lanes 0..15) or second (lanes 16..31) density level; a shuffle tree takes the lexicographic maximum of
// the 32 vectors.  Writes mean[4], norms[4] (zero -> eps) and raw[4] (before the zero replacement: what pass A verifies rows
// against).  VOL: the records were written by other GPUs -- read them past L1.
template <bool VOL>
__device__ inline void norms_from_records(const double *recs, int stride, int nrec, double divisor, int lex, double *mean_out, double *norms_out, double *raw_out, int lane)
{
    auto ld = [](const double *p) -> double { return VOL ? __ldcv(p) : *p; };
    double mv = 0.0;
    if (lane < 4) {
        double t = ld(recs + kRecSum + lane);
        for (int r = 1; r < nrec; ++r) t += ld(recs + (size_t)r * stride + kRecSum + lane);
        mv = t / divisor;
        if (mean_out) mean_out[lane] = mv;
    }
    double m[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) m[v] = __shfl_sync(kFull, mv, v);
    double c[4] = {-1.0, -1.0, -1.0, -1.0};
    if (lex) {
        const int slot = lane & 15;
        Slot2 S;
        slot_init(slot, S);
        for (int r = 0; r < nrec; ++r) slot_merge(slot, slot_load<VOL ? 2 : 0>(recs + (size_t)r * stride, slot), S);
        const bool second = lane >= 16;
        const double ext = second ? S.e2 : S.e, b1 = second ? S.c1 : S.a1, b2 = second ? S.c2 : S.a2, b3 = second ? S.c3 : S.a3;
        if (ext > neg_inf() && ext < pos_inf()) {   // (no such density level anywhere, or a non-finite state: no candidate)
            c[0] = fabs(ext - m[0]);
            c[1] = fabs(b1 - m[1]);
            c[2] = fabs(b2 - m[2]);
            c[3] = fabs(b3 - m[3]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double d[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) d[v] = __shfl_xor_sync(kFull, c[v], o);
            if (lex_less<4>(c, d)) {
#pragma unroll
                for (int v = 0; v < 4; ++v) c[v] = d[v];
            }
        }
    } else {
        double b = -1.0;
        if (lane < 4)
            for (int r = 0; r < nrec; ++r) {
                const double *R = recs + (size_t)r * stride;
                b = jl_max(b, fabs(ld(R + kRecCmax + lane) - mv));
                b = jl_max(b, fabs(ld(R + kRecCmin + lane) - mv));
            }
#pragma unroll
        for (int v = 0; v < 4; ++v) c[v] = __shfl_sync(kFull, b, v);
    }
    if (lane == 0) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            if (raw_out) raw_out[v] = c[v];
            norms_out[v] = c[v] == 0.0 ? kEps : c[v];
        }
    }
}

// ---- per-row side table ---------------------------------------------------------------------------------------------
// aux[row] = -1 for a plain row, else an index into RowAux: the row's entry in the merged boundary table and its slice
// of the halo routing table (a row can be in the halo of several peers)
struct RowAux {
    int bc;            // index into the merged BC table, or -1
    int sbeg, send;    // [sbeg, send) into route_peer / route_dst
};

struct StageArgs {
    int stage;         // 1, 2, 3
    int apply_bc2;     // BC pass 2 of the previous rhs! has not been applied yet (stages fused behind pass B)
    double dt;
    double *u, *uprev, *du;
    int64_t n;         // owned rows
    const int *aux;    // nullable
    const RowAux *rows;
    // merged boundary table (k_boundary_merged)
    const int *bc_kind;
    const double *bc_normals, *bc_values;
    // reductions
    double *partial;   // gridDim.x block records
    double *grec;      // one record per group of kStageGroup blocks (leaves / component extremes only)
    unsigned int *gticket;   // one ticket per group, zero between launches
    unsigned int *ticket;
    double divisor;
    int lex;
    double *stats;     // sum[4] | mean[4] | norms[4] | ... | raw norms at [kStatsRaw]
    // peer-memory exchange (nranks > 1)
    P2PPeers P;
    P2PLocal *L;
    const int *route_peer;
    const long long *route_dst;
};
constexpr int kStatsRaw = 16;   // stats[16..19]: norms before the zero replacement

constexpr int NORMS_NONE = 0, NORMS_LEX = 1, NORMS_COMP = 2;

// resident blocks per SM (the grid is one wave of them) and depth of the operand ring in shared memory
#ifndef MFT_STAGE_OCC
#define MFT_STAGE_OCC 2
#endif
#ifndef MFT_STAGE_DEPTH
#define MFT_STAGE_DEPTH 3
#endif
constexpr int kStageDepth = MFT_STAGE_DEPTH;
constexpr int kStageGroup = 16;   // blocks whose records the last of them to finish merges into one group record
constexpr int kStageSlotBytes = 256 * (3 * 32 + 4);   // per ring stage and block: du, u, uprev rows + aux of 8 warps x 32 rows
constexpr int kStageSmemBytes = kStageDepth * kStageSlotBytes;

// grid = ctx red_blocks x 256 threads, the grid of k_sum_mean: same rows per thread and same reduction tree => same sums
template <int NMODE, bool MULTI>
__global__ void __launch_bounds__(256, MFT_STAGE_OCC) k_stage_fused(const StageArgs A)
{
    constexpr int V = 4;
    constexpr bool NORMS = NMODE != NORMS_NONE;
    __shared__ double wrec[8][kRecDoubles];
    __shared__ double sh[8][V];
    __shared__ bool is_last;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Vec<V> *u = reinterpret_cast<Vec<V> *>(A.u);
    Vec<V> *uprev = reinterpret_cast<Vec<V> *>(A.uprev);
    Vec<V> *du = reinterpret_cast<Vec<V> *>(A.du);
    pdl_launch_dependents();
    pdl_wait();   // every operand of this kernel is the predecessor's output (du from pass B)
    unsigned long long e_u = 0;
    if constexpr (MULTI) {
        e_u = A.L->epoch[0] + 1;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            // credits for the two-kernel exchange protocol (k_p2p_put): everything stream-ordered before this kernel has
            // consumed the halos of both fields' last epochs
            for (int f = 0; f < 2; ++f) {
                const unsigned long long ef = A.L->epoch[f];
                for (int i = 0; i < A.P.nsrc; ++i) st_release_sys(&A.P.win[A.P.src[i]]->credit[f][A.P.rank], ef);
            }
        }
    }
    double s[V] = {0.0, 0.0, 0.0, 0.0};
    double cmx[NMODE == NORMS_COMP ? V : 1], cmn[NMODE == NORMS_COMP ? V : 1];
    if constexpr (NMODE == NORMS_COMP) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            cmx[v] = neg_inf();
            cmn[v] = pos_inf();
        }
    }
    double run_max = neg_inf(), run_min = pos_inf(), thr_max = neg_inf(), thr_min = pos_inf();
    bool has2_max = false, has2_min = false;
    if (NMODE == NORMS_LEX && lane < 2) wrec[w][kRecExt + lane] = wrec[w][kRecExt2 + lane] = lane == 0 ? neg_inf() : pos_inf();
    __syncwarp();
    const double dt = A.dt, dt2 = 2.0 * A.dt;
    const int stage = A.stage;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    // Operand rings: every warp streams ITS next row batches (du, u, uprev, aux of 32 consecutive rows) into shared memory with
    // bulk-async copies (cp.async.bulk = the 1-D TMA path) kStageDepth batches ahead, completion on an mbarrier per (warp,
    // slot); no block-level barrier in the loop.  The kernel is a pure stream (128 B per row): what bounds it is bytes in
    // flight, and 8 warps x 3 slots x 3.2 KB per block keep ~150 KB per SM in flight without a single register.
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ uint64_t full[8][kStageDepth];
    struct Slot {
        Vec<V> k[32], uo[32], up[32];
        int ax[32];
    };
    static_assert(sizeof(Slot) * 8 * kStageDepth == kStageSmemBytes, "ring size");
    Slot *ring = reinterpret_cast<Slot *>(ring_raw) + (size_t)w * kStageDepth;
    const int64_t base0 = (int64_t)blockIdx.x * blockDim.x + (int64_t)w * 32;
    auto issue = [&](int it) {   // lane 0 only
        const int64_t base = base0 + (int64_t)it * stride;
        if (base >= A.n) return;
        const int slot = it % kStageDepth;
        const uint32_t nrows = (uint32_t)(A.n - base < 32 ? A.n - base : 32);
        const uint32_t rb = nrows * 32u, ab = ((nrows * 4u + 15u) / 16u) * 16u;   // (aux is padded to a multiple of 16 bytes)
        mbar_expect_tx(&full[w][slot], rb * (stage != 1 ? 3u : 2u) + ab);
        bulk_g2s(ring[slot].k, du + base, rb, &full[w][slot]);
        bulk_g2s(ring[slot].uo, u + base, rb, &full[w][slot]);
        if (stage != 1) bulk_g2s(ring[slot].up, uprev + base, rb, &full[w][slot]);
        bulk_g2s(ring[slot].ax, A.aux + base, ab, &full[w][slot]);
    };
    if (lane == 0) {
        for (int d = 0; d < kStageDepth; ++d) mbar_init(&full[w][d], 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (lane == 0)
        for (int d = 0; d < kStageDepth; ++d) issue(d);
    struct RowIn {
        Vec<V> k, uo, up;
        int ax;
    };
    RowIn cur;
    int it = 0;
    for (int64_t i = base0 + lane; i - lane < A.n; i += stride, ++it) {
        const int slot = it % kStageDepth;
        mbar_wait(&full[w][slot], (uint32_t)(it / kStageDepth) & 1u);
        cur.ax = -1;
        if (i < A.n) {
            cur.k = ring[slot].k[lane];
            cur.uo = ring[slot].uo[lane];
            if (stage != 1) cur.up = ring[slot].up[lane];
            cur.ax = ring[slot].ax[lane];
        }
        __syncwarp();   // every lane holds its row in registers: the slot can be refilled
        if (lane == 0) issue(it + kStageDepth);
        const bool valid = i < A.n;
        Vec<V> un;
#pragma unroll
        for (int v = 0; v < V; ++v) un.a[v] = 0.0;
        if (valid) {
            const int ax = cur.ax;
            Vec<V> k = cur.k;
            Vec<V> uo = cur.uo;
            int bcj = -1, kind = -1;
            RowAux ra{-1, 0, 0};
            if (ax >= 0) {
                ra = A.rows[ax];
                bcj = ra.bc;
                if (bcj >= 0) kind = A.bc_kind[bcj];
            }
            double nx = 0.0, ny = 0.0;
            if (kind == 1) {   // slip wall: unit normal as k_boundary_merged forms it
                const double nx0 = A.bc_normals[2 * bcj], ny0 = A.bc_normals[2 * bcj + 1];
                const double nrm = sqrt(nx0 * nx0 + ny0 * ny0);
                nx = nx0 / nrm;
                ny = ny0 / nrm;
            }
            if (A.apply_bc2 && kind >= 0) {
                // BC pass 2 of the rhs! that produced k (rbfsolver.jl:424-427): u and du at the boundary point
                if (kind == 0) {
#pragma unroll
                    for (int v = 0; v < V; ++v) k.a[v] = 0.0;   // u already holds the Dirichlet values of that rhs!
                } else {
                    const double vdotn = uo.a[1] * nx + uo.a[2] * ny;
                    uo.a[1] = uo.a[1] - vdotn * nx;
                    uo.a[2] = uo.a[2] - vdotn * ny;
                    k.a[1] = 0.0;
                    k.a[2] = 0.0;
                }
                st_vec(du + i, k);
            }
            if (stage == 1) {
                st_vec(uprev + i, uo);
#pragma unroll
                for (int v = 0; v < V; ++v) un.a[v] = fma(dt, k.a[v], uo.a[v]);
            } else {
                const Vec<V> up = cur.up;
                if (stage == 2) {
#pragma unroll
                    for (int v = 0; v < V; ++v) un.a[v] = fma(dt, k.a[v], fma(3.0, up.a[v], uo.a[v])) / 4.0;
                } else {
#pragma unroll
                    for (int v = 0; v < V; ++v) un.a[v] = fma(dt2, k.a[v], fma(2.0, uo.a[v], up.a[v])) / 3.0;
                }
            }
            if (kind == 0) {   // BC pass 1 of the next rhs!: Dirichlet values of the new stage time
#pragma unroll
                for (int v = 0; v < V; ++v) un.a[v] = A.bc_values[(int64_t)bcj * V + v];
            } else if (kind == 1) {
                const double vdotn = un.a[1] * nx + un.a[2] * ny;
                un.a[1] = un.a[1] - vdotn * nx;
                un.a[2] = un.a[2] - vdotn * ny;
            }
            st_vec(u + i, un);
            if constexpr (MULTI) {
                for (int q = ra.sbeg; q < ra.send; ++q)
                    st_vec(reinterpret_cast<Vec<V> *>(A.P.field[0][A.route_peer[q]]) + A.route_dst[q], un);
            }
        }
        if constexpr (NORMS) {
            if (valid) {
#pragma unroll
                for (int v = 0; v < V; ++v) s[v] += un.a[v];
            }
            if constexpr (NMODE == NORMS_LEX) {
                lex_side_update<0>(wrec[w], valid, un.a[0], un.a[1], un.a[2], un.a[3], run_max, thr_max, has2_max, lane);
                lex_side_update<1>(wrec[w], valid, un.a[0], un.a[1], un.a[2], un.a[3], run_min, thr_min, has2_min, lane);
            } else if (valid) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    cmx[v] = fmax(cmx[v], un.a[v]);
                    cmn[v] = fmin(cmn[v], un.a[v]);
                }
            }
        }
    }
    if constexpr (!NORMS && !MULTI) return;

    // ---- block record -> partial[blockIdx.x] -----------------------------------------------------------------------
    if constexpr (NORMS) {
        // the sums: the reduction tree of k_sum_mean (same grid, same order => the same bits)
#pragma unroll
        for (int v = 0; v < V; ++v)
            for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_down_sync(kFull, s[v], o);
        if (lane == 0)
            for (int v = 0; v < V; ++v) sh[w][v] = s[v];
        if constexpr (NMODE == NORMS_COMP) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                cmx[v] = warp_extreme<false>(cmx[v], true);
                cmn[v] = warp_extreme<true>(cmn[v], true);
            }
            if (lane == 0)
                for (int v = 0; v < V; ++v) {
                    wrec[w][kRecCmax + v] = cmx[v];
                    wrec[w][kRecCmin + v] = cmn[v];
                }
        }
    }
    __syncthreads();
    double *prec = A.partial + (size_t)blockIdx.x * kRecDoubles;
    if constexpr (NORMS) {
        if (threadIdx.x == 0) {
            for (int v = 0; v < V; ++v) {
                double t = 0.0;
                for (int k = 0; k < 8; ++k) t += sh[k][v];
                prec[kRecSum + v] = t;
            }
        }
        if constexpr (NMODE == NORMS_LEX) {
            if (threadIdx.x < 16) {
                const int slot = threadIdx.x;
                Slot2 S;
                slot_init(slot, S);
                for (int k = 0; k < 8; ++k) slot_merge(slot, slot_load<0>(wrec[k], slot), S);
                slot_store(prec, slot, S);
            }
        } else if (threadIdx.x < 8) {
            const int v = threadIdx.x & 3;
            const bool mn = threadIdx.x >= 4;
            double b = mn ? pos_inf() : neg_inf();
            for (int k = 0; k < 8; ++k) b = mn ? fmin(b, wrec[k][kRecCmin + v]) : fmax(b, wrec[k][kRecCmax + v]);
            prec[(mn ? kRecCmin : kRecCmax) + v] = b;
        }
    }
    // ---- group records: the last block of every group of kStageGroup blocks merges the group's leaves (parallel over the
    // groups: no single serial tail; whoever combines the whole grid afterwards reads gridDim.x / 16 records, one round trip)
    __shared__ double fin[kRecDoubles];
    __shared__ Slot2 xw[8][16];
    __shared__ bool glast;
    const int nb = (int)gridDim.x;
    if constexpr (NORMS) {
        const int g = (int)blockIdx.x / kStageGroup;
        const int gsize = nb - g * kStageGroup < kStageGroup ? nb - g * kStageGroup : kStageGroup;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            glast = atomicAdd(&A.gticket[g], 1u) == (unsigned)gsize - 1u;
        }
        __syncthreads();
        if (glast) {
            __threadfence();
            double *G = A.grec + (size_t)g * kRecDoubles;
            if constexpr (NMODE == NORMS_LEX) {
                const int slot = threadIdx.x & 15, j = threadIdx.x >> 4;
                Slot2 S;
                slot_init(slot, S);
                if (j < gsize) S = slot_load<1>(A.partial + (size_t)(g * kStageGroup + j) * kRecDoubles, slot);
                slot_merge(slot, slot_shfl_xor(S, 16), S);   // lanes slot and slot + 16 hold the same slot
                if (lane < 16) xw[w][slot] = S;
                __syncthreads();
                if (threadIdx.x < 16) {
                    for (int k = 1; k < 8; ++k) slot_merge(slot, xw[k][slot], S);
                    slot_store(G, slot, S);
                }
            } else if (threadIdx.x < 8) {
                const int v = threadIdx.x & 3;
                const bool mn = threadIdx.x >= 4;
                double b = mn ? pos_inf() : neg_inf();
                for (int k = 0; k < gsize; ++k) {
                    const double x = __ldcg(&A.partial[(size_t)(g * kStageGroup + k) * kRecDoubles + (mn ? kRecCmin : kRecCmax) + v]);
                    b = mn ? fmin(b, x) : fmax(b, x);
                }
                G[(mn ? kRecCmin : kRecCmax) + v] = b;
            }
            if (threadIdx.x == 0) A.gticket[g] = 0;
        }
    }
    // fences are cumulative: the block barrier makes every thread's stores (incl. the remote halo rows and a group record) visible
    // to thread 0, whose fence then orders them before the ticket (and, in the last block, before the flags)
    __syncthreads();
    if (threadIdx.x == 0) {
        if constexpr (MULTI) __threadfence_system();
        else __threadfence();
        is_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    // ---- last block: this rank's record = block sums (tree of k_sum_mean) + merged group records (one L2 round trip) --------
    const int ng = (nb + kStageGroup - 1) / kStageGroup;
    if constexpr (NORMS) {
#pragma unroll
        for (int v = 0; v < V; ++v) s[v] = 0.0;
        for (int b = threadIdx.x; b < nb; b += blockDim.x)
            for (int v = 0; v < V; ++v) s[v] += __ldcg(&A.partial[(size_t)b * kRecDoubles + kRecSum + v]);
#pragma unroll
        for (int v = 0; v < V; ++v)
            for (int o = 16; o > 0; o >>= 1) s[v] += __shfl_down_sync(kFull, s[v], o);
        __syncthreads();
        if (lane == 0)
            for (int v = 0; v < V; ++v) sh[w][v] = s[v];
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int v = 0; v < V; ++v) {
                double t = 0.0;
                for (int k = 0; k < 8; ++k) t += sh[k][v];
                fin[kRecSum + v] = t;
            }
        }
        if constexpr (NMODE == NORMS_LEX) {
            // thread = (j, slot): 16 scanners per (side, leaf) slot over the group records, then the scanners of a slot merge
            const int slot = threadIdx.x & 15, j = threadIdx.x >> 4;
            Slot2 S;
            slot_init(slot, S);
            for (int b = j; b < ng; b += 16) slot_merge(slot, slot_load<1>(A.grec + (size_t)b * kRecDoubles, slot), S);
            slot_merge(slot, slot_shfl_xor(S, 16), S);   // lanes slot and slot + 16 hold the same slot
            if (lane < 16) xw[w][slot] = S;
            __syncthreads();
            if (threadIdx.x < 16) {
                for (int k = 1; k < 8; ++k) slot_merge(slot, xw[k][slot], S);
                slot_store(fin, slot, S);
            }
        } else if (threadIdx.x < 8) {
            const int v = threadIdx.x & 3;
            const bool mn = threadIdx.x >= 4;
            double b = mn ? pos_inf() : neg_inf();
            for (int k = 0; k < ng; ++k) {
                const double x = __ldcg(&A.grec[(size_t)k * kRecDoubles + (mn ? kRecCmin : kRecCmax) + v]);
                b = mn ? fmin(b, x) : fmax(b, x);
            }
            fin[(mn ? kRecCmin : kRecCmax) + v] = b;
        }
        __syncthreads();
    }
    if constexpr (MULTI) {
        const unsigned long long en = A.L->epoch_n + 1;
        const int par = (int)(en & 1);
        if constexpr (NORMS) {
            // this rank's record into every rank's window (own included)
            for (int t = threadIdx.x; t < A.P.nranks * kRecDoubles; t += blockDim.x) {
                const int r = t / kRecDoubles, d = t % kRecDoubles;
                A.P.win[r]->rec[par][A.P.rank][d] = fin[d];
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence_system();   // ONE fence orders the halo rows and the record before all the flags below
            if constexpr (NORMS)
                for (int r = 0; r < A.P.nranks; ++r) st_relaxed_sys(&A.P.win[r]->rec_flag[par][A.P.rank], en);
            for (int i = 0; i < A.P.ndst; ++i) st_relaxed_sys(&A.P.win[A.P.dst[i]]->data_flag[0][A.P.rank], e_u);
            A.L->epoch[0] = e_u;
            if constexpr (NORMS) A.L->epoch_n = en;
            *A.ticket = 0;
        }
    } else {
        // one GPU: the norms are final here; pass A reads them from stats[] (stream order)
        if (w == 0) {
            if (lane < V) A.stats[lane] = fin[kRecSum + lane];
            norms_from_records<false>(fin, kRecDoubles, 1, A.divisor, A.lex, A.stats + V, A.stats + 2 * V, A.stats + kStatsRaw, lane);
            if (lane == 0) *A.ticket = 0;
        }
    }
}

// Several GPUs: block 0 of pass A turns the ranks' records into the norms while the other blocks run their main loops;
// epilogues wait for norm_ready.  Executed by the threads of one warp (lane 0 does the arithmetic).
__device__ inline void p2p_norms_merge(const P2PPeers &P, P2PLocal *L, double divisor, int lex, double *stats, int lane)
{
    const unsigned long long en = L->epoch_n;
    const int par = (int)(en & 1);
    P2PWindow *win = P.win[P.rank];
    if (lane < P.nranks) spin_until(&win->rec_flag[par][lane], en, &L->error);
    __syncwarp();
    norms_from_records<true>(&win->rec[par][0][0], kRecDoubles, P.nranks, divisor, lex, stats + 4, stats + 8, stats + kStatsRaw, lane);
    __syncwarp();
    if (lane == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(&L->norm_ready), "l"(en) : "memory");
    }
}

}  // namespace mft
