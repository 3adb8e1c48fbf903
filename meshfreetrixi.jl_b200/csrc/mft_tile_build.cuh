// mft_tile_build.cuh -- the union-tile layout of ONE tile (R = 1), written once for the host and the device.
//
// build_tiler_host (mft_layout_host.inl) is the reference implementation of the tile format: sorted stencil union, bank-coloured
// slots, second record copy tuned by a local search, per-phase copy choice, step words and compact weight blocks.  The code below
// restates its per-tile work on raw arrays with caller-provided scratch and without the C++ standard library, so that the same
// source runs
//   * on the GPU, one tile per warp (lane 0 walks the sequential greedy / local-search logic; tiles are independent, a million
//     points are 8224 tiles = 1.7 tiles per resident warp): k_tile_sizes / k_tile_emit in mft_b200.cu, and
//   * on the CPU (mft_debug_tile_build_compare), where its output is compared BYTE FOR BYTE with build_tiler_host's.
// Byte identity is the contract: every loop below follows the order of the host builder, ties are broken the same way, the
// pseudo-random second-copy permutation uses the same generator jumped to the tile's position.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define MFT_HD __host__ __device__ inline
#else
#define MFT_HD inline
#endif

namespace mft {
namespace tb {

constexpr int kNB = 8;         // 16-byte bank groups seen by one LDS.128 phase (8 lanes)
constexpr int kLanes = 32;     // lanes (rows) per slice
#ifndef MFT_TILE_WARPS
#define MFT_TILE_WARPS 4
#endif
constexpr int kWarps = MFT_TILE_WARPS;
constexpr int kRows = kLanes * kWarps;   // rows per tile

struct Op {   // operator rows in caller numbering, entries of a row in summation order
    const long long *ptr;
    const int *col;
    const double *wx, *wy;
    const int *perm, *iperm;   // device -> caller row, caller -> device column; nullable (identity)
    long long nrows_dev;
    int n_tot;
};

struct Flags {
    int colour, two_copies, tune;
};

// per-worker scratch, carved out of one byte buffer (scratch_bytes / scratch_bind)
struct Scratch {
    int kmax, cap_nu, cap_req;
    int *cols;               // kRows * kmax: the tile's columns, then its sorted union
    unsigned short *node;    // kRows * kmax: union index of entry cpos of the tile's row i at [i * kmax + cpos]
    int *len;                // kRows
    // pass 2 only
    unsigned short *adj;     // cap_nu^2 co-request counts
    int *deg, *order, *bank, *slot, *slot1, *bank1;   // cap_nu each
    int *rq_ptr, *rq_node, *nr_ptr, *nr_idx, *stamp;
    char *rq_ok, *rq_pad;
};

MFT_HD size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// bytes of one worker's scratch: pass 1 needs the first three arrays only (cap_nu = 0)
MFT_HD size_t scratch_bytes(int kmax, int cap_nu)
{
    size_t b = 0;
    b += align16(sizeof(int) * (size_t)kRows * kmax);
    b += align16(sizeof(unsigned short) * (size_t)kRows * kmax);
    b += align16(sizeof(int) * kRows);
    if (cap_nu > 0) {
        const size_t cap_req = (size_t)kWarps * kmax * (kLanes / kNB);
        b += align16(sizeof(unsigned short) * (size_t)cap_nu * cap_nu);
        b += 6 * align16(sizeof(int) * (size_t)cap_nu);
        b += align16(sizeof(int) * (cap_req + 1));         // rq_ptr
        b += align16(sizeof(int) * cap_req * kNB);         // rq_node
        b += align16(sizeof(int) * ((size_t)cap_nu + 1));  // nr_ptr
        b += align16(sizeof(int) * cap_req * kNB);         // nr_idx
        b += align16(sizeof(int) * cap_req);               // stamp
        b += 2 * align16(cap_req);                         // rq_ok, rq_pad
    }
    return b;
}

MFT_HD void scratch_bind(Scratch &S, unsigned char *base, int kmax, int cap_nu)
{
    S.kmax = kmax;
    S.cap_nu = cap_nu;
    S.cap_req = kWarps * kmax * (kLanes / kNB);
    unsigned char *p = base;
    auto take = [&](size_t bytes) {
        unsigned char *q = p;
        p += align16(bytes);
        return q;
    };
    S.cols = reinterpret_cast<int *>(take(sizeof(int) * (size_t)kRows * kmax));
    S.node = reinterpret_cast<unsigned short *>(take(sizeof(unsigned short) * (size_t)kRows * kmax));
    S.len = reinterpret_cast<int *>(take(sizeof(int) * kRows));
    S.adj = nullptr;
    if (cap_nu > 0) {
        const size_t cr = (size_t)S.cap_req;
        S.adj = reinterpret_cast<unsigned short *>(take(sizeof(unsigned short) * (size_t)cap_nu * cap_nu));
        S.deg = reinterpret_cast<int *>(take(sizeof(int) * (size_t)cap_nu));
        S.order = reinterpret_cast<int *>(take(sizeof(int) * (size_t)cap_nu));
        S.bank = reinterpret_cast<int *>(take(sizeof(int) * (size_t)cap_nu));
        S.slot = reinterpret_cast<int *>(take(sizeof(int) * (size_t)cap_nu));
        S.slot1 = reinterpret_cast<int *>(take(sizeof(int) * (size_t)cap_nu));
        S.bank1 = reinterpret_cast<int *>(take(sizeof(int) * (size_t)cap_nu));
        S.rq_ptr = reinterpret_cast<int *>(take(sizeof(int) * (cr + 1)));
        S.rq_node = reinterpret_cast<int *>(take(sizeof(int) * cr * kNB));
        S.nr_ptr = reinterpret_cast<int *>(take(sizeof(int) * ((size_t)cap_nu + 1)));
        S.nr_idx = reinterpret_cast<int *>(take(sizeof(int) * cr * kNB));
        S.stamp = reinterpret_cast<int *>(take(sizeof(int) * cr));
        S.rq_ok = reinterpret_cast<char *>(take(cr));
        S.rq_pad = reinterpret_cast<char *>(take(cr));
    }
}

// state of x -> a x + c (mod 2^64) after k more steps (the generator of the second-copy permutations)
MFT_HD uint64_t lcg_jump(uint64_t x, uint64_t k)
{
    uint64_t cur_a = 6364136223846793005ULL, cur_c = 1442695040888963407ULL, acc_a = 1, acc_c = 0;
    for (; k; k >>= 1) {
        if (k & 1) {
            acc_a *= cur_a;
            acc_c = acc_c * cur_a + cur_c;
        }
        cur_c = (cur_a + 1) * cur_c;
        cur_a *= cur_a;
    }
    return acc_a * x + acc_c;
}

// <= 8 points, each offering two bank groups: can they take pairwise different groups, none of them a blocked one?
// Points are edges between their two groups (a blocked option folds the edge onto the other group): an assignment exists iff no
// connected component holds more edges than groups -- the same answer as the augmenting-path matching of the host builder
// (BankMatcher::perfect; compared exhaustively in tests/test_abi_cpu.py), without recursion.
MFT_HD bool two_choice_feasible(const int *opt0, const int *opt1, int kk, unsigned blocked)
{
    int par[kNB], ecnt[kNB], vcnt[kNB];
    for (int b = 0; b < kNB; ++b) {
        par[b] = b;
        ecnt[b] = 0;
        vcnt[b] = 1;
    }
    for (int i = 0; i < kk; ++i) {
        int a = opt0[i], b = opt1[i];
        if ((blocked >> a) & 1u) a = b;
        if ((blocked >> b) & 1u) b = a;
        if ((blocked >> a) & 1u) return false;
        while (par[a] != a) a = par[a];
        while (par[b] != b) b = par[b];
        if (a != b) {
            par[b] = a;
            ecnt[a] += ecnt[b];
            vcnt[a] += vcnt[b];
        }
        if (++ecnt[a] > vcnt[a]) return false;
    }
    return true;
}

MFT_HD void heap_sort(int *a, int n)
{
    auto sift = [&](int i, int m) {
        const int x = a[i];
        for (;;) {
            int ch = 2 * i + 1;
            if (ch >= m) break;
            if (ch + 1 < m && a[ch + 1] > a[ch]) ++ch;
            if (a[ch] <= x) break;
            a[i] = a[ch];
            i = ch;
        }
        a[i] = x;
    };
    for (int i = n / 2 - 1; i >= 0; --i) sift(i, n);
    for (int m = n - 1; m > 0; --m) {
        const int x = a[0];
        a[0] = a[m];
        a[m] = x;
        sift(0, m);
    }
}

// The tile's sorted stencil union (S.cols[0..nu)), the row lengths (S.len) and the union index of every entry (S.node);
// returns nu, or -1 when a row is longer than the scratch allows.
MFT_HD int tile_union(const Op &A, long long t, Scratch &S)
{
    const long long d0 = t * kRows, d1 = A.nrows_dev < d0 + kRows ? A.nrows_dev : d0 + kRows;
    int nc = 0;
    for (int i = 0; i < kRows; ++i) S.len[i] = 0;
    for (long long d = d0; d < d1; ++d) {
        const long long r = A.perm ? A.perm[d] : d;
        const long long b = A.ptr[r], e = A.ptr[r + 1];
        if (e - b > S.kmax) return -1;
        S.len[d - d0] = (int)(e - b);
        for (long long p = b; p < e; ++p) S.cols[nc++] = A.iperm ? A.iperm[A.col[p]] : A.col[p];
    }
    heap_sort(S.cols, nc);
    int nu = 0;
    for (int i = 0; i < nc; ++i)
        if (i == 0 || S.cols[i] != S.cols[i - 1]) S.cols[nu++] = S.cols[i];
    for (long long d = d0; d < d1; ++d) {
        const long long r = A.perm ? A.perm[d] : d;
        const long long b = A.ptr[r], e = A.ptr[r + 1];
        unsigned short *nd = S.node + (size_t)(d - d0) * S.kmax;
        for (long long p = b; p < e; ++p) {
            const int j = A.iperm ? A.iperm[A.col[p]] : A.col[p];
            int lo = 0, hi = nu;   // lower_bound
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (S.cols[mid] < j) lo = mid + 1;
                else hi = mid;
            }
            nd[p - b] = (unsigned short)lo;
        }
    }
    return nu;
}

// pass 1: union size of tile t and (W, L) of its slices (R = 1: both are the longest row of the slice)
MFT_HD int tile_sizes(const Op &A, long long t, Scratch &S, int *wl)
{
    const int nu = tile_union(A, t, S);
    if (nu < 0) return nu;
    const long long d0 = t * kRows, d1 = A.nrows_dev < d0 + kRows ? A.nrows_dev : d0 + kRows;
    const int ns_tile = (int)((d1 - d0 + kLanes - 1) / kLanes);
    for (int si = 0; si < ns_tile; ++si) {
        int W = 0;
        for (int l = 0; l < kLanes; ++l) W = S.len[si * kLanes + l] > W ? S.len[si * kLanes + l] : W;
        wl[2 * (t * kWarps + si)] = W;
        wl[2 * (t * kWarps + si) + 1] = W;
    }
    return nu;
}

// distinct union points that lanes [ph*8, ph*8+8) of slice si request at step cpos, in lane order; pad: a lane has no such step
MFT_HD int phase_group(const Scratch &S, int si, int cpos, int ph, int *grp, bool *pad)
{
    int kk = 0;
    bool pd = false;
    for (int l = ph * kNB; l < (ph + 1) * kNB; ++l) {
        const int i = si * kLanes + l;
        if (cpos >= S.len[i]) {
            pd = true;
            continue;
        }
        const int nd = S.node[(size_t)i * S.kmax + cpos];
        bool dup = false;
        for (int q = 0; q < kk; ++q) dup |= grp[q] == nd;
        if (!dup) grp[kk++] = nd;
    }
    if (pad) *pad = pd;
    return kk;
}

struct Out {   // the arrays of the layout (HostTileR / DevTileR)
    unsigned char *blob;
    const long long *boff;
    const int *wl;
    const int *uoff;
    int *ulist;
    unsigned short *uslot;
};

// pass 2: slots, second copy, step words, weights, union list of tile t.  lcg_skip: steps of the second-copy generator consumed by
// the tiles before t.  Returns the tile's slot count (its dummy record takes slot `nslots`), or -1 (scratch too small).
MFT_HD int tile_emit(const Op &A, const Flags &F, long long t, uint64_t lcg_skip, Scratch &S, const Out &O)
{
    const int nu = tile_union(A, t, S);
    if (nu < 0 || nu > S.cap_nu) return -1;
    const long long d0 = t * kRows, d1 = A.nrows_dev < d0 + kRows ? A.nrows_dev : d0 + kRows;
    const long long s0 = d0 / kLanes;
    const int ns_tile = (int)((d1 - d0 + kLanes - 1) / kLanes);
    const bool two_copies = F.two_copies != 0, tune = F.tune != 0;
    int *slot = S.slot, *slot1 = S.slot1, *bank = S.bank, *bank1 = S.bank1, *deg = S.deg, *order = S.order;
    int grp[kNB];
    auto sliceW = [&](int si) {
        int W = 0;
        for (int l = 0; l < kLanes; ++l) W = S.len[si * kLanes + l] > W ? S.len[si * kLanes + l] : W;
        return W;
    };
    int nslots = nu;
    if (!F.colour || nu <= kNB) {
        for (int q = 0; q < nu; ++q) slot[q] = q;
    } else {
        unsigned short *adj = S.adj;
        for (size_t i = 0; i < (size_t)nu * nu; ++i) adj[i] = 0;
        for (int si = 0; si < ns_tile; ++si) {
            const int W = sliceW(si);
            for (int cpos = 0; cpos < W; ++cpos)
                for (int ph = 0; ph < kLanes / kNB; ++ph) {
                    const int kk = phase_group(S, si, cpos, ph, grp, nullptr);
                    for (int x = 0; x < kk; ++x)
                        for (int y = 0; y < kk; ++y)
                            if (x != y) {
                                unsigned short &e = adj[(size_t)grp[x] * nu + grp[y]];
                                if (e < 0xffff) ++e;
                            }
                }
        }
        for (int a = 0; a < nu; ++a) {
            const unsigned short *row = adj + (size_t)a * nu;
            int sum = 0;
            for (int b = 0; b < nu; ++b) sum += row[b];
            deg[a] = sum;
        }
        // stable sort by descending degree (insertion sort: equal degrees keep their index order)
        for (int i = 0; i < nu; ++i) {
            int j = i;
            while (j > 0 && deg[order[j - 1]] < deg[i]) {
                order[j] = order[j - 1];
                --j;
            }
            order[j] = i;
        }
        for (int q = 0; q < nu; ++q) bank[q] = -1;
        int fill[kNB] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto choose = [&](int a) {
            long long cost[kNB] = {0, 0, 0, 0, 0, 0, 0, 0};
            const unsigned short *row = adj + (size_t)a * nu;
            for (int b = 0; b < nu; ++b)
                if (row[b] && bank[b] >= 0) cost[bank[b]] += row[b];
            int bestb = 0;
            for (int k = 1; k < kNB; ++k)
                if (cost[k] < cost[bestb] || (cost[k] == cost[bestb] && fill[k] < fill[bestb])) bestb = k;
            return bestb;
        };
        for (int i = 0; i < nu; ++i) {
            const int a = order[i];
            bank[a] = choose(a);
            ++fill[bank[a]];
        }
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; i < nu; ++i) {
                const int a = order[i];
                --fill[bank[a]];
                bank[a] = -1;
                bank[a] = choose(a);
                ++fill[bank[a]];
            }
        if (tune && two_copies) {
            for (int o = 0; o < nu; o += kNB) {
                const int oe = nu < o + kNB ? nu : o + kNB;
                unsigned used = 0;
                int dup[kNB], nd = 0;
                for (int q = o; q < oe; ++q) {
                    if (used & (1u << bank[q])) dup[nd++] = q;
                    else used |= 1u << bank[q];
                }
                int b = 0;
                for (int i = 0; i < nd; ++i) {
                    while (used & (1u << b)) ++b;
                    bank[dup[i]] = b;
                    used |= 1u << b;
                }
            }
            nslots = 0;
            for (int q = 0; q < nu; ++q) {
                slot[q] = (q / kNB) * kNB + bank[q];
                nslots = nslots > slot[q] + 1 ? nslots : slot[q] + 1;
            }
        } else {
            int level[kNB] = {0, 0, 0, 0, 0, 0, 0, 0};
            nslots = 0;
            for (int q = 0; q < nu; ++q) {
                slot[q] = level[bank[q]]++ * kNB + bank[q];
                nslots = nslots > slot[q] + 1 ? nslots : slot[q] + 1;
            }
        }
    }
    if (two_copies) {
        uint64_t lcg = lcg_jump(0x9e3779b97f4a7c15ULL, lcg_skip);
        for (int q = 0; q < nu; ++q) slot1[q] = q;
        for (int q = nu - 1; q > 0; --q) {
            lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
            const int j = (int)((lcg >> 33) % (uint64_t)(q + 1));
            const int x = slot1[q];
            slot1[q] = slot1[j];
            slot1[j] = x;
        }
        nslots = nslots > nu ? nslots : nu;
        if (tune && nu > kNB) {
            int *rq_ptr = S.rq_ptr, *rq_node = S.rq_node, *nr_ptr = S.nr_ptr, *nr_idx = S.nr_idx, *stamp = S.stamp;
            char *rq_ok = S.rq_ok, *rq_pad = S.rq_pad;
            int nreq = 0, nrn = 0;
            rq_ptr[0] = 0;
            for (int si = 0; si < ns_tile; ++si) {
                const int W = sliceW(si);
                for (int cpos = 0; cpos < W; ++cpos)
                    for (int ph = 0; ph < kLanes / kNB; ++ph) {
                        bool pad = false;
                        const int kk = phase_group(S, si, cpos, ph, grp, &pad);
                        if (kk + (pad ? 1 : 0) < 2) continue;
                        for (int q = 0; q < kk; ++q) rq_node[nrn++] = grp[q];
                        rq_pad[nreq] = pad ? 1 : 0;
                        rq_ptr[++nreq] = nrn;
                    }
            }
            for (int q = 0; q <= nu; ++q) nr_ptr[q] = 0;
            for (int e = 0; e < nrn; ++e) nr_ptr[rq_node[e] + 1]++;
            for (int q = 0; q < nu; ++q) nr_ptr[q + 1] += nr_ptr[q];
            {
                int *cur = S.deg;   // (the colouring is done with it)
                for (int q = 0; q < nu; ++q) cur[q] = nr_ptr[q];
                for (int r = 0; r < nreq; ++r)
                    for (int e = rq_ptr[r]; e < rq_ptr[r + 1]; ++e) nr_idx[cur[rq_node[e]]++] = r;
            }
            for (int o = 0; o < nu; o += kNB) {
                const int oe = nu < o + kNB ? nu : o + kNB;
                for (int q = o; q < oe; ++q) {
                    int rank = 0;
                    for (int q2 = o; q2 < oe; ++q2) rank += slot1[q2] < slot1[q] ? 1 : 0;
                    bank1[q] = rank;
                }
            }
            auto matched = [&](int r) -> bool {
                int o0[kNB], o1[kNB];
                const int e0 = rq_ptr[r], kk = rq_ptr[r + 1] - e0;
                for (int i = 0; i < kk; ++i) {
                    o0[i] = slot[rq_node[e0 + i]] % kNB;
                    o1[i] = bank1[rq_node[e0 + i]];
                }
                return two_choice_feasible(o0, o1, kk, rq_pad[r] ? 1u : 0u);
            };
            for (int r = 0; r < nreq; ++r) {
                rq_ok[r] = matched(r) ? 1 : 0;
                stamp[r] = -1;
            }
            for (int sweep = 0; sweep < 2; ++sweep)
                for (int q = 0; q < nu; ++q) {
                    int bad_q = 0;
                    for (int e = nr_ptr[q]; e < nr_ptr[q + 1]; ++e) bad_q += rq_ok[nr_idx[e]] ? 0 : 1;
                    if (bad_q == 0) continue;
                    const int o = (q / kNB) * kNB, oe = nu < o + kNB ? nu : o + kNB;
                    int best2 = -1, best_gain = 0;
                    for (int q2 = o; q2 < oe; ++q2) {
                        if (q2 == q) continue;
                        const int mark = q * kNB + (q2 - o);
                        int before = bad_q, after = 0;
                        for (int e = nr_ptr[q]; e < nr_ptr[q + 1]; ++e) stamp[nr_idx[e]] = mark;
                        for (int e = nr_ptr[q2]; e < nr_ptr[q2 + 1]; ++e)
                            if (stamp[nr_idx[e]] != mark) before += rq_ok[nr_idx[e]] ? 0 : 1;
                        int x = bank1[q];
                        bank1[q] = bank1[q2];
                        bank1[q2] = x;
                        for (int e = nr_ptr[q]; e < nr_ptr[q + 1] && before - after > best_gain; ++e) after += matched(nr_idx[e]) ? 0 : 1;
                        for (int e = nr_ptr[q2]; e < nr_ptr[q2 + 1] && before - after > best_gain; ++e)
                            if (stamp[nr_idx[e]] != mark) after += matched(nr_idx[e]) ? 0 : 1;
                        x = bank1[q];
                        bank1[q] = bank1[q2];
                        bank1[q2] = x;
                        if (before - after > best_gain) {
                            best_gain = before - after;
                            best2 = q2;
                        }
                    }
                    if (best2 >= 0) {
                        const int x = bank1[q];
                        bank1[q] = bank1[best2];
                        bank1[best2] = x;
                        for (int e = nr_ptr[q]; e < nr_ptr[q + 1]; ++e) rq_ok[nr_idx[e]] = matched(nr_idx[e]) ? 1 : 0;
                        for (int e = nr_ptr[best2]; e < nr_ptr[best2 + 1]; ++e) rq_ok[nr_idx[e]] = matched(nr_idx[e]) ? 1 : 0;
                    }
                }
            for (int q = 0; q < nu; ++q) slot1[q] = (q / kNB) * kNB + bank1[q];
            nslots = (nu + kNB - 1) / kNB * kNB;
        }
    }
    if (nslots + 1 > 4095) return nslots;   // the caller reports it; nothing is written for this tile
    // ---- the slices ----------------------------------------------------------------------------------------------------------
    for (int si = 0; si < ns_tile; ++si) {
        const long long s = s0 + si;
        const int W = O.wl[2 * s], L = O.wl[2 * s + 1];
        const size_t word_bytes = (size_t)W * kLanes * 2, wblk = (size_t)L * kLanes * 8;
        unsigned char *at0 = O.blob + O.boff[s];
        unsigned short *word = reinterpret_cast<unsigned short *>(at0);
        for (int q = 0; q < W * kLanes; ++q) word[q] = (unsigned short)nslots;
        for (int l = 0; l < kLanes; ++l) {
            const int i = si * kLanes + l;
            for (int cpos = 0; cpos < S.len[i]; ++cpos)
                word[(size_t)cpos * kLanes + l] = (unsigned short)(slot[S.node[(size_t)i * S.kmax + cpos]] | (1u << 12));
        }
        if (two_copies) {
            for (int cpos = 0; cpos < W; ++cpos)
                for (int ph = 0; ph < kLanes / kNB; ++ph) {
                    bool pad_any = false;
                    const int kk = phase_group(S, si, cpos, ph, grp, &pad_any);
                    const bool pad = tune && nu > kNB && pad_any;
                    if (kk + (pad ? 1 : 0) < 2) continue;
                    int best_bits = 0, best_max = 99;
                    for (int bits = 0; bits < (1 << kk) && best_max > 1; ++bits) {
                        int cnt[kNB] = {pad ? 1 : 0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
                        for (int q = 0; q < kk; ++q) {
                            const int c2 = ++cnt[(((bits >> q) & 1) ? slot1[grp[q]] : slot[grp[q]]) % kNB];
                            mx = mx > c2 ? mx : c2;
                        }
                        if (mx < best_max) {
                            best_max = mx;
                            best_bits = bits;
                        }
                    }
                    for (int l = ph * kNB; l < (ph + 1) * kNB; ++l) {
                        const int i = si * kLanes + l;
                        if (cpos >= S.len[i]) continue;
                        const int nd = S.node[(size_t)i * S.kmax + cpos];
                        int q = 0;
                        while (grp[q] != nd) ++q;
                        if ((best_bits >> q) & 1) word[(size_t)cpos * kLanes + l] = (unsigned short)(slot1[nd] | (1u << 12) | 0x4000u);
                    }
                }
        }
        double *wx = reinterpret_cast<double *>(at0 + word_bytes);
        double *wy = reinterpret_cast<double *>(at0 + word_bytes + wblk);
        for (int l = 0; l < kLanes; ++l) {
            const long long d = s * kLanes + l;
            if (d >= A.nrows_dev) continue;
            const long long cr = A.perm ? A.perm[d] : d;
            int pos = 0;
            for (long long p = A.ptr[cr]; p < A.ptr[cr + 1]; ++p, ++pos) {
                wx[(size_t)pos * kLanes + l] = A.wx[p];
                wy[(size_t)pos * kLanes + l] = A.wy[p];
            }
        }
    }
    const size_t u0 = (size_t)O.uoff[t];
    for (int q = 0; q < nu; ++q) {
        O.ulist[u0 + q] = S.cols[q];
        O.uslot[2 * (u0 + q)] = (unsigned short)slot[q];
        O.uslot[2 * (u0 + q) + 1] = (unsigned short)(two_copies ? slot1[q] : slot[q]);
    }
    O.ulist[u0 + nu] = A.n_tot;
    O.uslot[2 * (u0 + nu)] = (unsigned short)nslots;
    O.uslot[2 * (u0 + nu) + 1] = (unsigned short)nslots;
    return nslots;
}

}  // namespace tb
}  // namespace mft
