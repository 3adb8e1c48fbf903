// mft_layout_host.inl -- host side of mft_finalize: operator rows from the caller's CSC / ELL input (transpose, summation-key
// sort) and the device layouts built from them (sliced ELL, row-pair ELL, union tiles) + the host self test of the union-tile
// format.  Pure host code; included by mft_b200.cu (uses its mft_ctx, Csr2, DevEll / DevTileR, fail(), CHECK()).
// ------------------------------------------------------------------------------------------------------
// finalize: build device layouts
// ------------------------------------------------------------------------------------------------------
static void sort_rows_by_key(Csr2 &A, const std::vector<int64_t> &keys, bool paired)
{
    if (keys.empty()) return;
    std::vector<int64_t> ord;
    std::vector<int32_t> tc;
    std::vector<double> tx, ty;
    for (int64_t r = 0; r < A.nrows; ++r) {
        const int64_t b = A.ptr[r], e = A.ptr[r + 1], len = e - b;
        bool sorted = true;
        for (int64_t p = b + 1; p < e; ++p)
            if (keys[A.col[p - 1]] > keys[A.col[p]]) {
                sorted = false;
                break;
            }
        if (sorted) continue;
        ord.resize(len);
        std::iota(ord.begin(), ord.end(), (int64_t)0);
        std::stable_sort(ord.begin(), ord.end(), [&](int64_t a, int64_t c2) { return keys[A.col[b + a]] < keys[A.col[b + c2]]; });
        tc.resize(len);
        tx.resize(len);
        if (paired) ty.resize(len);
        for (int64_t q = 0; q < len; ++q) {
            tc[q] = A.col[b + ord[q]];
            tx[q] = A.wx[b + ord[q]];
            if (paired) ty[q] = A.wy[b + ord[q]];
        }
        for (int64_t q = 0; q < len; ++q) {
            A.col[b + q] = tc[q];
            A.wx[b + q] = tx[q];
            if (paired) A.wy[b + q] = ty[q];
        }
    }
}

// rows of A^T from rows of A (entries of each output row in ascending source-row order)
static void transpose_rows(const Csr2 &A, int64_t ncols, bool paired, Csr2 &T)
{
    T.nrows = ncols;
    T.ptr.assign(ncols + 1, 0);
    for (int64_t p = 0; p < (int64_t)A.col.size(); ++p) T.ptr[A.col[p] + 1]++;
    for (int64_t i = 0; i < ncols; ++i) T.ptr[i + 1] += T.ptr[i];
    T.col.resize(A.col.size());
    T.wx.resize(A.col.size());
    if (paired) T.wy.resize(A.col.size());
    std::vector<int64_t> fill(T.ptr.begin(), T.ptr.end() - 1);
    for (int64_t r = 0; r < A.nrows; ++r)
        for (int64_t p = A.ptr[r]; p < A.ptr[r + 1]; ++p) {
            const int64_t q = fill[A.col[p]]++;
            T.col[q] = (int32_t)r;
            T.wx[q] = A.wx[p];
            if (paired) T.wy[q] = A.wy[p];
        }
}

// CSC columns -> "rows of the transpose" directly (column i of D = row i of D'), ascending row index
static void csc_to_colrows(const HostCsc &X, const HostCsc *Y, int64_t n, Csr2 &T)
{
    T.nrows = n;
    T.ptr.assign(X.colptr.begin(), X.colptr.end());
    T.col.resize(X.rowval.size());
    for (size_t p = 0; p < X.rowval.size(); ++p) T.col[p] = (int32_t)X.rowval[p];
    T.wx = X.nz;
    if (Y) T.wy = Y->nz;
}

// build sliced ELL for device rows [0, nrows_dev): device row d <- caller row perm[d]; columns remapped by iperm
static int build_ell(mft_ctx *c, const Csr2 &A, int64_t nrows_dev, bool paired, DevEll &out)
{
    const int64_t nsl = (nrows_dev + kSlice - 1) / kSlice;
    std::vector<int> off(nsl + 1, 0);
    auto caller_row = [&](int64_t d) -> int64_t { return c->have_perm ? c->perm[d] : d; };
    int64_t nnz = 0;
    for (int64_t s = 0; s < nsl; ++s) {
        int64_t w = 0;
        for (int64_t d = s * kSlice; d < std::min(nrows_dev, (s + 1) * kSlice); ++d) {
            const int64_t r = caller_row(d);
            const int64_t len = A.ptr[r + 1] - A.ptr[r];
            w = std::max(w, len);
            nnz += len;
        }
        const int64_t tot = (int64_t)off[s] + w;
        if (tot > 0x7fffffffLL / kSlice * 16) return fail(MFT_EINVAL, "operator too large for 32-bit slice offsets");
        off[s + 1] = (int)tot;
    }
    const int64_t ncols = off[nsl];
    const int colb = paired ? kColBytes2 : kColBytes1;
    std::vector<unsigned char> blob((size_t)ncols * colb + 128, 0);
    int maxw = 0;
    for (int64_t s = 0; s < nsl; ++s) {
        const int w = off[s + 1] - off[s];
        maxw = std::max(maxw, w);
        unsigned char *b = blob.data() + (size_t)off[s] * colb;
        int *idx = reinterpret_cast<int *>(b);
        double *wx = reinterpret_cast<double *>(b + (size_t)w * kSlice * 4);
        double *wy = reinterpret_cast<double *>(b + (size_t)w * kSlice * 12);
        // padding: the dummy record (index n_tot) with weight 0 -> adds an exact zero, one shared sector per request
        for (int q = 0; q < w * kSlice; ++q) idx[q] = (int)c->n_tot;
        for (int64_t d = s * kSlice; d < std::min(nrows_dev, (s + 1) * kSlice); ++d) {
            const int64_t r = caller_row(d);
            const int lane = (int)(d - s * kSlice);
            int cpos = 0;
            for (int64_t p = A.ptr[r]; p < A.ptr[r + 1]; ++p, ++cpos) {
                const int64_t j = A.col[p];
                const size_t at = (size_t)cpos * kSlice + lane;
                idx[at] = c->have_perm ? c->iperm[j] : (int)j;
                wx[at] = A.wx[p];
                if (paired) wy[at] = A.wy[p];
            }
        }
    }
    out.nslices = (int)nsl;
    out.colb = colb;
    out.maxw = maxw;
    out.ncols_total = ncols;
    out.nnz = nnz;
    CHECK(out.blob.upload(blob));
    CHECK(out.off.upload(off));
    return MFT_OK;
}

// pair-slice blobs: device rows (2l, 2l+1) share one lane; the lane walks the union of the two rows' entries in the
// reference order (ascending key), with a zero weight where a row lacks the entry
static int build_ell_pairs(mft_ctx *c, const Csr2 &A, int64_t nrows_dev, DevEll &out)
{
    const int64_t rows_per_slice = 2 * kSlice;
    const int64_t nsl = (nrows_dev + rows_per_slice - 1) / rows_per_slice;
    auto caller_row = [&](int64_t d) -> int64_t { return c->have_perm ? c->perm[d] : d; };
    auto key = [&](int32_t col) -> int64_t { return c->keys.empty() ? (int64_t)col : c->keys[col]; };
    struct Ent { int32_t col; double w[4]; };
    std::vector<std::vector<Ent>> lists((size_t)nsl * kSlice);
    std::vector<int> off(nsl + 1, 0);
    int maxw = 0;
    int64_t nnz = 0;
    for (int64_t s = 0; s < nsl; ++s) {
        int w = 0;
        for (int l = 0; l < kSlice; ++l) {
            const int64_t dA = s * rows_per_slice + 2 * l, dB = dA + 1;
            std::vector<Ent> &U = lists[s * kSlice + l];
            int64_t pa = 0, ea = 0, pb = 0, eb = 0;
            if (dA < nrows_dev) { const int64_t r = caller_row(dA); pa = A.ptr[r]; ea = A.ptr[r + 1]; }
            if (dB < nrows_dev) { const int64_t r = caller_row(dB); pb = A.ptr[r]; eb = A.ptr[r + 1]; }
            nnz += (ea - pa) + (eb - pb);
            while (pa < ea || pb < eb) {
                Ent e{};
                const bool takeA = pa < ea && (pb >= eb || key(A.col[pa]) <= key(A.col[pb]));
                const bool takeB = pb < eb && (pa >= ea || key(A.col[pb]) <= key(A.col[pa]));
                if (takeA && takeB && A.col[pa] != A.col[pb]) {
                    // equal keys on different columns cannot happen for a permutation of keys; fall back to A first
                    e.col = A.col[pa]; e.w[0] = A.wx[pa]; e.w[1] = A.wy[pa]; ++pa;
                } else if (takeA && takeB) {
                    e.col = A.col[pa]; e.w[0] = A.wx[pa]; e.w[1] = A.wy[pa]; e.w[2] = A.wx[pb]; e.w[3] = A.wy[pb]; ++pa; ++pb;
                } else if (takeA) {
                    e.col = A.col[pa]; e.w[0] = A.wx[pa]; e.w[1] = A.wy[pa]; ++pa;
                } else {
                    e.col = A.col[pb]; e.w[2] = A.wx[pb]; e.w[3] = A.wy[pb]; ++pb;
                }
                U.push_back(e);
            }
            w = std::max(w, (int)U.size());
        }
        maxw = std::max(maxw, w);
        off[s + 1] = off[s] + w;
    }
    const int64_t ncols = off[nsl];
    std::vector<unsigned char> blob((size_t)ncols * kColBytesPair + 128, 0);
    for (int64_t s = 0; s < nsl; ++s) {
        const int w = off[s + 1] - off[s];
        unsigned char *b = blob.data() + (size_t)off[s] * kColBytesPair;
        int *idx = reinterpret_cast<int *>(b);
        double *wq = reinterpret_cast<double *>(b + (size_t)w * kSlice * 4);
        for (int q = 0; q < w * kSlice; ++q) idx[q] = (int)c->n_tot;
        for (int l = 0; l < kSlice; ++l) {
            const std::vector<Ent> &U = lists[s * kSlice + l];
            for (size_t cpos = 0; cpos < U.size(); ++cpos) {
                const size_t at = cpos * kSlice + l;
                idx[at] = c->have_perm ? c->iperm[U[cpos].col] : (int)U[cpos].col;
                for (int q = 0; q < 4; ++q) wq[(size_t)q * w * kSlice + at] = U[cpos].w[q];
            }
        }
    }
    out.nslices = (int)nsl;
    out.colb = kColBytesPair;
    out.maxw = maxw;
    out.ncols_total = ncols;
    out.nnz = nnz;
    CHECK(out.blob.upload(blob));
    CHECK(out.off.upload(off));
    return MFT_OK;
}

// Union tiles (mft_tile_kernels.cuh).  Tile = kTileWarps slices; slice = 32 lanes x R rows: lane l of slice s owns device
// rows (s*32 + l)*R + r and walks the union of their entries in summation order; step word = slot | row mask << 12; the
// weights stay compact per row.  The tile's union list is sorted by device index (coalesced loads) and ends with the dummy
// record; uslot[] gives each entry its shared-memory slot.  With `colour` the slots are chosen so that points requested
// together by the 8 lanes of an LDS.128 phase fall into different 16-byte bank groups (slot mod 8) where possible:
// greedy weighted colouring of the co-request graph + two refinement sweeps.
struct HostTileR {
    int R = 0, nslices = 0, ntiles = 0, maxW = 0, maxL = 0, sstride = 0, ncopy = 1;
    int64_t nnz = 0, nsteps = 0;
    std::vector<unsigned char> blob;
    std::vector<long long> boff;
    std::vector<int> wl, uoff, ulist;
    std::vector<unsigned short> uslot;
};

// Run fn(worker) on `nthreads` host threads (the caller's thread is worker 0) and wait for all of them.
template <class F>
static void host_parallel(int nthreads, F fn)
{
    std::vector<std::thread> pool;
    pool.reserve(nthreads > 1 ? nthreads - 1 : 0);
    for (int w = 1; w < nthreads; ++w) {
        try {
            pool.emplace_back([&fn, w]() { fn(w); });
        } catch (const std::system_error &) {
            break;  // thread limit reached: the workers that did start (and this thread) share the work through the atomic cursor
        }
    }
    fn(0);
    for (auto &th : pool) th.join();
}

// host threads for the layout builders: MFT_HOST_THREADS, else the hardware concurrency (at most 64)
static int host_threads(int64_t work_items)
{
    int n = 0;
    if (const char *e = getenv("MFT_HOST_THREADS")) n = atoi(e);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    n = std::max(1, std::min(n, 64));
    return (int)std::max<int64_t>(1, std::min<int64_t>(n, work_items));
}

// state of x -> a x + c (mod 2^64) after `k` more steps (jump-ahead by repeated squaring of the affine map)
static inline uint64_t lcg_jump(uint64_t x, uint64_t k)
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

// <= 8 points, each offering two bank groups (its slot in copy 0 / copy 1): is there an assignment with all groups distinct?
// (Kuhn's augmenting paths on an 8 x 8 bipartite graph)
struct BankMatcher {
    int opt[8][2];
    int owner[8];
    int kk = 0;
    unsigned blocked = 0;   // bank groups that are taken already (the dummy record read by the padding lanes of the phase)
    bool augment(int i, unsigned &seen)
    {
        for (int o = 0; o < 2; ++o) {
            const int b = opt[i][o];
            if ((seen | blocked) & (1u << b)) continue;
            seen |= 1u << b;
            if (owner[b] < 0 || augment(owner[b], seen)) {
                owner[b] = i;
                return true;
            }
        }
        return false;
    }
    bool perfect()
    {
        for (int b = 0; b < 8; ++b) owner[b] = -1;
        for (int i = 0; i < kk; ++i) {
            unsigned seen = 0;
            if (!augment(i, seen)) return false;
        }
        return true;
    }
};

static int build_tiler_host(const mft_ctx *c, const Csr2 &A, int64_t nrows_dev, int R, bool colour, bool two_copies, HostTileR &out,
                            bool tune_copy1 = false)
{
    if (R > 2) two_copies = false;  // the copy-select bit (14) is a row-mask bit for R = 4
    const uint64_t lcg0 = 0x9e3779b97f4a7c15ULL;
    const int64_t rows_per_slice = (int64_t)kSlice * R, rows_per_tile = rows_per_slice * kTileWarps;
    const int64_t nsl = (nrows_dev + rows_per_slice - 1) / rows_per_slice;
    const int64_t ntl = (nrows_dev + rows_per_tile - 1) / rows_per_tile;
    auto caller_row = [&](int64_t d) -> int64_t { return c->have_perm ? c->perm[d] : d; };
    auto dev_col = [&](int64_t j) -> int { return c->have_perm ? c->iperm[j] : (int)j; };
    auto key = [&](int32_t col) -> int64_t { return c->keys.empty() ? (int64_t)col : c->keys[col]; };
    constexpr int NB = 8;  // 16-byte bank groups seen by one LDS.128 phase (8 lanes)
    std::vector<long long> &boff = out.boff;
    std::vector<int> &wl = out.wl, &uoff = out.uoff, &ulist = out.ulist;
    std::vector<unsigned short> &uslot = out.uslot;
    std::vector<unsigned char> &blob = out.blob;
    boff.assign(nsl, 0);
    wl.assign(2 * nsl, 0);
    uoff.assign(ntl + 1, 0);
    struct Step { int node; unsigned mask; };
    // per-thread scratch of one tile
    struct Scratch {
        std::vector<int> cols;                   // the tile's union: sorted device columns (node q = cols[q])
        std::vector<std::vector<Step>> lanes;    // step lists of the tile's lanes
        std::vector<unsigned short> adj;         // nu x nu co-request counts
        std::vector<int> slot, slot1, order, bank, deg, grp, nb_ptr, nb_idx, nb_w;
        std::vector<int> rq_ptr, rq_node, nr_ptr, nr_idx, bank1;  // tune_copy1: the tile's requests and node -> requests
        std::vector<char> rq_ok, rq_pad;
    };
    // union of the tile's stencils (sorted) + the lane step lists of its slices (R-way merge by summation key);
    // returns the number of stored entries of the tile's rows
    auto tile_lanes = [&](int64_t t, Scratch &S) -> int64_t {
        const int64_t d0 = t * rows_per_tile, d1 = std::min(nrows_dev, d0 + rows_per_tile);
        std::vector<int> &cols = S.cols;
        cols.clear();
        for (int64_t d = d0; d < d1; ++d) {
            const int64_t r = caller_row(d);
            for (int64_t p = A.ptr[r]; p < A.ptr[r + 1]; ++p) cols.push_back(dev_col(A.col[p]));
        }
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        auto node_of = [&](int j) -> int { return (int)(std::lower_bound(cols.begin(), cols.end(), j) - cols.begin()); };
        const int64_t s0 = d0 / rows_per_slice;
        const int ns_tile = (int)((d1 - d0 + rows_per_slice - 1) / rows_per_slice);
        S.lanes.resize((size_t)kTileWarps * kSlice);
        int64_t nnz = 0;
        for (int si = 0; si < ns_tile; ++si)
            for (int l = 0; l < kSlice; ++l) {
                std::vector<Step> &U = S.lanes[(size_t)si * kSlice + l];
                U.clear();
                int64_t pp[4], ee[4];
                for (int r = 0; r < R; ++r) {
                    const int64_t d = ((s0 + si) * kSlice + l) * R + r;
                    pp[r] = ee[r] = 0;
                    if (d < nrows_dev) {
                        const int64_t cr = caller_row(d);
                        pp[r] = A.ptr[cr];
                        ee[r] = A.ptr[cr + 1];
                        nnz += ee[r] - pp[r];
                    }
                }
                for (;;) {
                    int best = -1;
                    for (int r = 0; r < R; ++r)
                        if (pp[r] < ee[r] && (best < 0 || key(A.col[pp[r]]) < key(A.col[pp[best]]))) best = r;
                    if (best < 0) break;
                    const int32_t col = A.col[pp[best]];
                    Step st{node_of(dev_col(col)), 0u};
                    for (int r = 0; r < R; ++r)
                        if (pp[r] < ee[r] && A.col[pp[r]] == col) {
                            st.mask |= 1u << r;
                            ++pp[r];
                        }
                    U.push_back(st);
                }
            }
        return nnz;
    };
    // W (steps) and L (longest row) of slice si of the tile whose lanes are in S
    auto slice_wl = [&](int64_t s, int si, const Scratch &S, int &W, int &L) {
        W = 0;
        L = 0;
        for (int l = 0; l < kSlice; ++l) {
            W = std::max(W, (int)S.lanes[(size_t)si * kSlice + l].size());
            for (int r = 0; r < R; ++r) {
                const int64_t d = (s * kSlice + l) * R + r;
                if (d < nrows_dev) {
                    const int64_t cr = caller_row(d);
                    L = std::max(L, (int)(A.ptr[cr + 1] - A.ptr[cr]));
                }
            }
        }
    };
    const int nthreads = host_threads(ntl);
    constexpr int64_t kChunk = 16;  // tiles a worker claims at a time
    // ---- pass 1: sizes.  Per tile the union size, per slice (W, L): every offset of the layout follows from them ----
    std::vector<int> nu_of((size_t)ntl, 0);
    {
        std::atomic<int64_t> next{0};
        host_parallel(nthreads, [&](int) {
            Scratch S;
            for (;;) {
                const int64_t t0 = next.fetch_add(kChunk);
                if (t0 >= ntl) break;
                for (int64_t t = t0; t < std::min(ntl, t0 + kChunk); ++t) {
                    tile_lanes(t, S);
                    nu_of[(size_t)t] = (int)S.cols.size();
                    const int64_t d0 = t * rows_per_tile, d1 = std::min(nrows_dev, d0 + rows_per_tile);
                    const int64_t s0 = d0 / rows_per_slice;
                    const int ns_tile = (int)((d1 - d0 + rows_per_slice - 1) / rows_per_slice);
                    for (int si = 0; si < ns_tile; ++si) slice_wl(s0 + si, si, S, wl[2 * (s0 + si)], wl[2 * (s0 + si) + 1]);
                }
            }
        });
    }
    std::vector<uint64_t> lcg_skip((size_t)ntl + 1, 0);  // steps of the copy-1 generator consumed before tile t
    {
        int64_t usum = 0;
        for (int64_t t = 0; t < ntl; ++t) {
            usum += nu_of[(size_t)t] + 1;  // + the dummy record
            if (usum > 0x7fffffffLL) return fail(MFT_EINVAL, "union lists exceed 32-bit offsets");
            uoff[t + 1] = (int)usum;
            lcg_skip[(size_t)t + 1] = lcg_skip[(size_t)t] + (two_copies ? (uint64_t)std::max(nu_of[(size_t)t] - 1, 0) : 0);
        }
        long long at = 0;
        for (int64_t s = 0; s < nsl; ++s) {
            boff[s] = at;
            at += (long long)wl[2 * s] * kSlice * 2 + 2LL * R * wl[2 * s + 1] * kSlice * 8;
        }
        blob.assign((size_t)at + 128, 0);
        ulist.assign((size_t)usum, 0);
        uslot.assign((size_t)usum * 2, 0);
    }
    // ---- pass 2: slots (bank colouring, second copy) and the slices, written in place ----
    struct Totals {
        int max_slot = 0, maxW = 0, maxL = 0, err_slots = 0;
        int64_t nnz = 0, nsteps = 0;
    };
    std::vector<Totals> totals((size_t)nthreads);
    std::atomic<int64_t> next{0};
    std::atomic<bool> failed{false};
    host_parallel(nthreads, [&](int worker) {
        Scratch S;
        Totals &T = totals[(size_t)worker];
        std::vector<int> &slot = S.slot, &slot1 = S.slot1, &order = S.order, &bank = S.bank, &deg = S.deg, &grp = S.grp, &nb_ptr = S.nb_ptr,
                         &nb_idx = S.nb_idx, &nb_w = S.nb_w;
        std::vector<unsigned short> &adj = S.adj;
        for (;;) {
            const int64_t tc = next.fetch_add(kChunk);
            if (tc >= ntl || failed.load()) break;
            for (int64_t t = tc; t < std::min(ntl, tc + kChunk); ++t) {
                const int64_t d0 = t * rows_per_tile, d1 = std::min(nrows_dev, d0 + rows_per_tile);
                T.nnz += tile_lanes(t, S);
                const std::vector<int> &cols = S.cols;
                const std::vector<std::vector<Step>> &lanes = S.lanes;
                const int nu = (int)cols.size();
                const int64_t s0 = d0 / rows_per_slice;
                const int ns_tile = (int)((d1 - d0 + rows_per_slice - 1) / rows_per_slice);
                // slots
                slot.assign(nu, 0);
                int nslots = nu;
                if (!colour || nu <= NB) {
                    for (int q = 0; q < nu; ++q) slot[q] = q;
                } else {
                    adj.assign((size_t)nu * nu, 0);
                    for (int si = 0; si < ns_tile; ++si) {
                        size_t W = 0;
                        for (int l = 0; l < kSlice; ++l) W = std::max(W, lanes[(size_t)si * kSlice + l].size());
                        for (size_t cpos = 0; cpos < W; ++cpos)
                            for (int ph = 0; ph < kSlice / NB; ++ph) {
                                grp.clear();
                                for (int l = ph * NB; l < (ph + 1) * NB; ++l) {
                                    const std::vector<Step> &U = lanes[(size_t)si * kSlice + l];
                                    if (cpos < U.size() && std::find(grp.begin(), grp.end(), U[cpos].node) == grp.end()) grp.push_back(U[cpos].node);
                                }
                                for (size_t x = 0; x < grp.size(); ++x)
                                    for (size_t y = 0; y < grp.size(); ++y)
                                        if (x != y) {
                                            unsigned short &e = adj[(size_t)grp[x] * nu + grp[y]];
                                            if (e < 0xffff) ++e;
                                        }
                            }
                    }
                    // adjacency lists (node, weight) from the dense counts
                    deg.assign(nu, 0);
                    nb_ptr.assign(nu + 1, 0);
                    nb_idx.clear();
                    nb_w.clear();
                    for (int a = 0; a < nu; ++a) {
                        const unsigned short *row = &adj[(size_t)a * nu];
                        int sum = 0;
                        for (int b = 0; b < nu; ++b)
                            if (row[b]) {
                                sum += row[b];
                                nb_idx.push_back(b);
                                nb_w.push_back(row[b]);
                            }
                        deg[a] = sum;
                        nb_ptr[a + 1] = (int)nb_idx.size();
                    }
                    order.resize(nu);
                    std::iota(order.begin(), order.end(), 0);
                    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return deg[a] > deg[b]; });
                    bank.assign(nu, -1);
                    int fill[NB] = {0};
                    auto choose = [&](int a) {
                        long long cost[NB] = {0};
                        for (int q = nb_ptr[a]; q < nb_ptr[a + 1]; ++q)
                            if (bank[nb_idx[q]] >= 0) cost[bank[nb_idx[q]]] += nb_w[q];
                        int bestb = 0;
                        for (int k = 1; k < NB; ++k)   // ties -> emptiest bank group (keeps the slot count low)
                            if (cost[k] < cost[bestb] || (cost[k] == cost[bestb] && fill[k] < fill[bestb])) bestb = k;
                        return bestb;
                    };
                    for (int a : order) {
                        bank[a] = choose(a);
                        ++fill[bank[a]];
                    }
                    for (int pass = 0; pass < 2; ++pass)
                        for (int a : order) {
                            --fill[bank[a]];
                            bank[a] = -1;
                            bank[a] = choose(a);
                            ++fill[bank[a]];
                        }
                    if (tune_copy1 && two_copies) {   // (alone, the octet constraint would undo the colouring: 1.68 -> 1.95 read conflicts)
                        // phase 1 stores entry q of the union list from thread q: the 8 lanes of an STS.128 phase hold an aligned
                        // octet of the list.  Keep the colouring where it already gives the octet 8 different bank groups and move
                        // the duplicates to the free groups: conflict-free stores, slot = octet * 8 + group (compact).
                        for (int o = 0; o < nu; o += NB) {
                            const int oe = std::min(nu, o + NB);
                            unsigned used = 0;
                            int dup[NB], nd = 0;
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
                            slot[q] = (q / NB) * NB + bank[q];
                            nslots = std::max(nslots, slot[q] + 1);
                        }
                    } else {
                        int level[NB] = {0};
                        nslots = 0;
                        for (int q = 0; q < nu; ++q) {
                            slot[q] = level[bank[q]]++ * NB + bank[q];
                            nslots = std::max(nslots, slot[q] + 1);
                        }
                    }
                }
                // second copy of the tile's records under an independent (pseudo-random) bank assignment: each distinct point a
                // phase requests may then be read from either copy ("two choices"), which the emit loop below exploits.
                // One generator runs through the tiles in order; a worker jumps it to its tile's position.
                if (two_copies) {
                    uint64_t lcg = lcg_jump(lcg0, lcg_skip[(size_t)t]);
                    slot1.resize(nu);   // a random permutation of 0..nu-1: balanced bank groups, independent of copy 0
                    std::iota(slot1.begin(), slot1.end(), 0);
                    for (int q = nu - 1; q > 0; --q) {
                        lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
                        std::swap(slot1[q], slot1[(int)((lcg >> 33) % (uint64_t)(q + 1))]);
                    }
                    nslots = std::max(nslots, nu);
                    if (tune_copy1 && nu > NB) {
                        // Local search on the bank groups of copy 1 (tools/tile_sim.py model: LDS.128 conflict degree 1.25 -> 1.01).
                        // A request = the distinct points the 8 lanes of one LDS.128 phase ask for; it is conflict-free iff its
                        // points can be matched to distinct bank groups, each point offering the group of its copy-0 slot and of
                        // its copy-1 slot (bipartite matching, <= 8 x 8).  Two sweeps over the points: move a point of copy 1 to
                        // the group that leaves the fewest of its requests unmatched; groups stay balanced (cap per group).
                        std::vector<int> &rq_ptr = S.rq_ptr, &rq_node = S.rq_node, &nr_ptr = S.nr_ptr, &nr_idx = S.nr_idx, &bank1 = S.bank1;
                        std::vector<char> &rq_ok = S.rq_ok;
                        // In this mode the dummy record (read by the padding lanes of a step) sits in slot roundup8(nu) of both copies,
                        // i.e. in bank group 0: a phase with padding lanes has that group taken.
                        std::vector<char> &rq_pad = S.rq_pad;
                        rq_ptr.assign(1, 0);
                        rq_node.clear();
                        rq_pad.clear();
                        for (int si = 0; si < ns_tile; ++si) {
                            size_t W = 0;
                            for (int l = 0; l < kSlice; ++l) W = std::max(W, lanes[(size_t)si * kSlice + l].size());
                            for (size_t cpos = 0; cpos < W; ++cpos)
                                for (int ph = 0; ph < kSlice / NB; ++ph) {
                                    grp.clear();
                                    bool pad = false;
                                    for (int l = ph * NB; l < (ph + 1) * NB; ++l) {
                                        const std::vector<Step> &U = lanes[(size_t)si * kSlice + l];
                                        if (cpos >= U.size()) pad = true;
                                        else if (std::find(grp.begin(), grp.end(), U[cpos].node) == grp.end()) grp.push_back(U[cpos].node);
                                    }
                                    if (grp.size() + (pad ? 1 : 0) < 2) continue;
                                    rq_node.insert(rq_node.end(), grp.begin(), grp.end());
                                    rq_ptr.push_back((int)rq_node.size());
                                    rq_pad.push_back(pad ? 1 : 0);
                                }
                        }
                        const int nreq = (int)rq_ptr.size() - 1;
                        nr_ptr.assign(nu + 1, 0);
                        for (int x : rq_node) nr_ptr[x + 1]++;
                        for (int q = 0; q < nu; ++q) nr_ptr[q + 1] += nr_ptr[q];
                        nr_idx.assign(rq_node.size(), 0);
                        {
                            std::vector<int> cur(nr_ptr.begin(), nr_ptr.end() - 1);
                            for (int r = 0; r < nreq; ++r)
                                for (int e = rq_ptr[r]; e < rq_ptr[r + 1]; ++e) nr_idx[cur[rq_node[e]]++] = r;
                        }
                        // copy 1 starts from a random permutation of the bank groups inside every octet of the union list (octets
                        // stay permutations: conflict-free phase-1 stores, compact slots) ...
                        bank1.resize(nu);
                        for (int o = 0; o < nu; o += NB) {
                            const int oe = std::min(nu, o + NB);
                            for (int q = o; q < oe; ++q) {
                                int rank = 0;
                                for (int q2 = o; q2 < oe; ++q2) rank += slot1[q2] < slot1[q] ? 1 : 0;   // slot1: random permutation
                                bank1[q] = rank;
                            }
                        }
                        constexpr int kSweeps = 2;   // more sweeps do not lower the conflict degree further (measured)
                        auto matched = [&](int r) -> bool {   // can the request's points take distinct bank groups?
                            BankMatcher M;
                            const int e0 = rq_ptr[r];
                            M.kk = rq_ptr[r + 1] - e0;
                            M.blocked = rq_pad[r] ? 1u : 0u;
                            for (int i = 0; i < M.kk; ++i) {
                                M.opt[i][0] = slot[rq_node[e0 + i]] % NB;
                                M.opt[i][1] = bank1[rq_node[e0 + i]];
                            }
                            return M.perfect();
                        };
                        rq_ok.assign(nreq, 0);
                        for (int r = 0; r < nreq; ++r) rq_ok[r] = matched(r) ? 1 : 0;
                        std::vector<int> &stamp = S.order;   // request -> last point that counted it (scratch)
                        stamp.assign(nreq, -1);
                        // ... and is improved by swaps inside an octet: the swap that leaves the fewest requests of the two points
                        // unmatched
                        for (int sweep = 0; sweep < kSweeps; ++sweep)
                            for (int q = 0; q < nu; ++q) {
                                int bad_q = 0;
                                for (int e = nr_ptr[q]; e < nr_ptr[q + 1]; ++e) bad_q += rq_ok[nr_idx[e]] ? 0 : 1;
                                if (bad_q == 0) continue;
                                const int o = (q / NB) * NB, oe = std::min(nu, o + NB);
                                int best2 = -1, best_gain = 0;
                                for (int q2 = o; q2 < oe; ++q2) {
                                    if (q2 == q) continue;
                                    const int mark = q * NB + (q2 - o);
                                    int before = bad_q, after = 0;
                                    for (int e = nr_ptr[q]; e < nr_ptr[q + 1]; ++e) stamp[nr_idx[e]] = mark;
                                    for (int e = nr_ptr[q2]; e < nr_ptr[q2 + 1]; ++e)
                                        if (stamp[nr_idx[e]] != mark) before += rq_ok[nr_idx[e]] ? 0 : 1;
                                    std::swap(bank1[q], bank1[q2]);
                                    for (int e = nr_ptr[q]; e < nr_ptr[q + 1] && before - after > best_gain; ++e) after += matched(nr_idx[e]) ? 0 : 1;
                                    for (int e = nr_ptr[q2]; e < nr_ptr[q2 + 1] && before - after > best_gain; ++e)
                                        if (stamp[nr_idx[e]] != mark) after += matched(nr_idx[e]) ? 0 : 1;
                                    std::swap(bank1[q], bank1[q2]);
                                    if (before - after > best_gain) {
                                        best_gain = before - after;
                                        best2 = q2;
                                    }
                                }
                                if (best2 >= 0) {
                                    std::swap(bank1[q], bank1[best2]);
                                    for (int e = nr_ptr[q]; e < nr_ptr[q + 1]; ++e) rq_ok[nr_idx[e]] = matched(nr_idx[e]) ? 1 : 0;
                                    for (int e = nr_ptr[best2]; e < nr_ptr[best2 + 1]; ++e) rq_ok[nr_idx[e]] = matched(nr_idx[e]) ? 1 : 0;
                                }
                            }
                        for (int q = 0; q < nu; ++q) slot1[q] = (q / NB) * NB + bank1[q];
                        nslots = (nu + NB - 1) / NB * NB;   // dummy slot: bank group 0 in both copies
                    }
                }
                if (nslots + 1 > 4095) {
                    T.err_slots = nslots;
                    failed.store(true);
                    break;
                }
                T.max_slot = std::max(T.max_slot, nslots);  // the dummy record takes slot `nslots`
                // emit the slices
                for (int si = 0; si < ns_tile; ++si) {
                    const int64_t s = s0 + si;
                    const int W = wl[2 * s], L = wl[2 * s + 1];
                    T.maxW = std::max(T.maxW, W);
                    T.maxL = std::max(T.maxL, L);
                    T.nsteps += W;
                    const size_t word_bytes = (size_t)W * kSlice * 2, wblk = (size_t)L * kSlice * 8;
                    const size_t at0 = (size_t)boff[s];
                    unsigned short *word = reinterpret_cast<unsigned short *>(blob.data() + at0);
                    for (int q = 0; q < W * kSlice; ++q) word[q] = (unsigned short)nslots;  // dummy slot, empty mask
                    for (int l = 0; l < kSlice; ++l) {
                        const std::vector<Step> &U = lanes[(size_t)si * kSlice + l];
                        for (size_t cpos = 0; cpos < U.size(); ++cpos)
                            word[cpos * kSlice + l] = (unsigned short)(slot[U[cpos].node] | (U[cpos].mask << 12));
                    }
                    if (two_copies) {
                        // per step and LDS.128 phase (8 lanes): the copy of each distinct point that minimises the largest number
                        // of distinct addresses in one bank group (exhaustive over <= 2^8 choices)
                        for (int cpos = 0; cpos < W; ++cpos)
                            for (int ph = 0; ph < kSlice / NB; ++ph) {
                                grp.clear();
                                for (int l = ph * NB; l < (ph + 1) * NB; ++l) {
                                    const std::vector<Step> &U = lanes[(size_t)si * kSlice + l];
                                    if ((size_t)cpos < U.size() && std::find(grp.begin(), grp.end(), U[cpos].node) == grp.end()) grp.push_back(U[cpos].node);
                                }
                                const int kk = (int)grp.size();
                                bool pad = false;   // tuned layout: padding lanes of this phase read the dummy record in bank group 0
                                if (tune_copy1 && nu > NB)
                                    for (int l = ph * NB; l < (ph + 1) * NB; ++l) pad |= (size_t)cpos >= lanes[(size_t)si * kSlice + l].size();
                                if (kk + (pad ? 1 : 0) < 2) continue;
                                int best_bits = 0, best_max = 99;
                                for (int bits = 0; bits < (1 << kk) && best_max > 1; ++bits) {
                                    int cnt[NB] = {pad ? 1 : 0}, mx = 0;
                                    for (int q = 0; q < kk; ++q) mx = std::max(mx, ++cnt[((bits >> q) & 1 ? slot1[grp[q]] : slot[grp[q]]) % NB]);
                                    if (mx < best_max) {
                                        best_max = mx;
                                        best_bits = bits;
                                    }
                                }
                                for (int l = ph * NB; l < (ph + 1) * NB; ++l) {
                                    const std::vector<Step> &U = lanes[(size_t)si * kSlice + l];
                                    if ((size_t)cpos >= U.size()) continue;
                                    const int q = (int)(std::find(grp.begin(), grp.end(), U[cpos].node) - grp.begin());
                                    if ((best_bits >> q) & 1)
                                        word[(size_t)cpos * kSlice + l] = (unsigned short)(slot1[U[cpos].node] | (U[cpos].mask << 12) | 0x4000u);
                                }
                            }
                    }
                    for (int l = 0; l < kSlice; ++l) {
                        for (int r = 0; r < R; ++r) {
                            const int64_t d = (s * kSlice + l) * R + r;
                            if (d >= nrows_dev) continue;
                            const int64_t cr = caller_row(d);
                            double *wx = reinterpret_cast<double *>(blob.data() + at0 + word_bytes + (size_t)r * wblk);
                            double *wy = reinterpret_cast<double *>(blob.data() + at0 + word_bytes + (size_t)(R + r) * wblk);
                            int pos = 0;
                            for (int64_t p = A.ptr[cr]; p < A.ptr[cr + 1]; ++p, ++pos) {
                                wx[(size_t)pos * kSlice + l] = A.wx[p];
                                wy[(size_t)pos * kSlice + l] = A.wy[p];
                            }
                        }
                    }
                }
                // the tile's union list and slot table; last entry = the dummy record (a finite state in u, zeros in g)
                const size_t u0 = (size_t)uoff[t];
                for (int q = 0; q < nu; ++q) {
                    ulist[u0 + q] = cols[q];
                    uslot[2 * (u0 + q)] = (unsigned short)slot[q];
                    uslot[2 * (u0 + q) + 1] = (unsigned short)(two_copies ? slot1[q] : slot[q]);
                }
                ulist[u0 + nu] = (int)c->n_tot;
                uslot[2 * (u0 + nu)] = (unsigned short)nslots;
                uslot[2 * (u0 + nu) + 1] = (unsigned short)nslots;
            }
        }
    });
    int max_slot = 0, maxW = 0, maxL = 0;
    int64_t nnz = 0, nsteps = 0;
    for (const Totals &T : totals) {
        if (T.err_slots) return fail(MFT_ENOTSUP, "union tile: %d slots in one block exceed the 12-bit slot field", T.err_slots);
        max_slot = std::max(max_slot, T.max_slot);
        maxW = std::max(maxW, T.maxW);
        maxL = std::max(maxL, T.maxL);
        nnz += T.nnz;
        nsteps += T.nsteps;
    }
    out.R = R;
    out.nslices = (int)nsl;
    out.ntiles = (int)ntl;
    out.maxW = maxW;
    out.maxL = maxL;
    out.sstride = ((max_slot + 1 + 7) / 8) * 8;
    out.nnz = nnz;
    out.nsteps = nsteps;
    out.ncopy = two_copies ? 2 : 1;
    if (ulist.empty()) {
        ulist.push_back(0);
        uslot.push_back(0);
        uslot.push_back(0);
    }
    return MFT_OK;
}

static int build_tiler_device(mft_ctx *c, const Csr2 &A, int64_t nrows_dev, bool colour, bool two_copies, bool tune, DevTileR &out);

static int build_tiler(mft_ctx *c, const Csr2 &A, int64_t nrows_dev, int R, bool colour, bool two_copies, DevTileR &out)
{
    // default tiles (one row per thread): laid out on the device (mft_layout_device.inl; same bytes as the host builder below)
    if (c->layout_device && R == 1 && nrows_dev > 0) return build_tiler_device(c, A, nrows_dev, colour, two_copies, (c->tile & 16) != 0, out);
    const auto t_host0 = std::chrono::steady_clock::now();
    HostTileR h;
    CHECK(build_tiler_host(c, A, nrows_dev, R, colour, two_copies, h, (c->tile & 16) != 0));
    out.ncopy = h.ncopy;
    out.R = h.R;
    out.nslices = h.nslices;
    out.ntiles = h.ntiles;
    out.maxW = h.maxW;
    out.maxL = h.maxL;
    out.sstride = h.sstride;
    out.nnz = h.nnz;
    out.nunion = (int64_t)h.ulist.size();
    out.nsteps = h.nsteps;
    CHECK(out.blob.upload(h.blob));
    CHECK(out.boff.upload(h.boff));
    CHECK(out.wl.upload(h.wl));
    CHECK(out.uoff.upload(h.uoff));
    CHECK(out.ulist.upload(h.ulist));
    CHECK(out.uslot.upload(h.uslot));
    if (getenv("MFT_TRACE"))
        fprintf(stderr, "[mft] union-tile layout on %d host threads: %d tiles, R = %d: %.3f s\n", host_threads(h.ntiles), h.ntiles, R,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_host0).count());
    return MFT_OK;
}

// Host-only self test of the union-tile format (no CUDA calls): a random banded operator is laid out by
// build_tiler_host, then the kernels' walk (step words, row masks, per-row weight cursors, slot table) is replayed on
// the CPU and compared bit for bit with the plain row sums in summation order.  Returns 0 when identical.
// build the union-tile layout of A for the (stack) ctx, replay the kernels' walk on the CPU and compare with the plain row sums
static int tile_selftest_run(mft_ctx &ctx, const Csr2 &A, int64_t n, int k, int R, int layout, int with_perm, uint64_t st, double *stats4,
                             int nstats = 4)
{
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(st >> 33); };
    HostTileR h;
    const auto t_build0 = std::chrono::steady_clock::now();
    CHECK(build_tiler_host(&ctx, A, ctx.n_local, R, (layout & 1) != 0, (layout & 2) != 0, h, (layout & 4) != 0));
    if (getenv("MFT_TRACE")) {   // build time + a checksum of the whole layout (compare builds / thread counts)
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
        fprintf(stderr, "[mft] tile layout: n=%lld k=%d R=%d layout=%d perm=%d  build %.3f s  fnv %016llx\n", (long long)n, k, R, layout, with_perm,
                std::chrono::duration<double>(std::chrono::steady_clock::now() - t_build0).count(), (unsigned long long)fnv);
    }
    std::vector<double> x((size_t)n + 1);
    for (auto &v : x) v = (double)(int)(rnd() % 4001 - 2000) / 128.0;   // indexed by DEVICE column; x[n] = dummy record
    auto caller_row = [&](int64_t d) -> int64_t { return ctx.have_perm ? ctx.perm[d] : d; };
    auto dev_col = [&](int64_t j) -> int64_t { return ctx.have_perm ? ctx.iperm[j] : j; };
    int64_t bad = 0, conflicts = 0, phases = 0;
    std::vector<double> smem((size_t)h.sstride * 2);
    for (int t = 0; t < h.ntiles; ++t) {
        std::fill(smem.begin(), smem.end(), std::nan(""));
        const int u0 = h.uoff[t], nu = h.uoff[t + 1] - u0;
        for (int q = 0; q < nu; ++q) {
            if (h.uslot[2 * (u0 + q)] >= h.sstride || h.uslot[2 * (u0 + q) + 1] >= h.sstride) return fail(MFT_EINVAL, "selftest: slot beyond sstride");
            if (q > 0 && q < nu - 1 && h.ulist[u0 + q] <= h.ulist[u0 + q - 1]) return fail(MFT_EINVAL, "selftest: union list not ascending");
            smem[h.uslot[2 * (u0 + q)]] = x[h.ulist[u0 + q]];
            if (h.ncopy == 2) smem[h.sstride + h.uslot[2 * (u0 + q) + 1]] = x[h.ulist[u0 + q]];
        }
        if (h.ulist[u0 + nu - 1] != n) return fail(MFT_EINVAL, "selftest: tile list does not end with the dummy record");
        const uint32_t dword = h.uslot[2 * (u0 + nu - 1)];
        for (int w = 0; w < kTileWarps; ++w) {
            const int64_t s = (int64_t)t * kTileWarps + w;
            if (s >= h.nslices) break;
            const int W = h.wl[2 * s], L = h.wl[2 * s + 1];
            const unsigned char *src = h.blob.data() + h.boff[s];
            const unsigned short *word = reinterpret_cast<const unsigned short *>(src);
            const double *wx = reinterpret_cast<const double *>(src + (size_t)W * kSlice * 2);
            const double *wy = wx + (size_t)R * L * kSlice;
            for (int c0 = 0; c0 < W; ++c0)
                for (int ph = 0; ph < 4; ++ph) {   // bank-conflict degree of this LDS.128 phase
                    int cnt[8] = {0}, seen[8], ns = 0;
                    for (int l = ph * 8; l < ph * 8 + 8; ++l) {
                        const int sl = word[c0 * kSlice + l] & 0x4fff;   // slot + copy bit: distinct addresses
                        bool dup = false;
                        for (int q = 0; q < ns; ++q) dup |= seen[q] == sl;
                        if (!dup) { seen[ns++] = sl; ++cnt[(sl & 0xfff) % 8]; }
                    }
                    conflicts += *std::max_element(cnt, cnt + 8);
                    ++phases;
                }
            for (int l = 0; l < kSlice; ++l) {
                double ax[4] = {0, 0, 0, 0}, ay[4] = {0, 0, 0, 0};
                int pos[4] = {0, 0, 0, 0};
                for (int c0 = 0; c0 < W + 3; ++c0) {   // + batch tail steps
                    const uint32_t wd = c0 < W ? word[c0 * kSlice + l] : dword;
                    const double xv = smem[(wd & 0xfff) + ((R <= 2 && (wd & 0x4000u)) ? (size_t)h.sstride : 0)];
                    for (int r = 0; r < R; ++r) {
                        const bool m = (wd >> (12 + r)) & 1u;
                        const double w1 = m ? wx[((size_t)r * L + pos[r]) * kSlice + l] : 0.0;
                        const double w2 = m ? wy[((size_t)r * L + pos[r]) * kSlice + l] : 0.0;
                        pos[r] += m;
                        ax[r] = ax[r] + w1 * xv;
                        ay[r] = ay[r] + w2 * xv;
                    }
                }
                for (int r = 0; r < R; ++r) {
                    const int64_t d = (s * kSlice + l) * R + r;
                    if (d >= ctx.n_local) continue;
                    const int64_t cr = caller_row(d);
                    double rx = 0.0, ry = 0.0;
                    for (int64_t p = A.ptr[cr]; p < A.ptr[cr + 1]; ++p) {
                        rx = rx + A.wx[p] * x[dev_col(A.col[p])];
                        ry = ry + A.wy[p] * x[dev_col(A.col[p])];
                    }
                    if (!(rx == ax[r] && ry == ay[r]) || pos[r] != (int)(A.ptr[cr + 1] - A.ptr[cr])) ++bad;
                }
            }
        }
    }
    if (stats4) {
        stats4[0] = phases ? (double)conflicts / (double)phases : 0.0;   // mean LDS.128 conflict degree
        stats4[1] = (double)h.nsteps * kSlice / (double)std::max<int64_t>(1, ctx.n_local);  // union steps per row
        stats4[2] = (double)h.ulist.size() / (double)std::max<int64_t>(1, ctx.n_local);    // union entries per row
        stats4[3] = (double)h.sstride;
        if (nstats >= 6) {
            // phase-1 stores: thread q writes entry q of the tile's union list; the 8 lanes of an STS.128 phase hold an aligned
            // octet of the list -> conflict degree = largest number of entries of the octet in one bank group, per copy
            double sum[2] = {0.0, 0.0};
            int64_t noct = 0;
            for (int t = 0; t < h.ntiles; ++t) {
                const int u0 = h.uoff[t], nu = h.uoff[t + 1] - u0;
                for (int o = 0; o < nu; o += 8, ++noct)
                    for (int cpy = 0; cpy < 2; ++cpy) {
                        int cnt[8] = {0}, mx = 0;
                        for (int q = o; q < std::min(nu, o + 8); ++q) mx = std::max(mx, ++cnt[h.uslot[2 * (u0 + q) + cpy] % 8]);
                        sum[cpy] += mx;
                    }
            }
            stats4[4] = noct ? sum[0] / (double)noct : 0.0;
            stats4[5] = noct ? sum[1] / (double)noct : 0.0;
        }
    }
    if (bad) return fail(MFT_EINVAL, "mft_debug_tile_selftest: %lld rows differ", (long long)bad);
    return MFT_OK;
}

// the random ragged banded operator of the self tests (rows = the first n - n/7 points; optional window-shuffled permutation with
// descending summation keys); returns the generator state for the caller's own draws
static uint64_t selftest_operator(mft_ctx &ctx, Csr2 &A, int64_t n, int k, int with_perm, unsigned seed)
{
    ctx.n_local = n - n / 7;  // some trailing "halo" columns without rows
    ctx.n_halo = n - ctx.n_local;
    ctx.n_tot = n;
    ctx.V = 4;
    uint64_t st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(st >> 33); };
    if (with_perm) {
        ctx.have_perm = true;
        ctx.perm.resize(n);
        std::iota(ctx.perm.begin(), ctx.perm.end(), 0);
        // shuffle inside windows so that locality survives (device rows near each other stay near)
        for (int64_t b = 0; b < ctx.n_local; b += 64)
            for (int64_t i = std::min(ctx.n_local, b + 64) - 1; i > b; --i) std::swap(ctx.perm[i], ctx.perm[b + rnd() % (i - b + 1)]);
        ctx.iperm.resize(n);
        for (int64_t d = 0; d < n; ++d) ctx.iperm[ctx.perm[d]] = (int32_t)d;
        ctx.keys.resize(n);
        for (int64_t i = 0; i < n; ++i) ctx.keys[i] = (int64_t)(n - 1 - i) * 3;  // descending keys: order != column order
    }
    A.nrows = n;
    A.ptr.assign(n + 1, 0);
    for (int64_t r = 0; r < n; ++r) {
        const int len = r < ctx.n_local ? std::max(1, k - (int)(rnd() % 4)) : 0;   // ragged rows
        std::vector<int32_t> cs;
        while ((int)cs.size() < len) {
            const int64_t j = std::min<int64_t>(n - 1, std::max<int64_t>(0, r + (int64_t)(rnd() % (6 * k)) - 3 * k));
            if (std::find(cs.begin(), cs.end(), (int32_t)j) == cs.end()) cs.push_back((int32_t)j);
        }
        auto keyf = [&](int32_t col) { return ctx.keys.empty() ? (int64_t)col : ctx.keys[col]; };
        std::sort(cs.begin(), cs.end(), [&](int32_t a, int32_t b) { return keyf(a) < keyf(b); });
        for (int32_t j : cs) {
            A.col.push_back(j);
            A.wx.push_back((double)(int)(rnd() % 2001 - 1000) / 64.0);
            A.wy.push_back((double)(int)(rnd() % 2001 - 1000) / 32.0);
        }
        A.ptr[r + 1] = (int64_t)A.col.size();
    }
    return st;
}

extern "C" int mft_debug_tile_selftest(int64_t n, int k, int R, int layout, int with_perm, unsigned seed, double *stats4)
{
    if (n <= 0 || k <= 0 || k > n || (R != 1 && R != 2 && R != 4)) return fail(MFT_EINVAL, "mft_debug_tile_selftest: bad arguments");
    NvtxRange range("tile layout selftest");
    mft_ctx ctx;
    Csr2 A;
    const uint64_t st = selftest_operator(ctx, A, n, k, with_perm, seed);
    return tile_selftest_run(ctx, A, n, k, R, layout, with_perm, st, stats4);
}

// the same self test on a caller-supplied sparsity (e.g. the kNN table of a real cloud or its transpose): rows = n_rows stencils
// over n columns (0-based CSR, columns of a row in summation order = as given); weights are pseudo-random dyadic numbers
static int selftest_operator_csr(mft_ctx &ctx, Csr2 &A, int64_t n, int64_t n_rows, const int64_t *rowptr, const int32_t *col, unsigned seed,
                                 uint64_t &st_out, int &kmax_out)
{
    ctx.n_local = n_rows;
    ctx.n_halo = n - n_rows;
    ctx.n_tot = n;
    ctx.V = 4;
    uint64_t st = seed * 6364136223846793005ULL + 1442695040888963407ULL;
    auto rnd = [&]() { st = st * 6364136223846793005ULL + 1442695040888963407ULL; return (uint32_t)(st >> 33); };
    A.nrows = n;
    A.ptr.assign(n + 1, rowptr[n_rows]);
    int kmax = 1;
    for (int64_t r = 0; r <= n_rows; ++r) A.ptr[r] = rowptr[r];
    const int64_t nnz = rowptr[n_rows];
    A.col.resize(nnz);
    A.wx.resize(nnz);
    A.wy.resize(nnz);
    for (int64_t r = 0; r < n_rows; ++r) kmax = std::max<int>(kmax, (int)(rowptr[r + 1] - rowptr[r]));
    for (int64_t p = 0; p < nnz; ++p) {
        if (col[p] < 0 || col[p] >= n) return fail(MFT_EINVAL, "mft_debug_tile_selftest_csr: column out of range");
        A.col[p] = col[p];
        A.wx[p] = (double)(int)(rnd() % 2001 - 1000) / 64.0;
        A.wy[p] = (double)(int)(rnd() % 2001 - 1000) / 32.0;
    }
    st_out = st;
    kmax_out = kmax;
    return MFT_OK;
}

extern "C" int mft_debug_tile_selftest_csr(int64_t n, int64_t n_rows, const int64_t *rowptr, const int32_t *col, int R, int layout,
                                           unsigned seed, double *stats6)
{
    if (n <= 0 || n_rows < 0 || n_rows > n || !rowptr || !col || (R != 1 && R != 2 && R != 4))
        return fail(MFT_EINVAL, "mft_debug_tile_selftest_csr: bad arguments");
    mft_ctx ctx;
    Csr2 A;
    uint64_t st = 0;
    int kmax = 1;
    CHECK(selftest_operator_csr(ctx, A, n, n_rows, rowptr, col, seed, st, kmax));
    return tile_selftest_run(ctx, A, n, kmax, R, layout, 0, st, stats6, 6);
}
