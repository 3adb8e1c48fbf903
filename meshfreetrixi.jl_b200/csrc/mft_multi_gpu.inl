// mft_multi_gpu.inl -- multi-GPU plumbing: NCCL communicator (loaded at run time), halo plan, peer-memory (CUDA IPC) exchange
// bootstrap, global norms across ranks.  Included at the end of mft_b200.cu.
// ------------------------------------------------------------------------------------------------------
// multi-GPU: NCCL (loaded at run time so single-GPU users do not need libnccl)
// ------------------------------------------------------------------------------------------------------
extern "C" int mft_nccl_unique_id(void *id128)
{
    if (!id128) return fail(MFT_EINVAL, "mft_nccl_unique_id: NULL");
    NcclApi *N = nccl_api();
    if (!N) return fail(MFT_ENCCL, "NCCL could not be loaded: %s", nccl_load_error());
    if (N->getUniqueId(id128) != 0) return fail(MFT_ENCCL, "ncclGetUniqueId failed");
    return MFT_OK;
}

extern "C" int mft_comm_init(mft_ctx *c, int nranks, int rank, const void *id128)
{
    NEED_CTX(c);
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(MFT_EINVAL, "mft_comm_init: bad arguments");
    NcclApi *N = nccl_api();
    if (!N) return fail(MFT_ENCCL, "NCCL could not be loaded: %s", nccl_load_error());
    c->nccl = N;
    if (N->commInitRank(&c->comm, nranks, id128, rank) != 0) return fail(MFT_ENCCL, "ncclCommInitRank failed");
    c->nranks = nranks;
    c->rank = rank;
    CHECK(c->gather_buf.alloc((int64_t)2 * nranks * c->V + 4 * c->V));
    // global point count (ndofs of the parallel domain, parallel_rbfsolver.jl:10-13) = sum of the owned counts
    double nl = (double)c->n_local, ng = 0.0;
    CU(cudaMemcpy(c->gather_buf.p, &nl, sizeof(double), cudaMemcpyHostToDevice));
    if (N->allReduce(c->gather_buf.p, c->gather_buf.p + 1, 1, NCCL_DOUBLE, NCCL_SUM, c->comm, c->stream) != 0)
        return fail(MFT_ENCCL, "ncclAllReduce failed: %s", N->lastError(c->comm));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(&ng, c->gather_buf.p + 1, sizeof(double), cudaMemcpyDeviceToHost));
    c->n_global = (int64_t)(ng + 0.5);
    return MFT_OK;
}

extern "C" int mft_set_halo(mft_ctx *c, int npeers, const int *peers, const int64_t *send_off, const int64_t *send_idx1,
                            const int64_t *recv_count)
{
    NEED_CTX(c);
    if (npeers < 0 || (npeers > 0 && (!peers || !send_off || !recv_count))) return fail(MFT_EINVAL, "mft_set_halo: bad arguments");
    c->peers.assign(peers, peers + npeers);
    c->send_off.assign(send_off, send_off + npeers + 1);
    c->recv_off.assign(npeers + 1, 0);
    for (int p = 0; p < npeers; ++p) c->recv_off[p + 1] = c->recv_off[p] + recv_count[p];
    if (c->recv_off[npeers] != c->n_halo) return fail(MFT_EINVAL, "mft_set_halo: receive counts sum to %lld, n_halo is %lld", (long long)c->recv_off[npeers], (long long)c->n_halo);
    c->n_send = npeers > 0 ? c->send_off[npeers] : 0;
    std::vector<int> rows((size_t)c->n_send);
    for (int64_t i = 0; i < c->n_send; ++i) {
        const int64_t p = send_idx1[i] - 1;
        if (p < 0 || p >= c->n_local) return fail(MFT_EINVAL, "mft_set_halo: send index %lld is not an owned point", (long long)send_idx1[i]);
        rows[i] = c->have_perm ? c->iperm[p] : (int)p;
    }
    CHECK(c->send_rows.upload(rows));
    c->send_rows_host = rows;
    CHECK(c->send_buf.alloc(std::max<int64_t>(1, c->n_send) * 2 * c->V));
    return MFT_OK;
}

// Fused step on several GPUs: block -> tile map of a union-tile operator.  A tile is a BAND tile if its stencil union holds a
// halo column (it must wait for the peers' rows) or -- forward operator -- if one of its rows is in a peer's halo (its pass A
// epilogue writes g into the peer's memory, which must not happen before the peer is done with the previous epoch: the same
// wait orders that).  Interior tiles come first and start at once; band tiles are scheduled last.  At least one band tile
// exists per kernel, so a finished kernel has always seen every source rank's flag of the epoch (the ordering argument in
// DESIGN.md section 7 rests on that).
static int build_tile_order(mft_ctx *c, DevTileR &e, bool with_send_rows)
{
    if (!e.ready()) return MFT_OK;
    std::vector<int> uoff((size_t)e.ntiles + 1), ulist((size_t)e.uoff.n > 0 ? (size_t)e.ulist.n : 0);
    CU(cudaMemcpy(uoff.data(), e.uoff.p, sizeof(int) * uoff.size(), cudaMemcpyDeviceToHost));
    if (!ulist.empty()) CU(cudaMemcpy(ulist.data(), e.ulist.p, sizeof(int) * ulist.size(), cudaMemcpyDeviceToHost));
    std::vector<char> band((size_t)e.ntiles, 0);
    const int64_t rows_per_tile = (int64_t)kSlice * kTileWarps * e.R;
    for (int t = 0; t < e.ntiles; ++t)
        for (int q = uoff[(size_t)t]; q < uoff[(size_t)t + 1]; ++q) {
            const int j = ulist[(size_t)q];
            if (j >= c->n_local && j < c->n_tot) {   // n_tot itself is the dummy record
                band[(size_t)t] = 1;
                break;
            }
        }
    if (with_send_rows)
        for (int row : c->send_rows_host) band[(size_t)(row / rows_per_tile)] = 1;
    int nband = 0;
    for (char b : band) nband += b;
    if (nband == 0 && e.ntiles > 0) band[(size_t)e.ntiles - 1] = 1;
    std::vector<int> order;
    order.reserve((size_t)e.ntiles);
    for (int t = 0; t < e.ntiles; ++t)
        if (!band[(size_t)t]) order.push_back(t);
    e.n_free = (int)order.size();
    for (int t = 0; t < e.ntiles; ++t)
        if (band[(size_t)t]) order.push_back(t);
    CHECK(e.order.upload(order));
    return MFT_OK;
}

// peer-memory norms in three separately launchable parts: 0 = local sum + publish, 1 = wait sums, max deviation from the
// global mean + publish, 2 = wait candidates and stage them for pass A
static int p2p_norms_part(mft_ctx *c, int part)
{
    ScopedTimer t(c, MFT_K_REDUCE);
    const int V = 4;
    const int64_t n = c->n_local;
    const Vec<4> *u = reinterpret_cast<const Vec<4> *>(c->u.p);
    P2PLocal *L = reinterpret_cast<P2PLocal *>(c->p2p_local.p);
    if (part == 0) {
        k_p2p_sum<<<c->red_blocks, 256, 0, c->stream>>>(u, n, c->partial.p, c->peers_dev, L);
    } else if (part == 1) {
        k_p2p_wait_norms<<<1, 32, 0, c->stream>>>(c->peers_dev, L, 0, nullptr);
        c->launches++;
        const double ng = (double)c->n_global;
        const double divisor = c->mean_div_vn ? (double)V * ng : ng;
        if (c->max_lex) k_p2p_maxdev<true><<<c->red_blocks, 256, 0, c->stream>>>(u, n, divisor, c->partial.p, c->peers_dev, L, c->stats.p + V);
        else k_p2p_maxdev<false><<<c->red_blocks, 256, 0, c->stream>>>(u, n, divisor, c->partial.p, c->peers_dev, L, c->stats.p + V);
    } else {
        k_p2p_wait_norms<<<1, 32, 0, c->stream>>>(c->peers_dev, L, 1, c->gather_buf.p + (int64_t)c->nranks * V);
    }
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

// global ode_mean / ode_maximum across ranks (MPI.Allreduce in src/auxiliary/mpi.jl:45-46,76): every rank reduces its
// owned points, the per-rank results are all-gathered, and the CONSUMER kernel combines them in rank order
// (k_maxdev_norms forms the mean from the gathered sums, pass A forms the norms from the gathered candidates):
// two kernels + two tiny all-gathers per stage.
static int launch_norms_multi(mft_ctx *c)
{
    ScopedTimer t(c, MFT_K_REDUCE);
    const int V = 4;
    NcclApi *N = c->nccl;
    if (c->p2p) {
        CHECK(p2p_norms_part(c, 0));
        CHECK(p2p_norms_part(c, 1));
        return p2p_norms_part(c, 2);
    }
    if (!c->comm) return fail(MFT_EINVAL, "multi-rank norms need mft_comm_init");
    const int64_t n = c->n_local;
    const Vec<4> *u = reinterpret_cast<const Vec<4> *>(c->u.p);
    double *gsum = c->gather_buf.p;                              // nranks x V
    double *gmax = c->gather_buf.p + (int64_t)c->nranks * V;     // nranks x V
    double *mine = c->gather_buf.p + (int64_t)2 * c->nranks * V; // 2V scratch (sum | mean, unused)
    double *mine2 = mine + 2 * V;                                // V scratch
    k_sum_mean<4><<<c->red_blocks, 256, 0, c->stream>>>(u, n, c->partial.p, c->ticket.p, 1.0, mine);
    c->launches++;
    LAUNCH_CHECK();
    if (N->allGather(mine, gsum, V, NCCL_DOUBLE, c->comm, c->stream) != 0)
        return fail(MFT_ENCCL, "ncclAllGather failed: %s", N->lastError(c->comm));
    const double ng = (double)c->n_global;
    const double divisor = c->mean_div_vn ? (double)V * ng : ng;
    if (c->max_lex)
        k_maxdev_norms<4, true><<<c->red_blocks, 256, 0, c->stream>>>(u, n, gsum, c->nranks, divisor, c->partial.p, c->ticket.p + 1, mine2, 0, c->stats.p + V);
    else
        k_maxdev_norms<4, false><<<c->red_blocks, 256, 0, c->stream>>>(u, n, gsum, c->nranks, divisor, c->partial.p, c->ticket.p + 1, mine2, 0, c->stats.p + V);
    c->launches++;
    LAUNCH_CHECK();
    if (N->allGather(mine2, gmax, V, NCCL_DOUBLE, c->comm, c->stream) != 0)
        return fail(MFT_ENCCL, "ncclAllGather failed: %s", N->lastError(c->comm));
    return MFT_OK;
}

// ------------------------------------------------------------------------------------------------------
// peer-memory exchange setup: CUDA IPC handles of {u, g, window} travel through the host program (all-gather)
// ------------------------------------------------------------------------------------------------------
extern "C" int mft_p2p_handles(mft_ctx *c, void *out3x64)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (!out3x64) return fail(MFT_EINVAL, "mft_p2p_handles: NULL");
    if (c->V != 4) return fail(MFT_ENOTSUP, "peer-memory exchange is implemented for Euler 2-D");
    if (!c->p2p_window.p) {
        CHECK(c->p2p_window.alloc((int64_t)sizeof(P2PWindow)));
        if (!c->p2p_local.p) CHECK(c->p2p_local.alloc((int64_t)sizeof(P2PLocal)));
        CU(cudaMemset(c->p2p_window.p, 0, sizeof(P2PWindow)));
        CU(cudaMemset(c->p2p_local.p, 0, sizeof(P2PLocal)));
        if (!c->g.p) {  // no viscosity source: still give peers something valid to map
            CHECK(c->g.alloc((c->n_tot + 1) * 2 * c->V));
            CU(cudaMemset(c->g.p, 0, sizeof(double) * (c->n_tot + 1) * 2 * c->V));
        }
    }
    cudaIpcMemHandle_t h[3];
    CU(cudaIpcGetMemHandle(&h[0], c->u.p));
    CU(cudaIpcGetMemHandle(&h[1], c->g.p));
    CU(cudaIpcGetMemHandle(&h[2], c->p2p_window.p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out3x64, h, sizeof h);
    return MFT_OK;
}

extern "C" int mft_p2p_connect(mft_ctx *c, int nranks, int rank, const void *all_handles, const int64_t *peer_dst_row,
                               int64_t n_global)
{
    NEED_CTX(c);
    if (nranks < 2 || nranks > kMaxRanks || rank < 0 || rank >= nranks || !all_handles)
        return fail(MFT_EINVAL, "mft_p2p_connect: bad arguments (2..%d ranks)", kMaxRanks);
    if (!c->p2p_window.p) return fail(MFT_EINVAL, "mft_p2p_connect: call mft_p2p_handles first");
    if (c->peers.empty() && c->n_halo > 0) return fail(MFT_EINVAL, "mft_p2p_connect: call mft_set_halo first");
    const cudaIpcMemHandle_t *H = reinterpret_cast<const cudaIpcMemHandle_t *>(all_handles);
    P2PPeers &P = c->peers_dev;
    memset(&P, 0, sizeof P);
    P.nranks = nranks;
    P.rank = rank;
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) {
            P.field[0][r] = c->u.p;
            P.field[1][r] = c->g.p;
            P.win[r] = reinterpret_cast<P2PWindow *>(c->p2p_window.p);
            continue;
        }
        void *q[3];
        for (int k = 0; k < 3; ++k) {
            cudaError_t e = cudaIpcOpenMemHandle(&q[k], H[r * 3 + k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) return fail(MFT_ECUDA, "cudaIpcOpenMemHandle(rank %d): %s (peer access over NVLink/PCIe is required)", r, cudaGetErrorString(e));
            c->ipc_opened.push_back(q[k]);
        }
        P.field[0][r] = q[0];
        P.field[1][r] = q[1];
        P.win[r] = reinterpret_cast<P2PWindow *>(q[2]);
    }
    // destinations / sources and the per-entry routing table
    std::vector<int> speer((size_t)c->n_send);
    std::vector<long long> sdst((size_t)c->n_send);
    for (size_t p = 0; p < c->peers.size(); ++p) {
        const int64_t ns = c->send_off[p + 1] - c->send_off[p];
        const int64_t nr = c->recv_off[p + 1] - c->recv_off[p];
        if (ns > 0) P.dst[P.ndst++] = c->peers[p];
        if (nr > 0) P.src[P.nsrc++] = c->peers[p];
        for (int64_t i = 0; i < ns; ++i) {
            speer[c->send_off[p] + i] = c->peers[p];
            sdst[c->send_off[p] + i] = (long long)(peer_dst_row[p] + i);
        }
    }
    CHECK(c->send_peer.upload(speer));
    CHECK(c->send_dst.upload(sdst));
    c->send_peer_host = speer;
    c->send_dst_host = sdst;
    CHECK(c->peers_dev_buf.upload(std::vector<P2PPeers>(1, P)));
    CHECK(build_row_aux(c));                        // now with the halo routes
    CHECK(build_tile_order(c, c->fwd_tiler, true));
    CHECK(build_tile_order(c, c->tra_tiler, false));
    for (auto &g : c->graphs)                       // captured steps hold the old tables
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    c->nranks = nranks;
    c->rank = rank;
    c->n_global = n_global;
    if (!c->gather_buf.p) CHECK(c->gather_buf.alloc((int64_t)2 * nranks * c->V + 4 * c->V));
    c->p2p = true;
    return MFT_OK;
}
