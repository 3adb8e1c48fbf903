// mft_time_loop.inl -- history callback, SSPRK33 / SSPRK43 steps (CUDA-graph replay), Zhang-Shu stage limiter entry points.
// Included by mft_b200.cu after rhs_device().
// ------------------------------------------------------------------------------------------------------
// history + time stepping
// ------------------------------------------------------------------------------------------------------
// time_deriv_weights! history.jl:131-152: w = scale * (A' \ b'), LU with partial pivoting
static void time_deriv_weights(int m, const double *t, double *w)
{
    double maxabs = 0.0;
    for (int i = 0; i < m; ++i) maxabs = std::fmax(maxabs, std::fabs(t[i]));
    const double scale = 1.0 / maxabs;
    double ts[8], M[64], b[8];
    for (int i = 0; i < m; ++i) ts[i] = t[i] * scale;
    for (int k = 0; k < m; ++k) {
        for (int i = 0; i < m; ++i) M[k * m + i] = std::pow(ts[i], (double)k);
        b[k] = (double)k * std::pow(ts[0], (double)(k - 1));
    }
    for (int col = 0; col < m; ++col) {
        int piv = col;
        double best = std::fabs(M[col * m + col]);
        for (int r = col + 1; r < m; ++r)
            if (std::fabs(M[r * m + col]) > best) {
                best = std::fabs(M[r * m + col]);
                piv = r;
            }
        if (piv != col) {
            for (int j = 0; j < m; ++j) std::swap(M[col * m + j], M[piv * m + j]);
            std::swap(b[col], b[piv]);
        }
        for (int r = col + 1; r < m; ++r) {
            const double l = M[r * m + col] / M[col * m + col];
            M[r * m + col] = l;
            for (int j = col + 1; j < m; ++j) M[r * m + j] = M[r * m + j] - l * M[col * m + j];
            b[r] = b[r] - l * b[col];
        }
    }
    for (int r = m - 1; r >= 0; --r) {
        double s = b[r];
        for (int j = r + 1; j < m; ++j) s = s - M[r * m + j] * b[j];
        b[r] = s / M[r * m + r];
    }
    for (int i = 0; i < m; ++i) w[i] = scale * b[i];
}

static int history_push_common(mft_ctx *c, double t, int64_t success_iter, bool given, int nterms,
                               const double *weights_or_null, int approx_order)
{
    if (c->nslots == 0) return MFT_OK;  // modify_cache! fallback: no-op without a residual-viscosity source (history.jl:87-89)
    NvtxRange range("update history");
    c->success_iter = success_iter;
    // shift_soln_history! history.jl:105-111 as a ring buffer: slot 0 = most recent
    c->hist_head = (c->hist_head + c->nslots - 1) % c->nslots;
    for (int s = c->nslots - 1; s >= 1; --s) c->time_history[s] = c->time_history[s - 1];
    c->time_history[0] = t;
    const int64_t len = c->n_tot * c->V;
    CU(cudaMemcpyAsync(c->hist[c->hist_head].p, c->u.p, sizeof(double) * len, cudaMemcpyDeviceToDevice, c->stream));
    // update_approx_du! history.jl:113-129
    ApproxDuArgs a{};
    a.out = c->approx_du.p;
    a.len = len;
    a.nterms = 0;
    if (success_iter > 0) {
        int ntp = nterms;
        if (!given) {
            ntp = (int)std::min<int64_t>(success_iter + 1, (int64_t)approx_order + 1);
            if (ntp > c->nslots) return fail(MFT_EINVAL, "mft_history_push: approx_order+1 = %d exceeds polydeg+1 = %d history slots", approx_order + 1, c->nslots);
            time_deriv_weights(ntp, c->time_history.data(), c->time_weights.data());
        } else {
            if (ntp > c->nslots || ntp > 8) return fail(MFT_EINVAL, "mft_history_push_weights: %d weights exceed %d history slots", ntp, c->nslots);
            for (int i = 0; i < ntp; ++i) c->time_weights[i] = weights_or_null[i];
        }
        a.nterms = ntp;
        for (int s = 0; s < ntp; ++s) {
            a.hist[s] = c->hist[(c->hist_head + s) % c->nslots].p;
            a.w[s] = c->time_weights[s];
        }
    }
    {
        ScopedTimer tm(c, MFT_K_OTHER);
        k_approx_du<<<c->red_blocks * 4, 256, 0, c->stream>>>(a);
        c->launches++;
        LAUNCH_CHECK();
    }
    return MFT_OK;  // asynchronous (stream order); downloads / mft_synchronize wait
}

extern "C" int mft_history_push(mft_ctx *c, double t, int64_t success_iter, int approx_order)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (approx_order < 0 || approx_order > 7) return fail(MFT_EINVAL, "mft_history_push: approx_order must be in [0,7]");
    return history_push_common(c, t, success_iter, false, 0, nullptr, approx_order);
}

extern "C" int mft_history_push_weights(mft_ctx *c, double t, int64_t success_iter, int n, const double *weights)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (n < 0 || (n > 0 && !weights)) return fail(MFT_EINVAL, "mft_history_push_weights: bad weights");
    return history_push_common(c, t, success_iter, true, n, weights, 0);
}

static int launch_limiter(mft_ctx *c, int npairs, const double *thresholds, const int *variables);

static int launch_stage(mft_ctx *c, int stage, double dt)
{
    NvtxRange range("SSPRK stage update");
    ScopedTimer tm(c, MFT_K_STAGE);
    const int64_t len = c->n_local * c->V;
    k_ssprk33_stage<<<c->red_blocks * 4, 256, 0, c->stream>>>(stage, dt, c->uprev.p, c->du.p, c->u.p, len);
    c->launches++;
    LAUNCH_CHECK();
    // stage_limiter!(u, integrator, p, t) after every stage update (OrdinaryDiffEq SSPRK33(stage_limiter!))
    if (!c->stage_lim_variables.empty())
        CHECK(launch_limiter(c, (int)c->stage_lim_variables.size(), c->stage_lim_thresholds.data(), c->stage_lim_variables.data()));
    return MFT_OK;
}

// ---- fused device-resident step (MFT_OPT_FUSED_STEP; kernels: mft_fused_kernels.cuh, band logic: mft_tile_kernels.cuh) ----
// Per stage three launches: k_stage_fused (BC pass 2 of the previous rhs!, stage update, BC pass 1, norm statistic, u halo
// puts + flags), pass A (band tiles wait for the u halo; g halo puts + flags), pass B (band tiles wait for the g halo).
static bool fused_step_ok(const mft_ctx *c)
{
    if (!c->fused_step || c->V != 4 || c->eq != MFT_EQ_EULER2D || c->srcs.size() != 1) return false;
    const int kind = c->srcs[0]->kind;
    if (kind != MFT_SRC_UPWIND && kind != MFT_SRC_RESIDUAL) return false;
    if (!c->fwd_tiler.ready() || !c->tra_tiler.ready()) return false;
    if (!c->bc_merged && !c->bcs.empty()) {
        bool any = false;
        for (auto *g : c->bcs) any |= g->nb > 0 && g->kind != MFT_BC_DO_NOTHING;
        if (any) return false;   // a point in two boundary groups: the group order matters, keep the per-group launches
    }
    if (!c->stage_lim_variables.empty()) return false;
    if (c->nranks > 1 && (!c->p2p || !c->fwd_tiler.order.p || !c->tra_tiler.order.p)) return false;
    return true;
}

static int launch_stage_fused(mft_ctx *c, int stage, double dt, bool apply_bc2)
{
    NvtxRange range("SSPRK stage update + boundary flux + norms (fused)");
    ScopedTimer tm(c, MFT_K_STAGE);
    const bool residual = c->srcs[0]->kind == MFT_SRC_RESIDUAL;
    const bool multi = c->nranks > 1;
    StageArgs a{};
    a.stage = stage;
    a.apply_bc2 = apply_bc2 ? 1 : 0;
    a.dt = dt;
    a.u = c->u.p;
    a.uprev = c->uprev.p;
    a.du = c->du.p;
    a.n = c->n_local;
    a.aux = c->row_aux.p;
    a.rows = c->row_aux_tab.p;
    a.bc_kind = c->bc_kind.p;
    a.bc_normals = c->bc_normals.p;
    a.bc_values = c->bc_values.p;
    a.partial = c->partial.p;
    a.grec = c->partial.p + (size_t)c->red_blocks * kRecDoubles;
    a.gticket = c->ticket.p + 8;
    a.ticket = c->ticket.p + 4;
    const double ng = multi ? (double)c->n_global : (double)c->n_local;
    a.divisor = c->mean_div_vn ? (double)c->V * ng : ng;
    a.lex = c->max_lex;
    a.stats = c->stats.p;
    a.L = reinterpret_cast<P2PLocal *>(c->p2p_local.p);
    a.route_peer = c->route_peer.p;
    a.route_dst = c->route_dst.p;
    if (multi) a.P = c->peers_dev;
    const int grid = c->red_blocks;
    const int nmode = !residual ? NORMS_NONE : c->max_lex ? NORMS_LEX : NORMS_COMP;
#define STAGE_K(NM, MU)                                                            \
    do {                                                                           \
        CHECK(ensure_smem(c, k_stage_fused<NM, MU>, kStageSmemBytes));             \
        CU(launch_k(k_stage_fused<NM, MU>, grid, 256, kStageSmemBytes, c->stream, c->pdl && c->pdl_next, a)); \
    } while (0)
    if (multi) {
        if (nmode == NORMS_LEX) STAGE_K(NORMS_LEX, true);
        else if (nmode == NORMS_COMP) STAGE_K(NORMS_COMP, true);
        else STAGE_K(NORMS_NONE, true);
    } else {
        if (nmode == NORMS_LEX) STAGE_K(NORMS_LEX, false);
        else if (nmode == NORMS_COMP) STAGE_K(NORMS_COMP, false);
        else STAGE_K(NORMS_NONE, false);
    }
#undef STAGE_K
    c->launches++;
    LAUNCH_CHECK();
    return MFT_OK;
}

static int ssprk33_step_fused(mft_ctx *c, double t, double dt, bool first_rhs)
{
    if (first_rhs) CHECK(rhs_device(c, t));   // k = f(u_n): the separate kernels (complete rhs! incl. BC pass 2)
    Source *s = c->srcs[0];
    const int visc = s->kind == MFT_SRC_UPWIND ? VISC_UPWIND : VISC_RESIDUAL;
    struct Guard {
        mft_ctx *c;
        ~Guard() { c->fused_active = c->pdl_next = false; }
    } guard{c};
    c->fused_active = true;
    c->fused_used = true;
    bool tables = false;   // stage-time Dirichlet tables put memcpy nodes between the kernels: no programmatic edge across them
    for (auto *g : c->bcs) tables |= g->stage_set[0] || g->stage_set[1];
    for (int stage = 1; stage <= 3; ++stage) {
        CHECK(select_stage_boundary_values(c, stage == 2 ? 1 : 0));   // rhs! of stage 1 and 3 is evaluated at t + dt, of stage 2 at t + dt/2
        c->pdl_next = stage > 1 && !tables && !c->timing;   // (the first stage kernel follows whatever ran before the step)
        CHECK(launch_stage_fused(c, stage, dt, stage > 1));
        c->pdl_next = !c->timing;
        NvtxRange r(visc == VISC_RESIDUAL ? "calc fluxes + calc SourceResidualViscosityTominec (fused)"
                                          : "calc fluxes + calc SourceUpwindViscosityTominec (fused)");
        CHECK(launch_pass_a(c, true, visc, s, false));
        CHECK(launch_pass_b(c));
    }
    c->fused_active = false;
    c->pdl_next = false;
    NvtxRange r("boundary flux");
    CHECK(launch_boundary(c, true));   // BC pass 2 of the last rhs!: u and du are complete when the step returns
    return MFT_OK;
}

static int ssprk33_step_launches(mft_ctx *c, double t, double dt, bool first_rhs)
{
    if (fused_step_ok(c)) return ssprk33_step_fused(c, t, dt, first_rhs);
    if (first_rhs) CHECK(rhs_device(c, t));  // k = f(u_n): first step only (FSAL afterwards); current Dirichlet tables = time t
    CHECK(launch_stage(c, 1, dt));
    CHECK(select_stage_boundary_values(c, 0));
    CHECK(rhs_device(c, t + dt));
    CHECK(launch_stage(c, 2, dt));
    CHECK(select_stage_boundary_values(c, 1));
    CHECK(rhs_device(c, t + dt / 2));
    CHECK(launch_stage(c, 3, dt));
    CHECK(select_stage_boundary_values(c, 0));
    CHECK(rhs_device(c, t + dt));
    return MFT_OK;
}

extern "C" int mft_ssprk_step(mft_ctx *c, int scheme, double t, double dt)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (scheme != MFT_SSPRK33) return fail(MFT_ENOTSUP, "mft_ssprk_step: only MFT_SSPRK33 is implemented");
    const bool first = !c->have_fsal;
    c->have_fsal = true;
    // The step is a fixed sequence of ~35 launches: replay it as one CUDA graph (the kernel arguments that vary
    // between steps -- dt and the success_iter==0 flag -- are part of the cache key; t only selects Dirichlet tables,
    // which the caller refreshes).  Eager path: per-kernel timing on, multi-rank (NCCL calls), or first use of a key.
    const bool graph_ok = c->use_graphs && !c->timing && (c->nranks == 1 || c->p2p || c->use_graphs >= 2);
    if (!graph_ok) return ssprk33_step_launches(c, t, dt, first);
    const int si_zero = c->success_iter == 0;
    mft_ctx::StepGraph *g = nullptr;
    for (auto &e : c->graphs)
        if (e.dt == dt && e.si_zero == si_zero && e.with_first_rhs == (int)first) g = &e;
    if (!g) {
        c->graphs.push_back(mft_ctx::StepGraph{dt, si_zero, (int)first, 0, 0, nullptr});
        g = &c->graphs.back();
    }
    g->uses++;
    if (g->uses == 1) return ssprk33_step_launches(c, t, dt, first);  // warm (also sets per-kernel smem attributes)
    if (!g->exec) {
        const int64_t l0 = c->launches;
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = ssprk33_step_launches(c, t, dt, first);
        const cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
        if (rc != MFT_OK) {
            if (graph) cudaGraphDestroy(graph);
            return rc;
        }
        if (ce != cudaSuccess) return fail(MFT_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(ce));
        g->nlaunch = c->launches - l0;
        c->launches = l0;
        CU(cudaGraphInstantiate(&g->exec, graph, 0));
        CU(cudaGraphDestroy(graph));
    }
    CU(cudaGraphLaunch(g->exec, c->stream));
    c->launches += g->nlaunch;
    return MFT_OK;  // asynchronous: mft_synchronize / downloads wait
}

// ---- Zhang-Shu positivity limiter (row f4; positivity_zhang_shu_point2d.jl:22-82, positivity_zhang_shu.jl:50-72) --------
extern "C" int mft_set_neighbors(mft_ctx *c, const int64_t *nbr1)
{
    NEED_CTX(c);
    if (!nbr1) return fail(MFT_EINVAL, "mft_set_neighbors: NULL array");
    if (c->k <= 0) return fail(MFT_EINVAL, "mft_set_neighbors: ctx was created with k=%d", c->k);
    const int64_t n = c->n_local, k = c->k;
    c->host_nbr.resize((size_t)(n * k));
    for (int64_t p = 0; p < n * k; ++p) {
        const int64_t j = nbr1[p] - 1;
        if (j < 0 || j >= c->n_tot) return fail(MFT_EINVAL, "mft_set_neighbors: neighbour %lld out of range", (long long)nbr1[p]);
        c->host_nbr[(size_t)p] = (int32_t)j;
    }
    // captured steps may hold the old table's address: start over
    for (auto &g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    c->zs_nbr.release();  // rebuilt (device numbering) at the next limiter call
    return MFT_OK;
}

static int zs_prepare(mft_ctx *c)
{
    if (c->V != 4 || c->eq != MFT_EQ_EULER2D) return fail(MFT_ENOTSUP, "Zhang-Shu limiter: Euler 2-D only");
    if (c->host_nbr.empty()) return fail(MFT_EINVAL, "Zhang-Shu limiter: mft_set_neighbors was not called");
    if (c->zs_nbr.p) return MFT_OK;
    const int64_t n = c->n_local, k = c->k;
    std::vector<int> tab((size_t)(n * k));
    for (int64_t d = 0; d < n; ++d) {
        const int64_t r = c->have_perm ? c->perm[d] : d;  // device row d holds caller point r
        for (int64_t q = 0; q < k; ++q) {
            const int32_t j = c->host_nbr[(size_t)(r * k + q)];
            tab[(size_t)(q * n + d)] = c->have_perm ? c->iperm[j] : j;
        }
    }
    CHECK(c->zs_nbr.upload(tab));
    CHECK(c->zs_tmp.alloc(n * c->V));
    CHECK(c->zs_flag.alloc(n));
    return MFT_OK;
}

// one limiter call = one pass per (threshold, variable) pair, in order, each pass on the state the previous one left
static int launch_limiter(mft_ctx *c, int npairs, const double *thresholds, const int *variables)
{
    CHECK(zs_prepare(c));
    ScopedTimer tm(c, MFT_K_OTHER);
    const int64_t n = c->n_local;
    for (int i = 0; i < npairs; ++i) {
        if (variables[i] != ZS_VAR_DENSITY && variables[i] != ZS_VAR_PRESSURE) return fail(MFT_EINVAL, "Zhang-Shu limiter: unknown variable %d", variables[i]);
        // multi-rank: the pass reads u at every stencil point of the owned rows -> refresh the halo copies first (the stage
        // update / the previous pass changed the owners' values).  Same exchange as at the start of rhs!; the credit protocol
        // of the peer-memory path allows any stream-ordered sequence of exchanges as long as all ranks issue the same one.
        CHECK(halo_exchange<4>(c, c->u.p));
        ZsArgs a{c->zs_nbr.p, c->k, n, c->u.p, c->zs_tmp.p, c->zs_flag.p, thresholds[i], c->eqp[0], variables[i]};
        k_zs_detect<<<grid_for(n, 128), 128, 0, c->stream>>>(a);
        k_zs_apply<<<grid_for(n, 256), 256, 0, c->stream>>>(n, c->zs_flag.p, c->zs_tmp.p, c->u.p);
        c->launches += 2;
        LAUNCH_CHECK();
    }
    return MFT_OK;
}

extern "C" int mft_limiter_zhang_shu(mft_ctx *c, int npairs, const double *thresholds, const int *variables,
                                     double *const *u_soa, int mem)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (npairs < 0 || (npairs > 0 && (!thresholds || !variables))) return fail(MFT_EINVAL, "mft_limiter_zhang_shu: bad arguments");
    if (mem == MFT_MEM_HOST) {
        CHECK(upload_soa(c, u_soa, c->u.p));
    } else if (mem != MFT_MEM_DEVICE) {
        return fail(MFT_EINVAL, "mft_limiter_zhang_shu: mem must be MFT_MEM_HOST or MFT_MEM_DEVICE");
    }
    c->have_fsal = false;  // u changed: f(u) has to be recomputed
    CHECK(launch_limiter(c, npairs, thresholds, variables));
    if (mem == MFT_MEM_HOST) {
        CHECK(download_soa(c, c->u.p, u_soa));
        CU(cudaStreamSynchronize(c->stream));
    }
    return MFT_OK;
}

extern "C" int mft_set_stage_limiter(mft_ctx *c, int npairs, const double *thresholds, const int *variables)
{
    NEED_CTX(c);
    if (npairs < 0 || (npairs > 0 && (!thresholds || !variables))) return fail(MFT_EINVAL, "mft_set_stage_limiter: bad arguments");
    c->stage_lim_thresholds.assign(thresholds, thresholds + npairs);
    c->stage_lim_variables.assign(variables, variables + npairs);
    // captured steps bake the launch sequence in: start over
    for (auto &g : c->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
    return MFT_OK;
}

// ---- SSPRK43 with embedded error estimate (the integrator the reference names: rbfsolver_test.jl:104-107) ------------
// The library does the four stages and returns the LOCAL sum of squared scaled errors and the local entry count; the
// caller combines ranks (sum both), forms EEst = sqrt(sumsq/count) (ode_norm, src/auxiliary/mpi.jl:15-19) and runs its
// own step-size controller (OrdinaryDiffEq's stays in charge in the Julia deployment), then commits or rolls back
// with mft_step_commit.
extern "C" int mft_ssprk43_step(mft_ctx *c, double t, double dt, double abstol, double reltol, double *sumsq_out,
                                int64_t *count_out)
{
    NEED_CTX(c);
    CHECK(mft_finalize(c));
    if (c->step_pending) return fail(MFT_EINVAL, "mft_ssprk43_step: previous step was neither committed nor rejected (mft_step_commit)");
    if (!c->stage_lim_variables.empty()) return fail(MFT_ENOTSUP, "mft_ssprk43_step: the stage limiter is wired into mft_ssprk_step (SSPRK33) only");
    const int64_t len = c->n_local * c->V, len_tot = c->n_tot * c->V;
    if (!c->utilde.p) {
        CHECK(c->utilde.alloc(len_tot));
        CHECK(c->kfsal.alloc(len_tot));
        CHECK(c->u_save.alloc(len_tot + c->V));
        CU(cudaMemsetAsync(c->utilde.p, 0, sizeof(double) * len_tot, c->stream));
    }
    if (!c->have_fsal) CHECK(rhs_device(c, t));  // k = f(u_n, t)
    c->have_fsal = true;
    // keep f(u_n) and u_n (rhs! also rewrites boundary / halo entries of u) for a possible rejection
    CU(cudaMemcpyAsync(c->kfsal.p, c->du.p, sizeof(double) * len_tot, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaMemcpyAsync(c->u_save.p, c->u.p, sizeof(double) * len_tot, cudaMemcpyDeviceToDevice, c->stream));
    const int grid = c->red_blocks * 4;
    auto stage = [&](int st) -> int {
        ScopedTimer tm(c, MFT_K_STAGE);
        k_ssprk43_stage<<<grid, 256, 0, c->stream>>>(st, dt, c->uprev.p, c->du.p, c->u.p, c->utilde.p, len);
        c->launches++;
        LAUNCH_CHECK();
        return MFT_OK;
    };
    CHECK(stage(1));
    CHECK(select_stage_boundary_values(c, 1));
    CHECK(rhs_device(c, t + dt / 2));
    CHECK(stage(2));
    CHECK(select_stage_boundary_values(c, 0));
    CHECK(rhs_device(c, t + dt));
    CHECK(stage(3));
    CHECK(select_stage_boundary_values(c, 1));
    CHECK(rhs_device(c, t + dt / 2));
    CHECK(stage(4));
    {
        ScopedTimer tm(c, MFT_K_REDUCE);
        k_error_sumsq<<<c->red_blocks, 256, 0, c->stream>>>(c->utilde.p, c->uprev.p, c->u.p, len, abstol, reltol, c->partial.p,
                                                         c->ticket.p + 2, c->stats.p + 3 * c->V);
        c->launches++;
        LAUNCH_CHECK();
    }
    CHECK(select_stage_boundary_values(c, 0));
    CHECK(rhs_device(c, t + dt));  // FSAL: k = f(u_{n+1}, t+dt)
    double ss = 0.0;
    CU(cudaMemcpyAsync(&ss, c->stats.p + 3 * c->V, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (sumsq_out) *sumsq_out = ss;
    if (count_out) *count_out = len;
    c->step_pending = true;
    return MFT_OK;
}

// accept != 0: keep u_{n+1} and its f; accept == 0: restore u_n and f(u_n) (a rejected step leaves no trace)
extern "C" int mft_step_commit(mft_ctx *c, int accept)
{
    NEED_CTX(c);
    if (!c->step_pending) return fail(MFT_EINVAL, "mft_step_commit: no step pending");
    if (!accept) {
        const int64_t len_tot = c->n_tot * c->V;
        CU(cudaMemcpyAsync(c->u.p, c->u_save.p, sizeof(double) * len_tot, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(c->du.p, c->kfsal.p, sizeof(double) * len_tot, cudaMemcpyDeviceToDevice, c->stream));
    }
    c->step_pending = false;
    return MFT_OK;
}
