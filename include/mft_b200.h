/*
 * mft_b200.h -- C ABI of libmft_b200.so: the B200 (sm_100a) implementation of MeshfreeTrixi.jl's
 * semidiscrete right-hand side `rhs!` and the pieces around it.
 *
 * This is the drop-in boundary.  The reference's reserved plug-in point is the empty engine type
 * `RBFFDEngineCUDA` (src/solvers/rbfsolver.jl:60-61, selected with
 * `PointCloudSolver(basis; engine = RBFFDEngineCUDA())`, src/solvers/pointcloudsolver/types.jl:71-74).
 * Julia methods specialised on that engine `ccall` the functions below (see INTEGRATION.md for the glue).
 * All file:line citations are relative to the reference repository root.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MFT_E* code on failure; mft_last_error() returns a
 *     thread-local message.  No exceptions cross the boundary.
 *   - the caller owns every host array; the library copies during set_* calls and never keeps host pointers
 *     (exception: MFT_MEM_HOST calls read/write the caller's arrays synchronously during the call).
 *   - indices arrive exactly as Julia stores them: 1-based Int64.
 *   - state arrays are SoA: `V` separate `double*` of length n_local+n_halo, i.e. StructArrays.components(u)
 *     (allocate_nested_array, src/solvers/pointcloudsolver/rbfsolver.jl:111-116).
 *   - a ctx is bound to one device, is not thread-safe and not re-entrant; calls are synchronous.
 *   - there is no CPU fallback: without a usable CUDA device mft_ctx_create fails with MFT_ENODEVICE.
 */
#ifndef MFT_B200_H
#define MFT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mft_ctx mft_ctx;

/* error codes */
#define MFT_OK 0
#define MFT_EINVAL (-1)   /* bad argument / call order */
#define MFT_ECUDA (-2)    /* CUDA runtime error */
#define MFT_ENODEVICE (-3)
#define MFT_ENOTSUP (-4)  /* combination not implemented (e.g. residual viscosity for advection) */
#define MFT_ENCCL (-5)
#define MFT_ENORMS (-6)   /* fused step: some rows exceeded the one-pass ode_maximum statistic (a ROUNDING tie decided the lexicographic order,
                           * DESIGN.md 3d); returned by mft_synchronize / mft_download_state so that it is never silent.  Re-run with
                           * MFT_OPT_FUSED_STEP = 0 (two-pass norms).  The counter is MFT_FIELD_NORM_MISSES. */

/* equations: Trixi CompressibleEulerEquations2D(gamma) / LinearScalarAdvectionEquation2D(a1,a2);
 * flux call site src/solvers/pointcloudsolver/rbfsolver.jl:259 */
#define MFT_EQ_EULER2D 0      /* params: gamma            ; nvars = 4 */
#define MFT_EQ_ADVECTION2D 1  /* params: a1, a2           ; nvars = 1 */

/* operator slots: cache.rbf_differentiation_matrices[1:2] (rbfsolver.jl:139,162) */
#define MFT_OP_DX 0
#define MFT_OP_DY 1

/* boundary kinds: functors of src/equations/PointCloudBCs.jl */
#define MFT_BC_DIRICHLET 0   /* :49-63   du_b = 0, u_b = g(x,t) (value table)            */
#define MFT_BC_SLIP_WALL 1   /* :87-106  du_b = (du1,0,0,du4), u_b = slip-projected        */
#define MFT_BC_DO_NOTHING 2  /* :108-115                                                     */

/* source kinds: callable structs of src/sources/hyperviscosity.jl */
#define MFT_SRC_HV_FLYER 0    /* :52-64    du += -gamma * H u          params: gamma                    + matrix */
#define MFT_SRC_HV_TOMINEC 1  /* :121-134  du += -gamma * (L'L) u      params: gamma                    + matrix */
#define MFT_SRC_UPWIND 2      /* :351-380  params: c_uw, dx_avg                                                  */
#define MFT_SRC_RESIDUAL 3    /* :382-409  params: c_rv, c_uw, dx_avg, polydeg                                   */
#define MFT_SRC_IGR 4         /* src/sources/IGR.jl:211-239  params: alpha [, maxiter = 20] ; sigma solved by CG (IterativeSolvers.cg!) */

#define MFT_MEM_HOST 0
#define MFT_MEM_DEVICE 1

/* options (mft_set_option) */
#define MFT_OPT_EXACT_ORDER 0      /* 1 (default): sums in the reference's order with separate mul/add ->
                                      bit-identical to the CSC SpMV of SparseArrays; 0: single sweep with FMAs */
#define MFT_OPT_MEAN_DIVISOR_VN 1  /* 1 (default): ode_mean divides by V*N (recursive_length, src/auxiliary/mpi.jl:42) */
#define MFT_OPT_MAX_LEXICOGRAPHIC 2/* 1 (default): maximum(::StructArray{SVector}) = lexicographic max (mpi.jl:71-81) */
#define MFT_OPT_DIAGNOSTICS 3      /* 1: keep eps_uw/eps_rv/eps/eps_c/residual for mft_get_field (default 0)     */
#define MFT_OPT_CUDA_GRAPH 4       /* 1 (default): mft_ssprk_step replays a captured CUDA graph on one GPU; 2: also multi-rank
                                      (NCCL calls captured into the graph); 0: always eager launches                      */
#define MFT_OPT_STAGE_WEIGHTS 5    /* bit 0 (forward operator, pass A) / bit 1 (transposed operator, pass B): 1 = a warp bulk-copies
                                      its whole operator slice (indices + weights) into shared memory, 0 = indices only, weights
                                      by coalesced loads + L2 bulk prefetch.  bit 2: exact-order pass A keeps ONE weight buffer and refills it with wy
                                      after the x sweep (smaller footprint -> larger L1).  Union-tile kernels stream their weights (measured faster) unless bit 3 is set:
                                      then bit 0 / bit 1 stage the compact weight blocks of one direction (two sweeps).
                                      Default 5 (measured best). */
#define MFT_OPT_REFINE_ORDER 7     /* 1 (default 0): within blocks of device rows (one union tile; 256 rows for the sliced-ELL kernels),
                                      order rows by D' row length: near-uniform slices of the transposed operator (fewer padding
                                      steps in pass B); tiles keep their rows and unions; the caller-visible numbering is unaffected */
#define MFT_OPT_SINGLE_SWEEP_EXACT 8/* 1: for the default 20-wide stencil use the single-sweep exact kernel (y-products parked in registers:
                                      one gather + one flux per neighbour, but 255 registers -> 8 warps/SM; measured 20 % slower);
                                      0 (default): the two-sweep exact kernel.  Same results bit for bit.                       */
#define MFT_OPT_PAIR_ROWS 9        /* bit 0: transposed operator (pass B), bit 1: forward operator (pass A, measured slower: FP64/latency-bound); default 1.
                                      An operator is stored per PAIR of consecutive rows (union of the two
                                      D' rows, zero weight where a row lacks an entry): each shared neighbour record is gathered once
                                      for both rows.  Same sums bit for bit.  0: one row per thread.                        */
#define MFT_OPT_TILE 10            /* bit 0: pass A, bit 1: pass B run over "union tiles" (Euler 2-D): a block of 128 rows loads the
                                      union of its stencils into shared memory once and rows read it with 16-bit local offsets
                                      (mft_tile_kernels.cuh); bit 2: the slots of a tile are bank-coloured (fewer LDS conflicts);
                                      bit 3: every record is kept twice under different bank assignments and each read picks the
                                      copy that avoids a conflict; bit 4 (opt-in): the bank groups of the second copy are tuned by
                                      a local search so that the points of an LDS.128 phase can be matched to distinct groups
                                      (simulated conflict degree 1.20 -> 1.02, layout only).  Same sums bit for bit.  Default 15. */
#define MFT_OPT_TILE_ROWS 11       /* rows per thread of the union-tile kernels, decimal digits: units = pass A, tens = pass B, each 1, 2
                                      or 4 (e.g. 42 = pass B 4 rows, pass A 2 rows).  A thread walks the union of its rows' stencils
                                      (16-bit word = slot | row mask << 12, weights compact per row).  Same sums bit for bit.   */
#define MFT_OPT_FUSED_STEP 12      /* 1 (default): mft_ssprk_step runs the fused device-resident step for Euler 2-D + one upwind / residual
                                    * viscosity source over union tiles: ONE kernel per stage does BC pass 2, the SSPRK33 stage update, BC
                                    * pass 1, ode_mean / ode_maximum (one-pass statistic) and -- on several GPUs -- the u halo puts; pass A
                                    * puts its g halo rows itself and tiles that touch the halo wait for it while the interior runs.
                                    * 0: separate stage / boundary / norm / put / wait kernels (the round-1 sequence). */
#define MFT_OPT_PDL 13             /* 1 (default): the kernels of a fused stage are launched as programmatic dependents of each other
                                    * (cudaLaunchAttributeProgrammaticStreamSerialization): a kernel starts its operator-only prologue
                                    * while its predecessor drains and waits (griddepcontrol.wait) before reading the predecessor's output */
#define MFT_OPT_LAYOUT_DEVICE 14   /* 1 (default): the union-tile layouts (one row per thread, the default) are built on the GPU at
                                    * mft_finalize -- one tile per warp, the per-tile code of csrc/mft_tile_build.cuh, byte-identical to
                                    * the host builder; 0: host threads (also used for the 2 / 4 rows-per-thread variants) */
#define MFT_OPT_PREFETCH_DISTANCE 6/* slices ahead for the L2 prefetch of operator data (weight blocks; union tiles: step words, weights, union
                                      list of the tile that many slices ahead).  Default 8 per SM; 0: off.               */

/* fields (mft_get_field): caches of create_tominec_rv_cache, hyperviscosity.jl:202-244 */
#define MFT_FIELD_EPS 0        /* N doubles   */
#define MFT_FIELD_EPS_UW 1     /* N doubles   */
#define MFT_FIELD_EPS_RV 2     /* N doubles   */
#define MFT_FIELD_EPS_C 3      /* N doubles (0,1,2) */
#define MFT_FIELD_RESIDUAL 4   /* V*N doubles, SoA */
#define MFT_FIELD_APPROX_DU 5  /* V*N doubles, SoA */
#define MFT_FIELD_NORMS 6      /* V doubles: n_inf_norms of update_residual_visc! */
#define MFT_FIELD_SIGMA 7      /* N doubles: cache.sigma of SourceIGR (IGR.jl:169-191) */
#define MFT_FIELD_IGR_STATUS 8 /* 3 doubles: CG iterations of the last solve, final |r|, initial |r| */
#define MFT_FIELD_NORM_MISSES 9/* 1 double: rows (summed over all rhs! so far) whose |u - mean| exceeded the norms of the fused step's one-pass
                                * statistic (csrc/mft_fused_kernels.cuh); 0 unless a rounding tie decided ode_maximum's lexicographic order */

#define MFT_SSPRK33 0

const char *mft_last_error(void);
int mft_version(void);  /* 100 * major + 10 * minor: 120 = 1.2 (1.1: setup pipeline, limiter, IGR, non-finite check; 1.2: stage-time
                           Dirichlet tables, tuned tile layout, CSR self test) */
int mft_device_count(void);

/* ---- lifetime -------------------------------------------------------------------------------------
 * replaces Trixi.create_cache for the CUDA engine (reference: rbfsolver.jl:130-165; parallel: parallel_rbfsolver.jl:17-54).
 * n_local owned points followed by n_halo halo points (layout of ParallelPointCloudDomain,
 * src/domains/PointCloudDomain/ParallelPointCloud.jl:144); n_halo = 0 on one GPU. */
int mft_ctx_create(mft_ctx **out, int device, int64_t n_local, int64_t n_halo, int nvars, int ndims, int k);
int mft_ctx_destroy(mft_ctx *ctx);

int mft_set_equation(mft_ctx *ctx, int kind, const double *params, int nparams);
int mft_set_option(mft_ctx *ctx, int option, double value);

/* device ordering: device row d holds caller point perm1[d] (1-based).  Owned points must stay in front of
 * halo points.  Optional; default identity.  mft_sfc_order() below produces a Hilbert ordering. */
int mft_set_permutation(mft_ctx *ctx, const int64_t *perm1);
/* reference summation rank of every caller point (default: its own index).  A multi-GPU caller passes global
 * point ids so that per-row sums run in ascending GLOBAL column order like the serial CSC SpMV. */
int mft_set_order_keys(mft_ctx *ctx, const int64_t *keys);

/* operators as Julia SparseMatrixCSC{Float64,Int64} fields (colptr n+1, rowval nnz, nzval nnz; 1-based), n = n_local+n_halo.
 * Dx and Dy must share one sparsity pattern (they do: compute_flux_operator, compute_operators.jl:443-452). */
int mft_set_operator_csc(mft_ctx *ctx, int slot, const int64_t *colptr, const int64_t *rowval, const double *nzval);
/* same operators as neighbour/weight tables: nbr1 (n_local x k, row-major, 1-based), wx, wy (n_local x k) */
int mft_set_operator_ell(mft_ctx *ctx, const int64_t *nbr1, const double *wx, const double *wy);

/* boundary groups, call order = NamedTuple order of `boundary_conditions` (calc_boundary_flux!, rbfsolver.jl:277-286).
 * idx1: 1-based point indices (BoundaryData.idx), normals: nb x 2 row-major, values: Dirichlet table SoA values[v*nb+j] or NULL */
int mft_add_boundary(mft_ctx *ctx, int kind, int64_t nb, const int64_t *idx1, const double *normals, const double *values);
int mft_update_boundary_values(mft_ctx *ctx, int group, const double *values);
/* Time-dependent Dirichlet data inside a device-resident step (mft_ssprk_step / mft_ssprk43_step).  The reference's BC
 * closures receive the stage time (calc_single_boundary_flux!, rbfsolver.jl:311-316) and one step
 * evaluates rhs! at t + dt and t + dt/2: hand the tables for both times to the library before the step (slot 0 = values at
 * t + dt, slot 1 = values at t + dt/2; same layout as mft_update_boundary_values); it makes each the current table in front
 * of the rhs! evaluated at that time.  The table current at the start of a run (time t0) is set with
 * mft_update_boundary_values.  Groups without stage tables keep their table for the whole step. */
int mft_set_stage_boundary_values(mft_ctx *ctx, int group, int slot, const double *values);

/* sources, call order = order of SourceTerms(...) (calc_sources!, rbfsolver.jl:388-395).  HV kinds take their
 * matrix (CSC as above); others pass NULLs. */
int mft_add_source(mft_ctx *ctx, int kind, const double *params, int nparams, const int64_t *colptr,
                   const int64_t *rowval, const double *nzval);

/* builds the device layouts; called implicitly by the first compute call */
int mft_finalize(mft_ctx *ctx);

/* ---- the hot path -----------------------------------------------------------------------------------
 * Trixi.rhs!(du,u,t,domain,...) rbfsolver.jl:397-428.  u is IN/OUT (strong BCs overwrite boundary points).
 * MFT_MEM_HOST: u_soa/du_soa are host arrays (H2D of u, D2H of u and du inside the call).
 * MFT_MEM_DEVICE: pointers are ignored; operates on the resident state (see mft_upload_state). */
int mft_rhs(mft_ctx *ctx, double t, double *const *u_soa, double *const *du_soa, int mem);
/* calc_fluxes! rbfsolver.jl:247-265: du += -Dx F(u) - Dy G(u)  (host arrays; du is accumulated into) */
int mft_calc_fluxes(mft_ctx *ctx, double *const *u_soa, double *const *du_soa);
/* source functor call source(du,u,t,...) for source number `index` (0-based, order of mft_add_source) */
int mft_apply_source(mft_ctx *ctx, int index, double t, double *const *u_soa, double *const *du_soa);
/* calc_boundary_flux! rbfsolver.jl:277-318 on host arrays (one pass) */
int mft_boundary_pass(mft_ctx *ctx, double t, double *const *u_soa, double *const *du_soa);

/* ---- resident state + time loop ---------------------------------------------------------------------- */
int mft_upload_state(mft_ctx *ctx, const double *const *u_soa);
int mft_download_state(mft_ctx *ctx, double *const *u_soa);
int mft_download_du(mft_ctx *ctx, double *const *du_soa);
/* HistoryCallback affect: modify_cache!(::SourceResidualViscosityTominec) history.jl:91-129 on the RESIDENT u.
 * The snapshot that enters the solution history is the device-resident state: after mft_ssprk_step / mft_upload_state that is
 * the state the callback means.  A host-driven integrator (mft_rhs with MFT_MEM_HOST) leaves the last STAGE value (post-BC)
 * resident, so such a caller uploads the callback's u first (mft_upload_state; julia/MeshfreeTrixiB200.jl does).
 * Weights from time_deriv_weights! (:131-152) are computed by the library ... */
int mft_history_push(mft_ctx *ctx, double t, int64_t success_iter, int approx_order);
/* ... or supplied by the caller (n weights, most recent first), so the Julia side can keep its own solve */
int mft_history_push_weights(mft_ctx *ctx, double t, int64_t success_iter, int n, const double *weights);
/* one SSPRK step on the resident state, FSAL structure (k=f(u_n) carried between steps), 3 rhs! per step */
int mft_ssprk_step(mft_ctx *ctx, int scheme, double t, double dt);
/* one SSPRK43 step (4 stages + FSAL rhs!) with OrdinaryDiffEq's embedded error estimate; the integrator the reference names
 * (rbfsolver_test.jl:104-107).  Returns the local sum_i (err_i / (abstol + max(|u_n,i|,|u_n+1,i|) reltol))^2 and the number
 * of local entries; the caller sums both over ranks, EEst = sqrt(sumsq/count) (ode_norm, src/auxiliary/mpi.jl:15-19), runs its
 * step-size controller and then calls mft_step_commit(ctx, accept).  A rejected step restores u_n and f(u_n). */
int mft_ssprk43_step(mft_ctx *ctx, double t, double dt, double abstol, double reltol, double *sumsq_out, int64_t *count_out);
int mft_step_commit(mft_ctx *ctx, int accept);
int mft_get_field(mft_ctx *ctx, int field, double *out);
int mft_synchronize(mft_ctx *ctx);
/* failure detection (the reference has none: Trixi's ode_unstable_check is imported at src/MeshfreeTrixi.jl:33 and unused):
 * number of NaN / +-Inf entries in the owned rows of the resident state */
int mft_count_nonfinite(mft_ctx *ctx, int64_t *count_out);
/* number of kernels launched by this ctx since creation (bench.py `gpu_launches`) */
int64_t mft_launch_count(mft_ctx *ctx);
/* CUDA-event time (ms) of kernel class `which` accumulated since the last reset; which<0 resets all. */
int mft_kernel_time_ms(mft_ctx *ctx, int which, double *ms, int64_t *launches);
#define MFT_K_PASS_A 0
#define MFT_K_PASS_B 1
#define MFT_K_REDUCE 2
#define MFT_K_STAGE 3
#define MFT_K_BC 4
#define MFT_K_OTHER 5
int mft_set_kernel_timing(mft_ctx *ctx, int enable);

/* CUDA-event stopwatch on the ctx's own stream (the stream every kernel of this ctx is launched on) */
int mft_timer_start(mft_ctx *ctx);
int mft_timer_stop(mft_ctx *ctx, double *elapsed_ms); /* records, synchronises, returns the elapsed device time */

/* pinned host memory helpers for MFT_MEM_HOST callers */
int mft_host_alloc(void **out, int64_t bytes);
int mft_host_free(void *p);
int mft_host_register(void *p, int64_t bytes);
int mft_host_unregister(void *p);

/* ---- setup helpers (host code, no GPU needed) ---------------------------------------------------------
 * Hilbert space-filling-curve order of a 2-D cloud: perm1_out[d] = 1-based index of the d-th point along the curve. */
int mft_sfc_order(int64_t n, const double *x, const double *y, int64_t *perm1_out);

/* ---- setup pipeline on the device ------------------------------------------
 * Stateless: host arrays in, host arrays out; the device is used for the search / the batched solves.
 *
 * mft_setup_knn replaces the KDTree + knn(kdtree, points, nv, true) calls of the PointData constructor
 * (src/domains/PointCloudDomain/geometry_primatives.jl:322-339): nbr1_out = n x k row-major, 1-based, ascending by
 * distance, the point itself first, exact distance ties by ascending index; dist_out (nullable) = the n x k distances
 * (dx_min = minimum, dx_avg = mean of column 2, as the reference forms them from its 2-NN query). */
int mft_setup_knn(int device, int64_t n, const double *x, const double *y, int k, int64_t *nbr1_out, double *dist_out);
/* the same search for the nq listed points only (query_idx1: 1-based indices into x, y; outputs nq x k): what one rank of a
 * partitioned cloud needs for its owned + halo rows (replaces the per-rank knn calls of partition_domain.jl:277-318) */
int mft_setup_knn_queries(int device, int64_t n, const double *x, const double *y, int k, int64_t nq, const int64_t *query_idx1,
                          int64_t *nbr1_out, double *dist_out);
/* mft_setup_rbf_weights replaces the per-point loop of compute_flux_operator
 * (src/solvers/pointcloudsolver/compute_operators.jl:409-453; deriv_order > 1: :549-594) for the polyharmonic spline
 * basis r^phs_power + monomials up to poly_degree: wx_out / wy_out = n x k row-major weights of d^m/dx^m and d^m/dy^m
 * (m = deriv_order, 1..4), aligned with nbr1 (= domain.pd.neighbors, 1-based).  Feed them to mft_set_operator_ell, or
 * assemble sparse(I, J, V) on the caller's side. */
int mft_setup_rbf_weights(int device, int64_t n, const double *x, const double *y, int k, const int64_t *nbr1, int phs_power,
                          int poly_degree, int deriv_order, double *wx_out, double *wy_out);
/* the same for n_rows stencils given as rows of nbr1 (n_rows x k, indices into the n points): a rank's owned + halo rows */
int mft_setup_rbf_weights_rows(int device, int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1,
                               int phs_power, int poly_degree, int deriv_order, double *wx_out, double *wy_out);
/* the same for the HybridGaussianPHS basis phi = alpha exp(-(epsilon r)^2) + beta r^phs_power
 * (RBF(HybridGaussianPHS(; Nrbf, alpha, beta, epsilon)), src/domains/PointCloudDomain/geometry_primatives.jl:117-132, 238-262) */
int mft_setup_rbf_weights_hybrid(int device, int64_t n, const double *x, const double *y, int64_t n_rows, int k, const int64_t *nbr1,
                                 int phs_power, double alpha, double beta, double epsilon, int poly_degree, int deriv_order,
                                 double *wx_out, double *wy_out);

/* ---- Zhang-Shu positivity limiter (stage callback) ------------------------------------------------------
 * replaces Trixi.limiter_zhang_shu!(u, threshold, variable, domain::PointCloudDomain{2}, ...)
 * (src/callbacks_stage/positivity_zhang_shu_point2d.jl:22-82) and the (thresholds, variables) recursion of
 * PositivityPreservingLimiterZhangShu (src/callbacks_stage/positivity_zhang_shu.jl:29-72).  Euler 2-D.  Multi-rank: every rank
 * passes the stencils of its owned rows in local numbering and calls the limiter collectively (one u halo refresh per pass). */
#define MFT_VAR_DENSITY 0   /* Trixi.density  */
#define MFT_VAR_PRESSURE 1  /* Trixi.pressure */
/* domain.pd.neighbors flattened: n_local x k row-major, 1-based, in the kNN list order (distance-sorted, self first) */
int mft_set_neighbors(mft_ctx *ctx, const int64_t *nbr1);
/* limiter!(u_ode, integrator, semi, t): one pass per (threshold, variable) pair, in order.  MFT_MEM_HOST: u_soa in/out;
 * MFT_MEM_DEVICE: acts on the resident state (u_soa ignored). */
int mft_limiter_zhang_shu(mft_ctx *ctx, int npairs, const double *thresholds, const int *variables, double *const *u_soa,
                          int mem);
/* SSPRK33(stage_limiter!): mft_ssprk_step applies the limiter after every stage update.  npairs = 0 removes it. */
int mft_set_stage_limiter(mft_ctx *ctx, int npairs, const double *thresholds, const int *variables);

/* Host-only self test of the union-tile operator format (MFT_OPT_TILE): lays out a random ragged banded operator for
 * n points (k entries per row, R = 1, 2 or 4 rows per thread, layout bit 0 = bank colouring, bit 1 = two record copies,
 * optional row permutation + order keys),
 * replays the kernels' walk on the CPU and compares it bit for bit with plain row sums.  Returns 0 when identical.
 * stats4 (nullable): mean LDS.128 bank-conflict degree, union steps per row, union entries per row, slots per tile. */
int mft_debug_tile_selftest(int64_t n, int k, int R, int layout, int with_perm, unsigned seed, double *stats4);
/* The same self test on a caller-supplied sparsity (the kNN table of a real cloud, or its transpose): n_rows stencils over n
 * columns, 0-based CSR, the columns of a row in summation order; pseudo-random dyadic weights.  layout bits as MFT_OPT_TILE >> 2
 * (1: bank-coloured slots, 2: two record copies, 4: tuned copies). */
int mft_debug_tile_selftest_csr(int64_t n, int64_t n_rows, const int64_t *rowptr, const int32_t *col, int R, int layout,
                                unsigned seed, double *stats6);  /* stats4 + STS.128 conflict degree of the phase-1 stores, copy 0 / copy 1 */

/* Host-only: the portable per-tile builder (the code the device runs, csrc/mft_tile_build.cuh) against the host builder on the
 * operator of mft_debug_tile_selftest / mft_debug_tile_selftest_csr: 0 when every array of the layout is identical. */
int mft_debug_tile_build_compare(int64_t n, int k, int layout, int with_perm, unsigned seed);
int mft_debug_tile_build_compare_csr(int64_t n, int64_t n_rows, const int64_t *rowptr, const int32_t *col, int layout, unsigned seed);
/* Host-only: the recursion-free bank-group feasibility test of the portable builder against the augmenting-path matcher of the
 * host builder (all instances with <= 3 points, then `trials` random ones with <= 8); *mismatches must come back 0. */
int mft_debug_matcher_compare(unsigned seed, int trials, long long *mismatches);
/* FNV-1a checksum (and size in bytes) of the union-tile layout a finalized context holds for the forward (0) or transposed (1)
 * operator: lets a test compare the device-built layout with the host-built one. */
int mft_debug_tiler_checksum(mft_ctx *ctx, int which, unsigned long long *fnv_out, long long *bytes_out);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) -------------------------------------------------
 * replaces MPICache + perform_halo_update! (src/domains/PointCloudDomain/ParallelPointCloud.jl:6-71,
 * src/auxiliary/mpi.jl:217-265) and the Allreduces of ode_mean/ode_maximum (mpi.jl:40-81). */
int mft_nccl_unique_id(void *id128);  /* rank 0 creates, the host program broadcasts the 128 bytes */
int mft_comm_init(mft_ctx *ctx, int nranks, int rank, const void *id128);
/* halo plan: for each peer p (rank id peers[p]) send owned points send_idx1[send_off[p]..send_off[p+1]) (1-based,
 * caller numbering) and receive recv_count[p] points into the halo tail, in peer order */
int mft_set_halo(mft_ctx *ctx, int npeers, const int *peers, const int64_t *send_off, const int64_t *send_idx1,
                 const int64_t *recv_count);

/* Peer-memory exchange over NVLink (alternative to the NCCL path, same results): ranks map each other's state
 * arrays with CUDA IPC and write halo blocks / partial norms straight into the peer's memory, flagged by epoch counters.
 * Kernel-only, hence a whole multi-GPU SSPRK step replays as one CUDA graph.
 *   1. every rank: mft_p2p_handles(ctx, buf192)            -> 3 x 64-byte IPC handles {u, g, flag window}
 *   2. host program all-gathers the handles (and, per peer, the row in the PEER's array where this rank's block starts:
 *      n_local(peer) + the peer's receive offset for this rank)
 *   3. every rank: mft_p2p_connect(ctx, nranks, rank, all_handles, peer_dst_row [one per peer of mft_set_halo], n_global) */
int mft_p2p_handles(mft_ctx *ctx, void *out3x64);
int mft_p2p_connect(mft_ctx *ctx, int nranks, int rank, const void *all_handles, const int64_t *peer_dst_row,
                    int64_t n_global);

#ifdef __cplusplus
}
#endif
#endif /* MFT_B200_H */
