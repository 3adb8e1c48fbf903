/*
 * mft_oracle.c -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * Plain-C CPU restatement of the MeshfreeTrixi.jl semidiscrete right-hand side `rhs!`
 * and the pieces around it (boundary passes, flux + sparse operator application,
 * stabilisation sources, time-history residual, SSPRK33 stage formulas).  It keeps the
 * reference's *execution structure*: SoA state (one vector per variable), CSC operators
 * with 1-based Int64 indices, one SpMV per variable per direction, serial loops.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  The CUDA library never links or calls it.
 *
 * Arithmetic contract (so the CUDA kernels can be compared bit-for-bit where possible):
 *   - compile with -ffp-contract=off; every fused multiply-add below is an explicit fma()
 *     and appears exactly where the reference's `@muladd` blocks (Trixi flux / cons2prim,
 *     history.jl) would form one.  SparseArrays' mul! is NOT under @muladd: separate * and +.
 *
 * Parity pins: the reference ships no golden vectors.  What pins this file
 * are the reference's own test identities, reproduced in tests/test_oracle_identities.py:
 *   test/divergence_test.jl:66-82, test/upwind_viscosity_test.jl:69-78, test/history_test.jl:14-32.
 * PARITY UNPINNED (no reference test, Julia not runnable here): update_residual_visc!,
 * update_visc!, both BC passes, whole-rhs!, time integration.  For those this restatement
 * is the reference.
 *
 * All file:line citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_EQ_EULER2D 0
#define ORC_EQ_ADVECTION2D 1

#define ORC_BC_DIRICHLET 0
#define ORC_BC_SLIP_WALL 1
#define ORC_BC_DO_NOTHING 2

#define ORC_SRC_HV_FLYER 0
#define ORC_SRC_HV_TOMINEC 1
#define ORC_SRC_UPWIND 2
#define ORC_SRC_RESIDUAL 3
#define ORC_SRC_IGR 4

/* Julia SparseMatrixCSC{Float64,Int64}: colptr (n+1), rowval (nnz), both 1-based. */
typedef struct {
    const int64_t *colptr;
    const int64_t *rowval;
    const double *nzval;
} orc_csc;

/* One boundary group (BoundaryData, src/domains/PointCloudDomain/geometry_primatives.jl:361-364). */
typedef struct {
    int32_t kind;
    int32_t pad_;
    int64_t n;
    const int64_t *idx;    /* 1-based point indices */
    const double *normals; /* n x 2, row-major (nx, ny) */
    const double *values;  /* Dirichlet table, SoA: values[v*n + j]; NULL otherwise */
} orc_bc;

/* One source term (callable structs of src/sources/hyperviscosity.jl). */
typedef struct {
    int32_t kind;
    int32_t mean_divisor_vn; /* 1: ode_mean divides by V*N (recursive_length, src/auxiliary/mpi.jl:42); 0: by N */
    int32_t max_lexicographic; /* 1: maximum(::StructArray{SVector}) = lexicographic max; 0: per-component max */
    int32_t pad_;
    orc_csc hv;             /* HV sources: the single hyperviscosity matrix */
    double gamma;           /* HV sources */
    double c_rv, c_uw, dx_avg;
    int64_t success_iter;
    double *eps_uw, *eps_rv, *eps; /* N each */
    int64_t *eps_c;                /* N */
    double *residual;              /* V x N SoA */
    double *approx_du;             /* V x N SoA */
    /* SourceIGR (src/sources/IGR.jl) */
    double igr_alpha;
    int64_t igr_maxiter;
    double *igr_sigma; /* N   : cache.sigma */
    double *igr_work;  /* >= 9N : rho_inv, igr_rhs, then {flux_y (4N)} reused as {tmp1, tmp2, tmp3, r, p, c, products} */
    int64_t igr_iters; /* out: CG iterations performed */
    double igr_res;    /* out: final |r| */
} orc_source;

typedef struct {
    int64_t n;
    int32_t nvars;
    int32_t eq;
    double eqp[2]; /* Euler: gamma ; advection: a1, a2 */
    orc_csc D[2];
    int32_t nbc, nsrc;
    orc_bc *bcs;
    orc_source *srcs;
    double *scratch_a; /* V x N : local_values_threaded[1] */
    double *scratch_b; /* V x N */
} orc_problem;

/* ------------------------------------------------------------------------------------------
 * SparseArrays mul! restatements (call sites
 * src/solvers/pointcloudsolver/rbfsolver.jl:14-24).  Julia stdlib, not under @muladd.
 * ---------------------------------------------------------------------------------------- */

/* C += alpha * A * x  (forward CSC; per output row the terms arrive in ascending column order) */
void orc_spmv_csc_accum(int64_t n, const orc_csc *A, const double *x, double alpha, double *C)
{
    for (int64_t col = 0; col < n; ++col) {
        const double ax = x[col] * alpha;
        for (int64_t p = A->colptr[col] - 1; p < A->colptr[col + 1] - 1; ++p) {
            const int64_t r = A->rowval[p] - 1;
            const double prod = A->nzval[p] * ax;
            C[r] = C[r] + prod;
        }
    }
}

/* C = A * x  (3-arg mul!: beta = 0 zero-fills, alpha = true) */
void orc_spmv_csc(int64_t n, const orc_csc *A, const double *x, double *C)
{
    for (int64_t i = 0; i < n; ++i) C[i] = 0.0;
    orc_spmv_csc_accum(n, A, x, 1.0, C);
}

/* C += alpha * A' * x  (adjoint CSC: tmp = sum over the column, then C[col] += tmp*alpha) */
void orc_spmv_csc_adj_accum(int64_t n, const orc_csc *A, const double *x, double alpha, double *C)
{
    for (int64_t col = 0; col < n; ++col) {
        double tmp = 0.0;
        for (int64_t p = A->colptr[col] - 1; p < A->colptr[col + 1] - 1; ++p) {
            const double prod = A->nzval[p] * x[A->rowval[p] - 1];
            tmp = tmp + prod;
        }
        const double t2 = tmp * alpha;
        C[col] = C[col] + t2;
    }
}

/* ------------------------------------------------------------------------------------------
 * Trixi physics (third-party, Trixi 0.6-0.7.5).  Inside @muladd.
 * call sites: src/solvers/pointcloudsolver/rbfsolver.jl:259, src/sources/hyperviscosity.jl:254
 * ---------------------------------------------------------------------------------------- */
static inline void euler_pressure_velocity(double gamma, double rho, double m1, double m2, double E,
                                           double *v1, double *v2, double *p)
{
    *v1 = m1 / rho;
    *v2 = m2 / rho;
    const double s = fma(m1, *v1, m2 * (*v2));   /* rho_v1*v1 + rho_v2*v2 */
    const double e = fma(-0.5, s, E);            /* rho_e - 0.5*(...)      */
    *p = (gamma - 1.0) * e;
}

void orc_flux(const orc_problem *P, int dim, const double *u, double *f)
{
    const int64_t n = P->n;
    if (P->eq == ORC_EQ_EULER2D) {
        const double gamma = P->eqp[0];
        for (int64_t e = 0; e < n; ++e) {
            const double rho = u[e], m1 = u[n + e], m2 = u[2 * n + e], E = u[3 * n + e];
            double v1, v2, p;
            euler_pressure_velocity(gamma, rho, m1, m2, E, &v1, &v2, &p);
            if (dim == 0) {
                f[e] = m1;
                f[n + e] = fma(m1, v1, p);
                f[2 * n + e] = m1 * v2;
                f[3 * n + e] = (E + p) * v1;
            } else {
                f[e] = m2;
                f[n + e] = m2 * v1;
                f[2 * n + e] = fma(m2, v2, p);
                f[3 * n + e] = (E + p) * v2;
            }
        }
    } else {
        const double a = P->eqp[dim];
        for (int64_t e = 0; e < n; ++e) f[e] = a * u[e];
    }
}

/* ------------------------------------------------------------------------------------------
 * Boundary pass: calc_boundary_flux! / calc_single_boundary_flux!
 * src/solvers/pointcloudsolver/rbfsolver.jl:277-318 ; functors src/equations/PointCloudBCs.jl
 * ---------------------------------------------------------------------------------------- */
void orc_boundary_pass(const orc_problem *P, double *u, double *du)
{
    const int64_t n = P->n;
    const int V = P->nvars;
    for (int g = 0; g < P->nbc; ++g) {
        const orc_bc *bc = &P->bcs[g];
        for (int64_t j = 0; j < bc->n; ++j) {
            const int64_t b = bc->idx[j] - 1;
            if (bc->kind == ORC_BC_DIRICHLET) {
                /* PointCloudBCs.jl:49-63: (FluxZero() = zeros, boundary_value_function(x,t)) */
                for (int v = 0; v < V; ++v) {
                    du[v * n + b] = 0.0;
                    u[v * n + b] = bc->values[v * bc->n + j];
                }
            } else if (bc->kind == ORC_BC_SLIP_WALL) {
                /* PointCloudBCs.jl:87-106 and apply_slip_velocity :15-21 (Euler 2-D only) */
                const double nx0 = bc->normals[2 * j], ny0 = bc->normals[2 * j + 1];
                const double nrm = sqrt(nx0 * nx0 + ny0 * ny0);
                const double nx = nx0 / nrm, ny = ny0 / nrm;
                const double m1 = u[n + b], m2 = u[2 * n + b];
                const double vdotn = m1 * nx + m2 * ny;
                const double s1 = vdotn * nx, s2 = vdotn * ny;
                u[n + b] = m1 - s1;
                u[2 * n + b] = m2 - s2;
                du[n + b] = 0.0;
                du[2 * n + b] = 0.0;
            } else {
                /* BoundaryConditionDoNothing, PointCloudBCs.jl:108-115 */
            }
        }
    }
}

/* calc_fluxes!  src/solvers/pointcloudsolver/rbfsolver.jl:247-265 */
void orc_calc_fluxes(const orc_problem *P, const double *u, double *du)
{
    const int64_t n = P->n;
    double *flux_values = P->scratch_a;
    for (int dim = 0; dim < 2; ++dim) {
        orc_flux(P, dim, u, flux_values);
        for (int v = 0; v < P->nvars; ++v)
            orc_spmv_csc_accum(n, &P->D[dim], flux_values + v * n, -1.0, du + v * n);
    }
}

/* ------------------------------------------------------------------------------------------
 * Reductions used by residual viscosity: ode_mean / ode_maximum, src/auxiliary/mpi.jl:40-81
 * Base.sum = pairwise mapreduce with 1024-element leaves (Julia Base; the @simd inside a leaf
 * may reassociate in the real thing -- unpinnable, flagged in DESIGN.md).
 * ---------------------------------------------------------------------------------------- */
static double pairwise_sum(const double *a, int64_t ifirst, int64_t ilast)
{
    if (ifirst == ilast) return a[ifirst];
    if (ilast - ifirst < 1024) {
        double v = a[ifirst] + a[ifirst + 1];
        for (int64_t i = ifirst + 2; i <= ilast; ++i) v = v + a[i];
        return v;
    }
    const int64_t imid = ifirst + ((ilast - ifirst) >> 1);
    const double v1 = pairwise_sum(a, ifirst, imid);
    const double v2 = pairwise_sum(a, imid + 1, ilast);
    return v1 + v2;
}

double orc_sum(const double *a, int64_t n)
{
    if (n == 0) return 0.0;
    if (n < 16) {
        double v = a[0];
        for (int64_t i = 1; i < n; ++i) v = v + a[i];
        return v;
    }
    return pairwise_sum(a, 0, n - 1);
}

/* NaN-propagating max, as Julia's max(::Float64, ::Float64) */
static inline double jl_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}

/* src/sources/hyperviscosity.jl:246-285 */
void orc_update_upwind_visc(const orc_problem *P, const orc_source *S, const double *u)
{
    const int64_t n = P->n;
    const double gamma = P->eqp[0];
    for (int64_t i = 0; i < n; ++i) {
        double rho = u[i], v1, v2, p;
        euler_pressure_velocity(gamma, rho, u[n + i], u[2 * n + i], u[3 * n + i], &v1, &v2, &p);
        const double speed = sqrt(v1 * v1 + v2 * v2);
        double sound;
        if (p < 0.0 || rho < 0.0) {
            sound = 0.0;
        } else {
            sound = sqrt(gamma * p / rho);
        }
        S->eps_uw[i] = S->c_uw * 0.5 * S->dx_avg * (speed + sound);
    }
}

/* src/sources/hyperviscosity.jl:289-329 */
void orc_update_residual_visc(const orc_problem *P, const orc_source *S, const double *du, const double *u,
                              double *norms_out /* 4, optional */)
{
    const int64_t n = P->n;
    const int V = P->nvars; /* 4 */
    double mean_u[4], nrm[4];
    for (int v = 0; v < V; ++v)
        for (int64_t i = 0; i < n; ++i) S->residual[v * n + i] = fabs(S->approx_du[v * n + i] - du[v * n + i]);
    const double len = S->mean_divisor_vn ? (double)(V * n) : (double)n;
    for (int v = 0; v < V; ++v) mean_u[v] = orc_sum(u + v * n, n) / len;
    if (S->max_lexicographic) {
        /* maximum over SVectors compares with isless(::AbstractVector, ::AbstractVector) = lexicographic */
        double best[4];
        for (int v = 0; v < V; ++v) best[v] = fabs(u[v * n] - mean_u[v]);
        for (int64_t i = 1; i < n; ++i) {
            double c[4];
            for (int v = 0; v < V; ++v) c[v] = fabs(u[v * n + i] - mean_u[v]);
            int less = 0; /* is best < c lexicographically ? */
            for (int v = 0; v < V; ++v) {
                if (best[v] < c[v]) { less = 1; break; }
                if (best[v] > c[v]) { less = 0; break; }
            }
            if (less) for (int v = 0; v < V; ++v) best[v] = c[v];
        }
        for (int v = 0; v < V; ++v) nrm[v] = best[v];
    } else {
        for (int v = 0; v < V; ++v) {
            double m = fabs(u[v * n] - mean_u[v]);
            for (int64_t i = 1; i < n; ++i) m = jl_max(m, fabs(u[v * n + i] - mean_u[v]));
            nrm[v] = m;
        }
    }
    for (int v = 0; v < V; ++v)
        if (nrm[v] == 0.0) nrm[v] = 2.220446049250313e-16; /* eps() */
    if (norms_out)
        for (int v = 0; v < V; ++v) norms_out[v] = nrm[v];
    const double h = S->dx_avg;
    for (int64_t i = 0; i < n; ++i) {
        double mx = S->residual[i] / nrm[0];
        for (int v = 1; v < V; ++v) mx = jl_max(mx, S->residual[v * n + i] / nrm[v]);
        S->eps_rv[i] = 0.5 * S->c_rv * (h * h) * mx;
    }
}

/* src/sources/hyperviscosity.jl:331-349 */
void orc_update_visc(const orc_problem *P, const orc_source *S)
{
    for (int64_t i = 0; i < P->n; ++i) {
        const double rv = S->eps_rv[i], uw = S->eps_uw[i];
        if (isnan(rv) || isinf(rv) || S->success_iter == 0) {
            if (isnan(uw) || isinf(uw)) {
                S->eps[i] = 2.220446049250313e-16;
                S->eps_c[i] = 2;
            } else {
                S->eps[i] = uw;
                S->eps_c[i] = 1;
            }
        } else {
            S->eps[i] = rv < uw ? rv : uw; /* min(eps_rv, eps_uw); uw may be NaN -> Julia min gives NaN */
            if (uw != uw) S->eps[i] = uw;
            S->eps_c[i] = rv < uw ? 0 : 1;
        }
    }
}

/* Dissipation  du += -Dx'(eps .* (Dx u)) - Dy'(eps .* (Dy u)) ; hyperviscosity.jl:364-379, 393-408 */
static void apply_eps_dissipation(const orc_problem *P, const orc_source *S, const double *u, double *du)
{
    const int64_t n = P->n;
    double *local_u = P->scratch_a;
    for (int dim = 0; dim < 2; ++dim) {
        for (int v = 0; v < P->nvars; ++v) orc_spmv_csc(n, &P->D[dim], u + v * n, local_u + v * n);
        for (int v = 0; v < P->nvars; ++v)
            for (int64_t i = 0; i < n; ++i) local_u[v * n + i] = S->eps[i] * local_u[v * n + i];
        for (int v = 0; v < P->nvars; ++v)
            orc_spmv_csc_adj_accum(n, &P->D[dim], local_u + v * n, -1.0, du + v * n);
    }
}

/* ------------------------------------------------------------------------------------------
 * SourceIGR, src/sources/IGR.jl (outside every @muladd scope: separate * and +).
 * PARITY UNPINNED: no reference test exercises it, and its linear solver is third party.
 * ---------------------------------------------------------------------------------------- */

/* mul!(y, A::IGRCompositeMatrix, v)  IGR.jl:55-70 */
static void igr_composite_mul(const orc_problem *P, const orc_source *S, const double *rho_inv, double *tmp1, double *tmp2,
                              double *tmp3, const double *v, double *y)
{
    const int64_t n = P->n;
    orc_spmv_csc(n, &P->D[0], v, tmp1);
    orc_spmv_csc(n, &P->D[1], v, tmp2);
    for (int64_t i = 0; i < n; ++i) tmp1[i] = rho_inv[i] * tmp1[i];
    for (int64_t i = 0; i < n; ++i) tmp2[i] = rho_inv[i] * tmp2[i];
    orc_spmv_csc(n, &P->D[0], tmp1, tmp3);
    orc_spmv_csc(n, &P->D[1], tmp2, tmp1);
    for (int64_t i = 0; i < n; ++i) y[i] = rho_inv[i] * v[i];
    for (int64_t i = 0; i < n; ++i) y[i] = y[i] - S->igr_alpha * (tmp3[i] + tmp1[i]);
}

/* norm / dot are BLAS calls in the reference (summation order unspecified); pairwise here */
static double igr_dot(const double *a, const double *b, double *scratch, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) scratch[i] = a[i] * b[i];
    return orc_sum(scratch, n);
}

/* IterativeSolvers.cg!(x, A, b; maxiter) (third party, unpinned): abstol = 0, reltol = sqrt(eps), no preconditioner,
 * initially_zero = false.  x is zero on entry here (update_sigma! zeroes sigma first, IGR.jl:181). */
static void igr_cg(const orc_problem *P, orc_source *S, const double *rho_inv, const double *b, double *x, double *w)
{
    const int64_t n = P->n;
    double *tmp1 = w, *tmp2 = w + n, *tmp3 = w + 2 * n, *r = w + 3 * n, *p = w + 4 * n, *c = w + 5 * n, *scr = w + 6 * n;
    for (int64_t i = 0; i < n; ++i) r[i] = b[i];
    igr_composite_mul(P, S, rho_inv, tmp1, tmp2, tmp3, x, c);
    for (int64_t i = 0; i < n; ++i) r[i] = r[i] - c[i];
    for (int64_t i = 0; i < n; ++i) p[i] = 0.0;
    double residual = sqrt(igr_dot(r, r, scr, n)), prev_residual = 1.0;
    const double tol = fmax(1.4901161193847656e-08 * residual, 0.0);
    int64_t it = 0;
    while (!(residual <= tol) && it < S->igr_maxiter) {
        const double beta = (residual * residual) / (prev_residual * prev_residual);
        for (int64_t i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
        igr_composite_mul(P, S, rho_inv, tmp1, tmp2, tmp3, p, c);
        const double alpha = (residual * residual) / igr_dot(p, c, scr, n);
        for (int64_t i = 0; i < n; ++i) x[i] = x[i] + alpha * p[i];
        for (int64_t i = 0; i < n; ++i) r[i] = r[i] - alpha * c[i];
        prev_residual = residual;
        residual = sqrt(igr_dot(r, r, scr, n));
        ++it;
    }
    S->igr_iters = it;
    S->igr_res = residual;
}

static void igr_apply(const orc_problem *P, orc_source *S, const double *u, double *du)
{
    const int64_t n = P->n;
    const double gamma = P->eqp[0];
    double *rho_inv = S->igr_work, *rhs = S->igr_work + n, *w = S->igr_work + 2 * n;
    double *u_prim = P->scratch_a, *flux_x = P->scratch_b; /* flux_y: w (4N needed -> use w..w+4N before the CG uses it) */
    double *flux_y = w;
    /* update_igr_rhs! :117-158 */
    for (int64_t i = 0; i < n; ++i) {
        double v1, v2, p;
        euler_pressure_velocity(gamma, u[i], u[n + i], u[2 * n + i], u[3 * n + i], &v1, &v2, &p);
        u_prim[i] = u[i];
        u_prim[n + i] = v1;
        u_prim[2 * n + i] = v2;
        u_prim[3 * n + i] = p;
    }
    for (int v = 0; v < 4; ++v) orc_spmv_csc(n, &P->D[0], u_prim + v * n, flux_x + v * n);
    for (int v = 0; v < 4; ++v) orc_spmv_csc(n, &P->D[1], u_prim + v * n, flux_y + v * n);
    for (int64_t i = 0; i < n; ++i) {
        double trace = 0.0;
        trace = trace + flux_x[i];         /* flux_x[i][1] */
        trace = trace + flux_y[n + i];     /* flux_y[i][2] */
        const double fx2 = flux_x[n + i], fx3 = flux_x[2 * n + i], fy2 = flux_y[n + i], fy3 = flux_y[2 * n + i];
        const double trace_squared = (fx2 * fx2 + (2.0 * fy2) * fx3) + fy3 * fy3;
        rhs[i] = S->igr_alpha * (trace * trace + trace_squared);
    }
    /* update_sigma! :169-191 */
    for (int64_t i = 0; i < n; ++i) S->igr_sigma[i] = 0.0;
    for (int64_t i = 0; i < n; ++i) rho_inv[i] = 1.0 / u[i];
    igr_cg(P, S, rho_inv, rhs, S->igr_sigma, w);
    /* flux_igr + mul_by_accum!(D[i], -1) :225-238 (fields with a zero flux receive w * (0 * -1): no change) */
    orc_spmv_csc_accum(n, &P->D[0], S->igr_sigma, -1.0, du + n);
    orc_spmv_csc_accum(n, &P->D[1], S->igr_sigma, -1.0, du + 2 * n);
}

void orc_apply_source(const orc_problem *P, orc_source *S, const double *u, double *du)
{
    const int64_t n = P->n;
    switch (S->kind) {
    case ORC_SRC_HV_FLYER:   /* hyperviscosity.jl:52-64  */
    case ORC_SRC_HV_TOMINEC: /* hyperviscosity.jl:121-134 */
        for (int v = 0; v < P->nvars; ++v)
            orc_spmv_csc_accum(n, &S->hv, u + v * n, -S->gamma, du + v * n);
        break;
    case ORC_SRC_UPWIND: /* hyperviscosity.jl:351-380 */
        orc_update_upwind_visc(P, S, u);
        for (int64_t i = 0; i < n; ++i) {
            S->eps[i] = S->eps_uw[i];
            S->eps_c[i] = 1;
        }
        apply_eps_dissipation(P, S, u, du);
        break;
    case ORC_SRC_RESIDUAL: /* hyperviscosity.jl:382-409 */
        orc_update_upwind_visc(P, S, u);
        orc_update_residual_visc(P, S, du, u, NULL);
        orc_update_visc(P, S);
        apply_eps_dissipation(P, S, u, du);
        break;
    case ORC_SRC_IGR: /* IGR.jl:211-239 */
        igr_apply(P, S, u, du);
        break;
    }
}

/* Trixi.rhs!  src/solvers/pointcloudsolver/rbfsolver.jl:397-428 ; u is IN/OUT */
void orc_rhs(const orc_problem *P, double *u, double *du)
{
    const int64_t n = P->n;
    for (int64_t i = 0; i < n * P->nvars; ++i) du[i] = 0.0; /* reset_du! :118-124 */
    orc_boundary_pass(P, u, du);
    orc_calc_fluxes(P, u, du);
    for (int s = 0; s < P->nsrc; ++s) orc_apply_source(P, &P->srcs[s], u, du); /* calc_sources! :388-395 */
    orc_boundary_pass(P, u, du);
}

/* ------------------------------------------------------------------------------------------
 * Time history: src/callbacks_step/history.jl:105-152  (inside @muladd)
 * ---------------------------------------------------------------------------------------- */

/* shift_soln_history! :105-111.  sol_history is V x (N x (polydeg+1)), slot-major per variable:
 * sol_history[v][slot*n + i]; slot 0 = most recent. */
void orc_shift_soln_history(int64_t n, int nvars, int nslots, double *time_history, double *sol_history,
                            double t, const double *u)
{
    for (int s = nslots - 1; s >= 1; --s) time_history[s] = time_history[s - 1];
    time_history[0] = t;
    for (int v = 0; v < nvars; ++v) {
        double *h = sol_history + (int64_t)v * n * nslots;
        memmove(h + n, h, sizeof(double) * (size_t)n * (size_t)(nslots - 1));
        memcpy(h, u + (int64_t)v * n, sizeof(double) * (size_t)n);
    }
}

/* time_deriv_weights! :131-152.  w = scale * (A' \ b'),  A[:,k] = t_^(k-1), b[k] = (k-1) t_eval^(k-2).
 * `\` on a dense square matrix = LU with partial pivoting (LAPACK getrf); restated as right-looking
 * Gaussian elimination with row pivoting. */
void orc_time_deriv_weights(int m, const double *t, double *w)
{
    double maxabs = 0.0;
    for (int i = 0; i < m; ++i) maxabs = fmax(maxabs, fabs(t[i]));
    const double scale = 1.0 / maxabs;
    double ts[16], M[16 * 16], b[16];
    for (int i = 0; i < m; ++i) ts[i] = t[i] * scale;
    const double t_eval = ts[0];
    /* M = A' : M[k][i] = ts[i]^k */
    for (int k = 0; k < m; ++k) {
        for (int i = 0; i < m; ++i) M[k * m + i] = pow(ts[i], (double)k);
        b[k] = (double)k * pow(t_eval, (double)(k - 1));
    }
    int piv[16];
    for (int c = 0; c < m; ++c) {
        int p = c;
        double best = fabs(M[c * m + c]);
        for (int r = c + 1; r < m; ++r)
            if (fabs(M[r * m + c]) > best) { best = fabs(M[r * m + c]); p = r; }
        piv[c] = p;
        if (p != c) {
            for (int j = 0; j < m; ++j) { double tmp = M[c * m + j]; M[c * m + j] = M[p * m + j]; M[p * m + j] = tmp; }
            double tb = b[c]; b[c] = b[p]; b[p] = tb;
        }
        for (int r = c + 1; r < m; ++r) {
            const double l = M[r * m + c] / M[c * m + c];
            M[r * m + c] = l;
            for (int j = c + 1; j < m; ++j) M[r * m + j] = M[r * m + j] - l * M[c * m + j];
            b[r] = b[r] - l * b[c];
        }
    }
    for (int r = m - 1; r >= 0; --r) {
        double s = b[r];
        for (int j = r + 1; j < m; ++j) s = s - M[r * m + j] * b[j];
        b[r] = s / M[r * m + r];
    }
    for (int i = 0; i < m; ++i) w[i] = scale * b[i];
    (void)piv;
}

/* update_approx_du! :113-129 */
void orc_update_approx_du(int64_t n, int nvars, int nslots, double *approx_du, double *time_weights,
                          const double *time_history, const double *sol_history, int64_t success_iter,
                          int approx_order)
{
    for (int64_t i = 0; i < n * nvars; ++i) approx_du[i] = 0.0;
    int64_t ntp = success_iter + 1 < approx_order + 1 ? success_iter + 1 : approx_order + 1;
    if (success_iter > 0) {
        orc_time_deriv_weights((int)ntp, time_history, time_weights);
        for (int s = 0; s < ntp; ++s)
            for (int v = 0; v < nvars; ++v) {
                const double *h = sol_history + (int64_t)v * n * nslots + (int64_t)s * n;
                double *a = approx_du + (int64_t)v * n;
                for (int64_t i = 0; i < n; ++i) a[i] = fma(time_weights[s], h[i], a[i]);
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * SSPRK33 (Shu-Osher) stage formulas, SURVEY.md Appendix B.5 (OrdinaryDiffEq is third-party and
 * unpinned; only call site rbfsolver_test.jl:104-107).  The stage expressions below DEFINE the
 * arithmetic both this oracle and the CUDA stage-update kernel use:
 *   s=1: u = fma(dt, k, uprev)
 *   s=2: u = fma(dt, k, fma(3, uprev, u)) / 4
 *   s=3: u = fma(2dt, k, fma(2, u, uprev)) / 3
 * ---------------------------------------------------------------------------------------- */
void orc_ssprk33_stage(int64_t len, int stage, double dt, const double *uprev, const double *k, double *u)
{
    if (stage == 1) {
        for (int64_t i = 0; i < len; ++i) u[i] = fma(dt, k[i], uprev[i]);
    } else if (stage == 2) {
        for (int64_t i = 0; i < len; ++i) u[i] = fma(dt, k[i], fma(3.0, uprev[i], u[i])) / 4.0;
    } else {
        const double dt2 = 2.0 * dt;
        for (int64_t i = 0; i < len; ++i) u[i] = fma(dt2, k[i], fma(2.0, u[i], uprev[i])) / 3.0;
    }
}

/* ------------------------------------------------------------------------------------------
 * SSPRK43 (OrdinaryDiffEq low-storage form, third party / unpinned; the integrator the
 * reference names: rbfsolver_test.jl:104-107).  With h = dt/2 and k = f(u):
 *   stage 1: u = fma(h, k, uprev)                 then k = f(u, t+h)
 *   stage 2: u = fma(h, k, u)                     then k = f(u, t+dt)
 *   stage 3: u = fma(h, k, u); utilde = fma(2,u,uprev)/3; u = fma(2,uprev,u)/3      then k = f(u, t+h)
 *   stage 4: u = fma(h, k, u); utilde = 0.5*(utilde - u)                            then k = f(u, t+dt)  (FSAL)
 * Error estimate: EEst = sqrt( sum_i (utilde_i / (abstol + max(|uprev_i|,|u_i|)*reltol))^2 / len )   (calculate_residuals
 * + ODE_DEFAULT_NORM / ode_norm, src/auxiliary/mpi.jl:15-19: RMS over all scalars).
 * ---------------------------------------------------------------------------------------- */
void orc_ssprk43_stage(int64_t len, int stage, double dt, const double *uprev, const double *k, double *u, double *utilde)
{
    const double h = dt / 2.0;
    if (stage == 1) {
        for (int64_t i = 0; i < len; ++i) u[i] = fma(h, k[i], uprev[i]);
    } else if (stage == 2) {
        for (int64_t i = 0; i < len; ++i) u[i] = fma(h, k[i], u[i]);
    } else if (stage == 3) {
        for (int64_t i = 0; i < len; ++i) {
            const double u3 = fma(h, k[i], u[i]);
            utilde[i] = fma(2.0, u3, uprev[i]) / 3.0;
            u[i] = fma(2.0, uprev[i], u3) / 3.0;
        }
    } else {
        for (int64_t i = 0; i < len; ++i) {
            const double un = fma(h, k[i], u[i]);
            u[i] = un;
            utilde[i] = 0.5 * (utilde[i] - un);
        }
    }
}

double orc_error_sumsq(int64_t len, const double *utilde, const double *uprev, const double *u, double abstol, double reltol)
{
    double s = 0.0;
    for (int64_t i = 0; i < len; ++i) {
        const double a = fabs(uprev[i]), b = fabs(u[i]);
        const double r = utilde[i] / (abstol + (a > b ? a : b) * reltol);
        s += r * r;
    }
    return s;
}

/* ------------------------------------------------------------------------------------------
 * Timed loop for the CPU baseline (bench.py cpu_baseline / --impl reference): `reps` rhs!
 * evaluations on the same state, exactly the reference's serial execution structure.
 * ---------------------------------------------------------------------------------------- */
void orc_rhs_repeat(const orc_problem *P, double *u, double *du, int reps)
{
    for (int r = 0; r < reps; ++r) orc_rhs(P, u, du);
}

/* ------------------------------------------------------------------------------------------
 * Zhang-Shu positivity limiter on the point cloud:
 * src/callbacks_stage/positivity_zhang_shu_point2d.jl:22-82 (one (threshold, variable) pair; the recursion over pairs
 * of positivity_zhang_shu.jl:50-72 is done by the caller).  nbr: n x k row-major, 0-based, in the kNN list order.
 * variable 0 = Trixi.density, 1 = Trixi.pressure (third party: (gamma-1)*(rho_e - 0.5*(rho_v1^2+rho_v2^2)/rho), no
 * product that @muladd can fuse).  The blend theta*u + (1-theta)*u_mean is inside @muladd; StaticArrays maps
 * muladd(scalar, SVector, SVector) to per-component muladd -> fma.  PARITY UNPINNED (no reference test uses it).
 * local_u, u_mean: scratch, 4n each (SoA like u). ---------------------------------------------------------------- */
static double zs_variable(int variable, double gamma, double rho, double m1, double m2, double E)
{
    if (variable == 0) return rho;
    return (gamma - 1.0) * (E - 0.5 * (m1 * m1 + m2 * m2) / rho);
}
static double jl_min(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return a < b ? a : b;
}
void orc_limiter_zhang_shu(int64_t n, int k, const int64_t *nbr, double gamma, double threshold, int variable, double *u,
                           double *local_u, double *u_mean)
{
    for (int64_t i = 0; i < 4 * n; ++i) local_u[i] = u_mean[i] = 0.0; /* set_to_zero! :28-29 */
    for (int64_t e = 0; e < n; ++e) {
        double value_min = INFINITY; /* typemax */
        for (int q = 0; q < k; ++q) {
            const int64_t i = nbr[e * k + q];
            value_min = jl_min(value_min, zs_variable(variable, gamma, u[i], u[n + i], u[2 * n + i], u[3 * n + i]));
        }
        if (!(value_min < threshold)) continue;
        for (int q = 0; q < k; ++q) {
            const int64_t i = nbr[e * k + q];
            for (int v = 0; v < 4; ++v) u_mean[v * n + e] = u_mean[v * n + e] + u[v * n + i];
        }
        for (int v = 0; v < 4; ++v) u_mean[v * n + e] = u_mean[v * n + e] / (double)k;
        const double value_mean = zs_variable(variable, gamma, u_mean[e], u_mean[n + e], u_mean[2 * n + e], u_mean[3 * n + e]);
        const double theta = (value_mean - threshold) / (value_mean - value_min);
        for (int v = 0; v < 4; ++v) local_u[v * n + e] = fma(theta, u[v * n + e], (1.0 - theta) * u_mean[v * n + e]);
    }
    for (int64_t e = 0; e < n; ++e) {
        int nonzero = 0;
        for (int v = 0; v < 4; ++v) nonzero |= local_u[v * n + e] != 0.0;
        if (nonzero)
            for (int v = 0; v < 4; ++v) u[v * n + e] = local_u[v * n + e];
    }
}
