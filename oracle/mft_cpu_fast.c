/* mft_cpu_fast.c -- TEST / MEASUREMENT INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * SURVEY.md section 8(d), CPU baseline (ii) "best-effort CPU": what a careful CPU implementation of the same rhs!
 * looks like -- NOT the reference's structure (that is mft_oracle.c, variant (i): CSC/Int64 operators, one SpMV per
 * variable per direction, serial loops, src/solvers/pointcloudsolver/rbfsolver.jl:247-265, src/sources/
 * hyperviscosity.jl:351-409).  Here: row-major operators (int32 columns, paired Dx/Dy weights), AoS state, ONE fused
 * row-parallel pass for flux divergence + D u + limiter + g and ONE for the adjoint apply, pthreads over rows on all
 * host cores (this image's gcc ships no OpenMP runtime).  Per-row sums run in the reference's order (ascending column, Dx terms then Dy terms, separate
 * multiply and add), so the result equals the oracle's up to the association of the global mean.
 *
 * Only bench.py's cpu_baseline leg and tests/ call this.  Compiled by oracle/Makefile with -O3 -pthread.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

typedef struct {
    int64_t n;
    const int64_t *ptr;   /* n+1: rows of D (ascending column inside a row)            */
    const int32_t *col;
    const double *wx, *wy;
    const int64_t *tptr;  /* n+1: rows of D' (ascending source row inside a row)       */
    const int32_t *tcol;
    const double *twx, *twy;
    int64_t nb;           /* Dirichlet boundary table (rows, V values each)            */
    const int32_t *bidx;
    const double *bval;
    double gamma, c_rv, c_uw, dx_avg;
    int success_iter_zero, mean_divisor_vn, max_lexicographic;
} fast_problem;

/* ---- a minimal parallel-for on pthreads (this image's gcc has no OpenMP runtime) ---------------------------------- */
static int g_threads = 0;
int fast_max_threads(void)
{
    if (g_threads <= 0) {
        const char *e = getenv("MFT_CPU_THREADS");
        long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
        g_threads = (int)(t < 1 ? 1 : (t > 256 ? 256 : t));
    }
    return g_threads;
}
void fast_set_threads(int t) { g_threads = t < 1 ? 1 : (t > 256 ? 256 : t); }

typedef void (*range_fn)(void *ctx, int tid, int64_t begin, int64_t end);
typedef struct {
    range_fn fn;
    void *ctx;
    int tid;
    int64_t begin, end;
} job;
static void *job_main(void *a)
{
    job *j = (job *)a;
    j->fn(j->ctx, j->tid, j->begin, j->end);
    return NULL;
}
static void parallel_for(int64_t n, range_fn fn, void *ctx)
{
    const int T = fast_max_threads();
    pthread_t th[256];
    job jobs[256];
    for (int t = 0; t < T; ++t) {
        jobs[t] = (job){fn, ctx, t, n * t / T, n * (t + 1) / T};
        if (t > 0) pthread_create(&th[t], NULL, job_main, &jobs[t]);
    }
    job_main(&jobs[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
}

static inline void prim(double gamma, const double *u, double *v1, double *v2, double *p)
{
    *v1 = u[1] / u[0];
    *v2 = u[2] / u[0];
    *p = (gamma - 1.0) * fma(-0.5, fma(u[1], *v1, u[2] * *v2), u[3]);
}

static inline double jl_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    return a > b ? a : b;
}

static inline int lex_less(const double *a, const double *b)
{
    for (int v = 0; v < 4; ++v) {
        if (a[v] < b[v]) return 1;
        if (a[v] > b[v]) return 0;
    }
    return 0;
}

typedef struct {
    const fast_problem *P;
    double *u, *du, *g;
    const double *approx_du;
    double mean[4], nrm[4];
    double part[256][4];
    int zero_du;
} rhs_ctx;

static void r_bc(void *c, int tid, int64_t b, int64_t e)
{
    rhs_ctx *x = (rhs_ctx *)c;
    (void)tid;
    for (int64_t j = b; j < e; ++j) {
        memcpy(x->u + 4 * (int64_t)x->P->bidx[j], x->P->bval + 4 * j, 4 * sizeof(double));
        if (x->zero_du) memset(x->du + 4 * (int64_t)x->P->bidx[j], 0, 4 * sizeof(double));
    }
}
static void r_sum(void *c, int tid, int64_t b, int64_t e)
{
    rhs_ctx *x = (rhs_ctx *)c;
    double s[4] = {0, 0, 0, 0};
    for (int64_t i = b; i < e; ++i)
        for (int v = 0; v < 4; ++v) s[v] += x->u[4 * i + v];
    memcpy(x->part[tid], s, sizeof s);
}
static void r_maxdev(void *c, int tid, int64_t b, int64_t e)
{
    rhs_ctx *x = (rhs_ctx *)c;
    double best[4] = {-1, -1, -1, -1};
    for (int64_t i = b; i < e; ++i) {
        double d[4];
        for (int v = 0; v < 4; ++v) d[v] = fabs(x->u[4 * i + v] - x->mean[v]);
        if (x->P->max_lexicographic) {
            if (lex_less(best, d)) memcpy(best, d, sizeof d);
        } else {
            for (int v = 0; v < 4; ++v) best[v] = jl_max(best[v], d[v]);
        }
    }
    memcpy(x->part[tid], best, sizeof best);
}
/* pass A: flux divergence, D u, limiter, g */
static void r_pass_a(void *c, int tid, int64_t rb, int64_t re)
{
    rhs_ctx *x = (rhs_ctx *)c;
    const fast_problem *P = x->P;
    const double gamma = P->gamma, *u = x->u, *nrm = x->nrm;
    (void)tid;
    for (int64_t i = rb; i < re; ++i) {
        double acc[4] = {0, 0, 0, 0}, gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0};
        const int64_t b = P->ptr[i], e = P->ptr[i + 1];
        for (int64_t q = b; q < e; ++q) {
            const double *uj = u + 4 * (int64_t)P->col[q];
            double v1, v2, p;
            prim(gamma, uj, &v1, &v2, &p);
            const double f[4] = {uj[1], fma(uj[1], v1, p), uj[1] * v2, (uj[3] + p) * v1};
            const double w = P->wx[q];
            for (int v = 0; v < 4; ++v) {
                acc[v] = acc[v] + w * (-f[v]);
                gx[v] = gx[v] + w * uj[v];
            }
        }
        for (int64_t q = b; q < e; ++q) {
            const double *uj = u + 4 * (int64_t)P->col[q];
            double v1, v2, p;
            prim(gamma, uj, &v1, &v2, &p);
            const double h[4] = {uj[2], uj[2] * v1, fma(uj[2], v2, p), (uj[3] + p) * v2};
            const double w = P->wy[q];
            for (int v = 0; v < 4; ++v) {
                acc[v] = acc[v] + w * (-h[v]);
                gy[v] = gy[v] + w * uj[v];
            }
        }
        const double *ui = u + 4 * i;
        double v1, v2, p;
        prim(gamma, ui, &v1, &v2, &p);
        const double speed = sqrt(v1 * v1 + v2 * v2);
        const double sound = (p < 0.0 || ui[0] < 0.0) ? 0.0 : sqrt(gamma * p / ui[0]);
        const double e_uw = P->c_uw * 0.5 * P->dx_avg * (speed + sound);
        double mx = fabs(x->approx_du[4 * i] - acc[0]) / nrm[0];
        for (int v = 1; v < 4; ++v) mx = jl_max(mx, fabs(x->approx_du[4 * i + v] - acc[v]) / nrm[v]);
        const double e_rv = 0.5 * P->c_rv * (P->dx_avg * P->dx_avg) * mx;
        double eps;
        if (isnan(e_rv) || isinf(e_rv) || P->success_iter_zero) {
            eps = (isnan(e_uw) || isinf(e_uw)) ? 2.220446049250313e-16 : e_uw;
        } else {
            eps = e_rv < e_uw ? e_rv : e_uw;
            if (e_uw != e_uw) eps = e_uw;
        }
        for (int v = 0; v < 4; ++v) {
            x->du[4 * i + v] = acc[v];
            x->g[8 * i + v] = eps * gx[v];
            x->g[8 * i + 4 + v] = eps * gy[v];
        }
    }
}
/* pass B: du -= Dx' gX + Dy' gY */
static void r_pass_b(void *c, int tid, int64_t rb, int64_t re)
{
    rhs_ctx *x = (rhs_ctx *)c;
    const fast_problem *P = x->P;
    (void)tid;
    for (int64_t i = rb; i < re; ++i) {
        double tx[4] = {0, 0, 0, 0}, ty[4] = {0, 0, 0, 0};
        for (int64_t q = P->tptr[i]; q < P->tptr[i + 1]; ++q) {
            const double *gj = x->g + 8 * (int64_t)P->tcol[q];
            const double a = P->twx[q], b2 = P->twy[q];
            for (int v = 0; v < 4; ++v) {
                tx[v] = tx[v] + a * gj[v];
                ty[v] = ty[v] + b2 * gj[4 + v];
            }
        }
        for (int v = 0; v < 4; ++v) x->du[4 * i + v] = (x->du[4 * i + v] + tx[v] * -1.0) + ty[v] * -1.0;
    }
}

/* one Euler + residual-viscosity rhs! on AoS arrays (4 doubles per point); u in/out (Dirichlet rows), du out;
 * g: scratch 8 doubles per point */
void fast_rhs_rv(const fast_problem *P, double *u, const double *approx_du, double *du, double *g)
{
    static rhs_ctx x;  /* part[] is large: keep it off the stack; one caller at a time (measurement code) */
    const int T = fast_max_threads();
    x.P = P;
    x.u = u;
    x.du = du;
    x.g = g;
    x.approx_du = approx_du;
    x.zero_du = 0;
    parallel_for(P->nb, r_bc, &x);                         /* BC pass 1 */
    parallel_for(P->n, r_sum, &x);                         /* ode_mean */
    const double len = P->mean_divisor_vn ? 4.0 * (double)P->n : (double)P->n;
    for (int v = 0; v < 4; ++v) {
        double s = 0.0;
        for (int t = 0; t < T; ++t) s += x.part[t][v];
        x.mean[v] = s / len;
    }
    parallel_for(P->n, r_maxdev, &x);                      /* ode_maximum(|u - mean|) */
    for (int v = 0; v < 4; ++v) x.nrm[v] = x.part[0][v];
    for (int t = 1; t < T; ++t) {
        if (P->max_lexicographic) {
            if (lex_less(x.nrm, x.part[t])) memcpy(x.nrm, x.part[t], 4 * sizeof(double));
        } else {
            for (int v = 0; v < 4; ++v) x.nrm[v] = jl_max(x.nrm[v], x.part[t][v]);
        }
    }
    for (int v = 0; v < 4; ++v)
        if (x.nrm[v] == 0.0) x.nrm[v] = 2.220446049250313e-16;
    parallel_for(P->n, r_pass_a, &x);
    parallel_for(P->n, r_pass_b, &x);
    x.zero_du = 1;
    parallel_for(P->nb, r_bc, &x);                         /* BC pass 2 */
}

void fast_rhs_rv_repeat(const fast_problem *P, double *u, const double *approx_du, double *du, double *g, int reps)
{
    for (int r = 0; r < reps; ++r) fast_rhs_rv(P, u, approx_du, du, g);
}
