"""Generates tests/golden/fixture_golden_f.npz: golden vectors for the "next" rows on the
reference's fixture cloud, from the CPU oracle -- Zhang-Shu limiter outputs and SourceIGR (sigma, du).  The operators are
the ones stored in fixture_golden.npz, so the vectors do not depend on the LAPACK build.  See make_golden.py for why the
oracle (and not the reference itself) is the source.  Re-run:  python tests/golden/make_golden_f.py"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
from cases import orc  # noqa: E402


def golden_ops():
    G = np.load(os.path.join(HERE, "fixture_golden.npz"))
    nb = G["neighbors"].astype(np.int64)
    n, k = nb.shape
    rows = np.repeat(np.arange(n), k)
    ops = []
    for w in (G["wx"], G["wy"]):
        A = sp.coo_matrix((w.reshape(-1), (rows, nb.reshape(-1))), shape=(n, n)).tocsc()
        A.sort_indices()
        ops.append(A)
    return nb, ops


def limiter_state(pts):
    u = cases.ic_smooth_euler(pts, 0.0)
    u[0, 100:130] *= 0.01
    u[3, 500:520] *= 0.2
    u[3, 900] = 0.01
    return u


def table_csc(nb, w):
    n, k = nb.shape
    A = sp.coo_matrix((np.asarray(w).reshape(-1), (np.repeat(np.arange(n), k), nb.reshape(-1))), shape=(n, n)).tocsc()
    A.sort_indices()
    return A


def advection_matrices(nb, G):
    """[Dx, Dy], Flyer H, Tominec H = lap' * lap from the stored r^5 tables"""
    D = [table_csc(nb, G["wx5"]), table_csc(nb, G["wy5"])]
    lap = table_csc(nb, G["lap"])
    return D, table_csc(nb, G["h4"]), sp.csc_matrix(lap.T @ lap)


def advection_problem(fx5, nb, G):
    """oracle problem of configs[0] on the stored tables; BC set of the config: inlet Dirichlet(IC), others do-nothing"""
    D, Hf, Ht = advection_matrices(nb, G)
    gam = (1.0 * fx5["dx_min"] ** 4, 1.0 * fx5["dx_min"] ** 4.5)       # hyperviscosity.jl:47, :116
    srcs = [orc.OracleSource(kind=orc.SRC_HV_FLYER, hv=orc.JuliaCSC(Hf), gamma=gam[0]),
            orc.OracleSource(kind=orc.SRC_HV_TOMINEC, hv=orc.JuliaCSC(Ht), gamma=gam[1])]
    ic = cases.ic_bump_advection
    obc = [orc.OracleBC(orc.BC_DIRICHLET, fx5["bidx"][0], fx5["bnrm"][0], value_fn=lambda x, t: ic(x, t))]
    obc += [orc.OracleBC(orc.BC_DO_NOTHING, fx5["bidx"][g], fx5["bnrm"][g]) for g in (1, 3, 2, 4)]
    return orc.OracleProblem(fx5["points"], 1, orc.EQ_ADVECTION2D, [1.0, 0.5], D[0], D[1], obc, srcs), gam


LIMITER = dict(thresholds=(0.05, 0.02), variables=(0, 1))      # (density, pressure)
# alpha = 0.01 dx_avg^2: on the boundary-imposed state (slip walls put a jump into the velocity field) larger alpha makes the
# reference's CG on its non-symmetric composite operator diverge (|r| grows 1e9-fold in 20 iterations), which is faithfully
# reproducible but amplifies last-bit differences; here it converges below sqrt(eps)|b| in ~8 iterations (early exit covered)
IGR_ALPHA_SCALE, IGR_MAXITER = 0.01, 20


def main():
    fx = cases.fixture_setup(p=3, N=3)
    pts = fx["points"]
    nb, ops = golden_ops()
    out = {}
    out["limiter_u"] = orc.limiter_zhang_shu(limiter_state(pts), nb, LIMITER["thresholds"], LIMITER["variables"], cases.GAMMA)
    alpha = IGR_ALPHA_SCALE * fx["dx_avg"] ** 2
    src = orc.source_igr(alpha=alpha, maxiter=IGR_MAXITER)
    P = orc.OracleProblem(pts, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1],
                          cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, cases.ic_smooth_euler), [src])
    u = cases.ic_smooth_euler(pts, 0.0)          # consistent with the Dirichlet data: no boundary jump feeding the CG
    out["igr_rhs_du"] = P.rhs(u, 0.0)
    out["igr_sigma"] = src.arrays["sigma"].copy()
    out["igr_iters"] = src.arrays["iters"]
    out["igr_alpha"] = alpha
    # ---- BASELINE configs[0]: linear advection a = (1, 0.5), PHS r^5 / degree 3, Flyer (k = 2) + Tominec hyperviscosity, SSPRK33,
    #      100 steps of dt = 0.1 dx_min.  The r^5 weight tables are stored (neighbour-table order,
    #      float64) so that the vectors do not depend on the LAPACK build that regenerates them.
    fx5 = cases.fixture_setup(p=5, N=3)
    D5 = orc.compute_flux_operator(pts, nb, 5, 3)
    H4 = orc.compute_flux_operator(pts, nb, 5, 3, 4)
    L2 = orc.compute_flux_operator(pts, nb, 5, 3, 2)
    import make_golden as mg0

    out["wx5"], out["wy5"] = mg0.rows_from_csc(D5[0], nb), mg0.rows_from_csc(D5[1], nb)
    out["h4"] = mg0.rows_from_csc(H4[0], nb) + mg0.rows_from_csc(H4[1], nb)       # Flyer: H = d4/dx4 + d4/dy4, same sparsity as D
    out["lap"] = mg0.rows_from_csc(L2[0], nb) + mg0.rows_from_csc(L2[1], nb)      # Tominec: H = lap' * lap
    P, gam = advection_problem(fx5, nb, out)
    out["adv_gamma_flyer"], out["adv_gamma_tominec"] = gam
    u0 = cases.ic_bump_advection(pts, 0.0)
    u = u0.copy()
    out["adv_rhs_du"] = P.rhs(u, 0.0)
    out["adv_rhs_u"] = u
    dt = 0.1 * fx5["dx_min"]
    out["adv_dt"] = dt
    out["adv_steps100_u"] = P.solve_ssprk33(u0, 0.0, dt, 100)[0]
    np.savez_compressed(os.path.join(HERE, "fixture_golden_f.npz"), **out)
    print("wrote fixture_golden_f.npz", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
