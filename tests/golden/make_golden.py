"""Generates tests/golden/fixture_golden.npz from the CPU oracle (oracle/).

The reference itself cannot be executed in this environment (pure Julia, no Julia toolchain, un-vendored
dependency -- see DESIGN.md), and it stores no golden vectors of its own.  These vectors are therefore the
oracle's outputs on the reference's fixture cloud test/data/cyl_0_05 (copied to tests/golden/cyl_0_05/), the oracle
being pinned by the reference's test identities (tests/test_oracle_identities.py).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cases  # noqa: E402
from cases import orc  # noqa: E402


def rows_from_csc(A, nb):
    """weights in neighbour-table order: w[i, c] = A[i, nb[i, c]]"""
    R = A.tocsr()
    out = np.empty(nb.shape)
    for i in range(nb.shape[0]):
        cols = R.indices[R.indptr[i]:R.indptr[i + 1]]
        vals = R.data[R.indptr[i]:R.indptr[i + 1]]
        lut = dict(zip(cols.tolist(), vals.tolist()))
        out[i] = [lut[j] for j in nb[i]]
    return out


def main():
    fx = cases.fixture_setup(p=3, N=3)
    pts, nb = fx["points"], fx["nb"]
    Dx, Dy = orc.compute_flux_operator(pts, nb, 3, 3)
    out = dict(neighbors=nb.astype(np.int32), dx_min=fx["dx_min"], dx_avg=fx["dx_avg"], wx=rows_from_csc(Dx, nb),
               wy=rows_from_csc(Dy, nb))

    def problem(sources=(), ic=cases.ic_smooth_euler):
        return orc.OracleProblem(pts, 4, orc.EQ_EULER2D, [cases.GAMMA], Dx, Dy,
                                 cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, ic), list(sources))

    # test/divergence_test.jl: calc_fluxes! on the gradient initial condition
    u0 = cases.ic_gradient(pts, 0.0)
    du = np.zeros_like(u0)
    problem(ic=cases.ic_gradient).calc_fluxes(u0, du)
    out["calc_fluxes_gradient_du"] = du
    # test/upwind_viscosity_test.jl: the source alone
    src = orc.source_upwind(fx["dx_avg"])
    du = np.zeros_like(u0)
    problem([src], ic=cases.ic_gradient).apply_source(0, u0, du)
    out["upwind_source_gradient_du"] = du
    out["upwind_source_gradient_eps"] = src.arrays["eps"].copy()
    # whole rhs! for the three source sets, state off the Dirichlet data
    for name, mk in (("none", lambda: []), ("upwind", lambda: [orc.source_upwind(fx["dx_avg"])]),
                     ("residual", lambda: [orc.source_residual(fx["dx_avg"], polydeg=3)])):
        u = cases.ic_smooth_euler(pts, 0.0) * 1.01
        out[f"rhs_{name}_du"] = problem(mk()).rhs(u, 0.0)
        out[f"rhs_{name}_u"] = u
    # Euler + residual viscosity + history callback, 30 SSPRK33 steps
    dt = 0.1 * fx["dx_min"] / 3.0
    P = problem([orc.source_residual(fx["dx_avg"], polydeg=3)])
    u_end, t_end = P.solve_ssprk33(cases.ic_smooth_euler(pts, 0.0), 0.0, dt, 30, approx_order=3)
    out["steps30_dt"] = dt
    out["steps30_u"] = u_end
    out["steps30_approx_du"] = P.sources[0].arrays["approx_du"].copy()
    np.savez_compressed(os.path.join(HERE, "fixture_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "fixture_golden.npz"), {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
