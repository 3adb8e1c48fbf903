"""The reference's driver script (rbfsolver_test.jl: PointCloudBasis -> PointCloudSolver -> PointCloudDomain from Medusa
files -> sources -> SemidiscretizationHyperbolic -> semidiscretize -> solve(SSPRK43, callbacks)) with the B200 engine.
Same names, same argument meaning; the only difference is `engine=RBFFDEngineCUDA()` in the solver.

    python examples/cylinder_euler_rv.py [--tend 0.02] [--out out_cyl]

Needs a CUDA device (the package has no CPU engine).  Cloud: the reference's own test fixture (test/data/cyl_0_05, 2154 points,
a channel with a cylinder), shipped under tests/golden/ as the golden-vector cloud."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mft_b200 as m  # noqa: E402


def initial_condition_flow(x, t, equations):
    """uniform inflow (rho, v1, v2, p) = (1.4, 0.8, 0, 1.0); the reference script uses initial_condition_constant"""
    n = x.shape[0]
    rho, v1, v2, p = 1.4, 0.8, 0.0, 1.0
    return np.stack([np.full(n, rho), np.full(n, rho * v1), np.full(n, rho * v2),
                     np.full(n, p / (equations.gamma - 1.0) + 0.5 * rho * (v1 * v1 + v2 * v2))])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tend", type=float, default=0.02)
    ap.add_argument("--out", default="out_cyl")
    args = ap.parse_args()

    approximation_order = 3
    basis = m.PointCloudBasis(m.Point2D(), approximation_order, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(diagnostics=True))      # <- the one changed line

    casename = os.path.join(ROOT, "tests", "golden", "cyl_0_05", "cyl_0_05")
    boundary_names = dict(inlet=1, outlet=2, bottom=3, top=4, cyl=5)
    domain = m.PointCloudDomain(solver, casename, boundary_names)

    equations = m.CompressibleEulerEquations2D(1.4)
    boundary_conditions = dict(inlet=m.BoundaryConditionDirichlet(initial_condition_flow),
                               outlet=m.BoundaryConditionDoNothing(),
                               bottom=m.boundary_condition_slip_wall, top=m.boundary_condition_slip_wall,
                               cyl=m.boundary_condition_slip_wall)

    history_callback = m.HistoryCallback(approx_order=approximation_order)
    source_rv = m.SourceResidualViscosityTominec(solver, equations, domain, c_rv=1.0, c_uw=1.0, polydeg=approximation_order)
    sources = m.SourceTerms(rv=source_rv)
    semi = m.SemidiscretizationHyperbolic(domain, equations, initial_condition_flow, solver,
                                          boundary_conditions=boundary_conditions, source_terms=sources)
    ode = m.semidiscretize(semi, (0.0, args.tend))

    save_callback = m.SolutionSavingCallback(dt=args.tend / 4, output_directory=args.out, prefix="cyl")
    time_int_tol = 1e-6
    sol = m.solve(ode, m.SSPRK43(), dt=1e-4, abstol=time_int_tol, reltol=time_int_tol,
                  callback=[history_callback, save_callback])

    accepted = sum(1 for entry in sol.log if entry[3])
    rho = sol.u[0]
    print(f"t = {sol.t:.5f}: {accepted} accepted / {len(sol.log) - accepted} rejected steps, {sol.nrhs} rhs! evaluations, "
          f"density in [{rho.min():.4f}, {rho.max():.4f}], non-finite entries: {semi.count_nonfinite()}")
    print("snapshots:", ", ".join(os.path.basename(f) for f in save_callback.files))
    semi.close()


if __name__ == "__main__":
    main()
