"""SolutionSavingCallback driven by the device-resident integrators: the files hold the
state / viscosity fields of the step they were written at.  (Written after this round's GPU budget was spent: first
hardware run is the round-end test pass.)"""
import os

import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu


def test_saving_callback_snapshots_match_the_oracle_trajectory(tmp_path):
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=fx["nv"])
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(diagnostics=True))
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = cases.ic_smooth_euler
    bc = dict(inlet=m.BoundaryConditionDirichlet(ic), outlet=m.BoundaryConditionDoNothing(), top=m.boundary_condition_slip_wall,
              bottom=m.boundary_condition_slip_wall, cyl=m.boundary_condition_slip_wall)
    srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs, operators=ops)
    dt, nsteps = 0.1 * fx["dx_min"] / 3.0, 6
    ode = m.semidiscretize(semi, (0.0, nsteps * dt))
    saver = m.SolutionSavingCallback(interval=3, output_directory=str(tmp_path), prefix="t")
    sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=[m.HistoryCallback(approx_order=3), saver], nsteps=nsteps)
    assert [os.path.basename(f) for f in saver.files] == [f"t_CompressibleEulerEquations2D_1_{i}.vtu" for i in (0, 3, 6)]
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1],
                          cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, ic), [orc.source_residual(fx["dx_avg"], polydeg=3)])
    for f, k in zip(saver.files, (0, 3, 6)):
        _, pdata, fdata, _ = m.vtk.read_vtu(f)
        u_ref = ode.u0 if k == 0 else P.solve_ssprk33(ode.u0, 0.0, dt, k, approx_order=3)[0]
        assert abs(fdata["time"][0] - k * dt) < 1e-14
        assert cases.relerr(np.stack([pdata["density"], pdata["momentum"][:, 0], pdata["momentum"][:, 1], pdata["density_energy"]]), u_ref) <= 1e-9
        assert {"eps", "eps_scalar", "eps_uw", "eps_rv", "approx_du", "residual", "pressure", "velocity", "index"} <= set(pdata)
    assert cases.relerr(sol.u, P.solve_ssprk33(ode.u0, 0.0, dt, nsteps, approx_order=3)[0]) <= 1e-9
    semi.close()
