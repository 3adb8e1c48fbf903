"""Time-dependent Dirichlet data inside the device-resident steppers (the reference's BC closures
receive the STAGE time, rbfsolver.jl:311-316; one SSPRK step evaluates rhs! at t + dt and t + dt/2).  The host mirror hands
the tables of both stage times to the library before every step (mft_set_stage_boundary_values); the step makes each the
current table in front of the rhs! evaluated at that time (device-to-device copies inside the captured graph).
Against the oracle, whose Dirichlet closure is called with the stage time.  (Written without GPU access: first hardware run
is the round-end pass; the host logic was dry-run against a stand-in library.)"""
import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu


def _ic_t(x, t, equations=None):
    """inflow data that oscillates on the time scale of a few steps"""
    return cases.ic_smooth_euler(x, 0.0) * (1.0 + 0.05 * np.sin(900.0 * t))


def _build(m, fx, sources, osources):
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=fx["nv"])
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA())
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    bcs = cases.DIVERGENCE_TEST_BCS
    kinds = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(_ic_t, time_dependent=True),
                 slip=lambda: m.boundary_condition_slip_wall, nothing=lambda: m.BoundaryConditionDoNothing())
    semi = m.SemidiscretizationHyperbolic(domain, eq, _ic_t, solver, boundary_conditions={k: kinds[v]() for k, v in bcs.items()},
                                          source_terms=m.SourceTerms(**sources(m, solver, eq, domain)), operators=ops)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, bcs, _ic_t),
                          osources(fx))
    frozen = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1],
                               cases.oracle_bcs(fx, bcs, lambda x, t: _ic_t(x, 0.0)), osources(fx))
    return semi, P, frozen


def test_ssprk33_resident_loop_sees_the_stage_times():
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    semi, P, frozen = _build(m, fx, lambda m, s, e, d: dict(rv=m.SourceResidualViscosityTominec(s, e, d, polydeg=3)),
                             lambda fx: [orc.source_residual(fx["dx_avg"], polydeg=3)])
    dt, nsteps = 0.1 * fx["dx_min"] / 3.0, 8          # first use of the graph key runs eagerly, then 7 replays
    ode = m.semidiscretize(semi, (0.0, nsteps * dt))
    sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=m.HistoryCallback(approx_order=3), nsteps=nsteps)
    ur, _ = P.solve_ssprk33(ode.u0, 0.0, dt, nsteps, approx_order=3)
    assert cases.relerr(sol.u, ur) <= 1e-9
    uf, _ = frozen.solve_ssprk33(ode.u0, 0.0, dt, nsteps, approx_order=3)
    assert cases.relerr(uf, ur) > 1e-5                # the test would notice tables frozen at t = 0
    inlet = fx["bidx"][cases.BOUNDARY_NAMES["inlet"] - 1]
    np.testing.assert_array_equal(sol.u[:, inlet], _ic_t(fx["points"][inlet], nsteps * dt))
    semi.close()


def test_ssprk43_adaptive_with_rejections_sees_the_stage_times():
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    semi, P, frozen = _build(m, fx, lambda m, s, e, d: dict(uw=m.SourceUpwindViscosityTominec(s, e, d)),
                             lambda fx: [orc.source_upwind(fx["dx_avg"])])
    u0 = _ic_t(fx["points"], 0.0)
    u_ref, t_ref, log_ref = orc.solve_ssprk43(P, u0, 0.0, 0.004, 1e-3, abstol=1e-6, reltol=1e-6)
    sol = m.solve(m.semidiscretize(semi, (0.0, 0.004)), m.SSPRK43(), dt=1e-3, abstol=1e-6, reltol=1e-6)
    assert [a[3] for a in sol.log] == [a[3] for a in log_ref], "accept/reject sequence differs"
    assert abs(sol.t - t_ref) < 1e-15 and cases.relerr(sol.u, u_ref) <= 1e-9
    uf, _, _ = orc.solve_ssprk43(frozen, u0, 0.0, 0.004, 1e-3, abstol=1e-6, reltol=1e-6)
    assert cases.relerr(uf, u_ref) > 1e-5
    semi.close()
