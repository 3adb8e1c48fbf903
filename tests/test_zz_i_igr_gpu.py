"""SourceIGR on the device (src/sources/IGR.jl) against the oracle, through the C ABI.
(Written after this round's GPU budget was spent: first hardware run is the round-end test pass; the kernel bodies, the
launch sequence and the reduction tree are covered on the CPU tier by tests/test_igr.py via tests/emu.)"""
import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu


def _problem(alpha, maxiter=20, reorder="hilbert", extra_sources=(), bcs=None, diagnostics=False):
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=fx["nv"])
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(reorder=reorder, diagnostics=diagnostics))
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = cases.ic_smooth_euler
    bcs = cases.DIVERGENCE_TEST_BCS if bcs is None else bcs
    kinds = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic), slip=lambda: m.boundary_condition_slip_wall,
                 nothing=lambda: m.BoundaryConditionDoNothing())
    srcs = dict(extra_sources)
    srcs["igr"] = m.SourceIGR(solver, eq, domain, alpha=alpha, linear_solver=m.cg_, maxiter=maxiter)
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions={k: kinds[v]() for k, v in bcs.items()},
                                          source_terms=m.SourceTerms(**srcs), operators=ops)
    return m, fx, ops, semi, bcs, ic


@pytest.mark.parametrize("alpha_scale,maxiter,reorder", [(20.0, 20, "hilbert"), (0.01, 20, None), (5.0, 2, "hilbert")])
def test_igr_source_functor_matches_oracle(alpha_scale, maxiter, reorder):
    fx0 = cases.fixture_setup(p=3, N=3)
    alpha = alpha_scale * fx0["dx_avg"] ** 2
    m, fx, ops, semi, bcs, ic = _problem(alpha, maxiter, reorder)
    u = ic(fx["points"], 0.0)
    du0 = 0.01 * np.sin(3 * u)
    du = du0.copy()
    semi.source_terms.igr(du, u, 0.0)
    src = orc.source_igr(alpha=alpha, maxiter=maxiter)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], [], [src])
    du_ref = du0.copy()
    P.apply_source(0, u, du_ref)
    it, res, res0 = semi.source_terms.igr.cache.igr_status
    assert it == src.arrays["iters"]
    sigma = semi.source_terms.igr.cache.sigma
    assert np.abs(sigma - src.arrays["sigma"]).max() <= 1e-9 * np.abs(src.arrays["sigma"]).max()
    assert abs(res - src.arrays["res"]) <= 1e-8 * res0
    assert np.array_equal(du[0], du0[0]) and np.array_equal(du[3], du0[3])
    assert cases.relerr(du, du_ref) <= 1e-9
    semi.close()


def test_igr_inside_rhs_and_time_loop():
    """rhs! with SourceIGR (1e-9 against the oracle) and a few graph-replayed SSPRK33 steps (1e-8).  alpha = 0.01 dx^2: on
    the boundary-imposed state larger alpha makes the reference's CG diverge (see tests/golden/make_golden_f.py)"""
    fx0 = cases.fixture_setup(p=3, N=3)
    alpha = 0.01 * fx0["dx_avg"] ** 2
    m, fx, ops, semi, bcs, ic = _problem(alpha)
    ode = m.semidiscretize(semi, (0.0, 1.0))
    u = ode.u0.copy()
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, bcs, ic),
                          [orc.source_igr(alpha=alpha)])
    u_ref = ode.u0.copy()
    du_ref = P.rhs(u_ref, 0.0)
    assert np.array_equal(u, u_ref)
    assert cases.relerr(du, du_ref) <= 1e-9
    dt, nsteps = 0.1 * fx["dx_min"] / 3.0, 4
    sol = m.solve(ode, m.SSPRK33(), dt=dt, nsteps=nsteps)
    ur, _ = P.solve_ssprk33(ode.u0, 0.0, dt, nsteps)
    assert cases.relerr(sol.u, ur) <= 1e-8
    semi.close()


def test_igr_after_upwind_viscosity_and_vtk_field(tmp_path):
    """two sources in NamedTuple order (upwind viscosity fused into the flux sweep, then IGR); sigma reaches the VTK file"""
    fx0 = cases.fixture_setup(p=3, N=3)
    alpha = 0.01 * fx0["dx_avg"] ** 2
    import mft_b200 as m0

    basis = m0.PointCloudBasis(m0.Point2D(), 3, approximation_type=m0.RBF(m0.PolyharmonicSpline(3)), nv=fx0["nv"])
    solver0 = m0.PointCloudSolver(basis, engine=m0.RBFFDEngineCUDA())
    domain0 = m0.PointCloudDomain(solver0, cases.FIXTURE, cases.BOUNDARY_NAMES)
    uw = m0.SourceUpwindViscosityTominec(solver0, m0.CompressibleEulerEquations2D(cases.GAMMA), domain0)
    m, fx, ops, semi, bcs, ic = _problem(alpha, extra_sources=dict(uw=uw), diagnostics=True)   # eps fields for the VTK file
    ode = m.semidiscretize(semi, (0.0, 1.0))
    u = ode.u0.copy()
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, bcs, ic),
                          [orc.source_upwind(fx["dx_avg"]), orc.source_igr(alpha=alpha)])
    du_ref = P.rhs(ode.u0.copy(), 0.0)
    assert cases.relerr(du, du_ref) <= 1e-9
    f = m.trixi2vtk(u, semi, 0.0, iter=0, output_directory=str(tmp_path), prefix="igr")
    _, pdata, _, _ = m.vtk.read_vtu(f)
    assert "sigma" in pdata and np.abs(pdata["sigma"] - P.sources[1].arrays["sigma"]).max() <= 1e-9 * np.abs(pdata["sigma"]).max()
    semi.close()
