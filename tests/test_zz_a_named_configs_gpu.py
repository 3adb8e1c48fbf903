"""BASELINE.json configs[2] and configs[3] at parity-test size (the full sizes are bench / scaling runs): same physics, same
boundary-condition mix, same sources, against the oracle -- 1e-12 per rhs!, 1e-9 after N steps (north_star tolerances).
  configs[2]: Euler isentropic vortex + upwind (first-order) viscosity, jittered cloud, Dirichlet on all sides
  configs[3]: Euler Sod shock tube + residual viscosity + history callback, slip walls top/bottom, Dirichlet left/right
              (discontinuous data: both branches of update_visc!'s min(eps_rv, eps_uw) are taken)
Uses only kernels that have run on hardware; the test code itself is new (first hardware run: round-end pass)."""
import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu

NAMES = dict(left=1, right=2, bottom=3, top=4)


def _build(m, cl, ic, bc_kinds, sources, reorder="hilbert"):
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(diagnostics=True, reorder=reorder))
    domain = m.PointCloudDomain(solver, cl, NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    mk = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic), slip=lambda: m.boundary_condition_slip_wall)
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions={k: mk[v]() for k, v in bc_kinds.items()},
                                          source_terms=m.SourceTerms(**sources(m, solver, eq, domain)))
    pd = domain.pd
    okind = dict(dirichlet=orc.BC_DIRICHLET, slip=orc.BC_SLIP_WALL)
    obc = [orc.OracleBC(okind[v], domain.boundary_tags[k].idx, domain.boundary_tags[k].normals,
                        value_fn=(lambda x, t: ic(x, t)) if v == "dirichlet" else None) for k, v in bc_kinds.items()]
    return semi, domain, pd, obc


def test_config2_vortex_upwind_viscosity():
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(96, 80, 10.0, 10.0 * 80 / 96, seed=1)
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, cases.GAMMA, center=(5.0, 4.0))   # noqa: E731
    semi, domain, pd, obc = _build(m, cl, ic, dict(left="dirichlet", right="dirichlet", bottom="dirichlet", top="dirichlet"),
                                   lambda m, s, e, d: dict(uw=m.SourceUpwindViscosityTominec(s, e, d, c_uw=1.0)))
    ops = semi.cache.rbf_differentiation_matrices
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], obc, [orc.source_upwind(pd.dx_avg)])
    ode = m.semidiscretize(semi, (0.0, 1.0))
    u = ode.u0 * (1.0 + 0.01 * np.sin(pd.points[:, 0]))          # off the Dirichlet data: the BC passes do real work
    u_ref = u.copy()
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    du_ref = P.rhs(u_ref, 0.0)
    assert np.array_equal(u, u_ref) and cases.relerr(du, du_ref) <= 1e-12
    dt, nsteps = 0.1 * pd.dx_min / 8.0, 20
    sol = m.solve(ode, m.SSPRK33(), dt=dt, nsteps=nsteps)
    ur, _ = P.solve_ssprk33(ode.u0, 0.0, dt, nsteps)
    assert cases.relerr(sol.u, ur) <= 1e-9
    semi.close()


@pytest.mark.parametrize("reorder", ["hilbert", None])
def test_config3_sod_residual_viscosity(reorder):
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(96, 48, 2.0, 1.0, seed=2)
    ic = lambda x, t, e=None: m.cloud.sod(x, cases.GAMMA, x_mid=1.0)   # noqa: E731
    semi, domain, pd, obc = _build(m, cl, ic, dict(left="dirichlet", right="dirichlet", bottom="slip", top="slip"),
                                   lambda m, s, e, d: dict(rv=m.SourceResidualViscosityTominec(s, e, d, c_rv=1.0, c_uw=1.0, polydeg=3)),
                                   reorder=reorder)
    ops = semi.cache.rbf_differentiation_matrices
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], obc, [orc.source_residual(pd.dx_avg, polydeg=3)])
    ode = m.semidiscretize(semi, (0.0, 1.0))
    dt, nsteps = 0.1 * pd.dx_min / 3.0, 20
    sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=m.HistoryCallback(approx_order=3), nsteps=nsteps)
    ur, _ = P.solve_ssprk33(ode.u0, 0.0, dt, nsteps, approx_order=3)
    assert np.isfinite(ur).all() and cases.relerr(sol.u, ur) <= 1e-9
    # the limiter took both branches, identically on both sides (eps_c: 0 = residual viscosity won, 1 = upwind cap)
    flag = semi.source_terms.rv.cache.eps_c
    ref_flag = P.sources[0].arrays["eps_c"]
    assert (flag != ref_flag).sum() <= 2 and 0 < (ref_flag == 1).sum() < len(ref_flag)   # (a near-tie of the two caps may flip)
    eps = semi.source_terms.rv.cache.eps
    assert np.abs(eps - P.sources[0].arrays["eps"]).max() <= 1e-9 * np.abs(eps).max()
    # and one more rhs! on the shocked state: the same state AND the same time-history residual on both sides (the oracle takes
    # the device's approx_du, as tests/test_gpu_scale.py does), so the north_star tolerance of a single rhs! applies: 1e-12
    P.sources[0].arrays["approx_du"][:] = semi.source_terms.rv.cache.approx_du
    u = ur.copy()
    u_ref = ur.copy()
    du = np.empty_like(u)
    m.rhs_(du, u, semi, nsteps * dt)
    assert cases.relerr(du, P.rhs(u_ref, nsteps * dt)) <= 1e-12
    semi.close()
