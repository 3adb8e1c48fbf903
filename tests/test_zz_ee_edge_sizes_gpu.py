"""Edge sizes through the C ABI: clouds smaller than one tile / one slice, sizes straddling the 32-row slice and 128-row
tile boundaries, wide and narrow stencils, an empty boundary group -- rhs! (flux divergence + upwind / residual viscosity)
against the oracle at 1e-12, for both kernel families (union tiles and thread-per-row).  Known kernels, new sizes (first
hardware run of this file: round-end pass)."""
import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu

NAMES = dict(left=1, right=2, bottom=3, top=4)


def _case(m, nx, ny, nv, tile, source, empty_group=False, refine_order=False, tile_rows=11):
    cl = m.cloud.jittered_lattice(nx, ny, 2.0, 2.0 * ny / nx, seed=5)
    if empty_group:      # BoundaryData with no points (a group that exists in the file but is empty on this cloud)
        cl.boundary_idxs[3] = np.zeros(0, dtype=np.int64)
        cl.boundary_normals[3] = np.zeros((0, 2))
    deg = 3 if nv >= 20 else 2
    basis = m.PointCloudBasis(m.Point2D(), deg, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=nv)
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(tile=tile, refine_order=refine_order, tile_rows=tile_rows))
    domain = m.PointCloudDomain(solver, cl, NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = lambda x, t, e=None: cases.ic_smooth_euler(x, t)   # noqa: E731
    bc = dict(left=m.BoundaryConditionDirichlet(ic), right=m.BoundaryConditionDoNothing(), bottom=m.boundary_condition_slip_wall,
              top=m.boundary_condition_slip_wall)
    pd = domain.pd
    if source == "upwind":
        srcs, osrc = m.SourceTerms(s=m.SourceUpwindViscosityTominec(solver, eq, domain)), [orc.source_upwind(pd.dx_avg)]
    elif source == "residual":
        srcs = m.SourceTerms(s=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3))
        osrc = [orc.source_residual(pd.dx_avg, polydeg=3)]
    else:
        srcs, osrc = m.SourceTerms(), []
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs)
    ops = semi.cache.rbf_differentiation_matrices
    kinds = dict(left=orc.BC_DIRICHLET, right=orc.BC_DO_NOTHING, bottom=orc.BC_SLIP_WALL, top=orc.BC_SLIP_WALL)
    obc = [orc.OracleBC(kinds[k], domain.boundary_tags[k].idx, domain.boundary_tags[k].normals,
                        value_fn=(lambda x, t: ic(x, t)) if kinds[k] == orc.BC_DIRICHLET else None) for k in bc]
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], obc, osrc)
    u = ic(pd.points, 0.0) * (1.0 + 0.02 * np.cos(3 * pd.points[:, 1]))
    u_ref = u.copy()
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    du_ref = P.rhs(u_ref, 0.0)
    assert np.array_equal(u, u_ref), (nx, ny, nv, tile, source)
    assert cases.relerr(du, du_ref) <= 1e-12, (nx, ny, nv, tile, source, cases.relerr(du, du_ref))
    # a couple of graph-replayed steps as well (stage kernels, FSAL bookkeeping at odd sizes)
    dt = 0.05 * pd.dx_min
    ode = m.semidiscretize(semi, (0.0, 3 * dt))
    hist = m.HistoryCallback(approx_order=3) if source == "residual" else None
    sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=hist, nsteps=3)
    ur, _ = P.solve_ssprk33(ode.u0, 0.0, dt, 3, approx_order=3 if source == "residual" else None)
    assert cases.relerr(sol.u, ur) <= 1e-10, (nx, ny, nv, tile, source)
    semi.close()


@pytest.mark.parametrize("tile", [31, 15, 0])
@pytest.mark.parametrize("nx,ny", [(4, 3), (5, 5), (8, 8), (10, 10), (11, 11), (15, 15), (16, 15)])
def test_odd_cloud_sizes(nx, ny, tile):
    """26, 45, 96, 140, 165, 285, 302 points with the default 20-wide stencil"""
    import mft_b200 as m

    _case(m, nx, ny, 20, tile, "upwind")


@pytest.mark.parametrize("tile", [31, 15, 0])
@pytest.mark.parametrize("nv", [13, 15, 30, 42])
def test_stencil_widths(nv, tile):
    import mft_b200 as m

    _case(m, 14, 12, nv, tile, "residual")


def test_empty_boundary_group_and_no_sources():
    import mft_b200 as m

    _case(m, 9, 7, 20, 15, None, empty_group=True)
    _case(m, 9, 7, 20, 0, "upwind", empty_group=True)
