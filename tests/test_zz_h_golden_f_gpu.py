"""The CUDA path reproduces the committed golden vectors of the "next" rows WITHOUT the oracle in the loop
(tests/golden/fixture_golden_f.npz, made by tests/golden/make_golden_f.py; kNN table / weights of fixture_golden.npz).
First hardware run is the round-end test pass (see tests/test_zz_g_setup_gpu.py)."""
import os
import sys

import numpy as np
import pytest

import cases

sys.path.insert(0, cases.GOLDEN)
import make_golden_f as mg  # noqa: E402

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(cases.GOLDEN, "fixture_golden.npz"))
GF = np.load(os.path.join(cases.GOLDEN, "fixture_golden_f.npz"))


def _semi(m, sources=None, **engine):
    _, ops = mg.golden_ops()
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(**engine))
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = cases.ic_smooth_euler
    kinds = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic), slip=lambda: m.boundary_condition_slip_wall,
                 nothing=lambda: m.BoundaryConditionDoNothing())
    srcs = m.SourceTerms(**(sources(m, solver, eq, domain) if sources else {}))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, source_terms=srcs, operators=ops,
                                          boundary_conditions={k: kinds[v]() for k, v in cases.DIVERGENCE_TEST_BCS.items()})
    return semi, domain


def test_device_setup_reproduces_golden_tables():
    import mft_b200 as m

    semi, domain = _semi(m, setup="device")
    assert np.array_equal(domain.pd.neighbors, G["neighbors"])
    assert domain.pd.dx_min == float(G["dx_min"]) and domain.pd.dx_avg == float(G["dx_avg"])
    wx, wy = m.setup_ops.rbf_fd_weights_device(domain.pd.points, domain.pd.neighbors, 3, 3)
    assert np.abs(wx - G["wx"]).max() <= 1e-8 * np.abs(G["wx"]).max() and np.abs(wy - G["wy"]).max() <= 1e-8 * np.abs(G["wy"]).max()
    semi.close()


def test_device_limiter_reproduces_golden():
    import mft_b200 as m

    semi, domain = _semi(m)
    lim = m.PositivityPreservingLimiterZhangShu(thresholds=mg.LIMITER["thresholds"], variables=(m.density, m.pressure))
    u = lim(mg.limiter_state(domain.pd.points), semi)
    assert np.array_equal(u, GF["limiter_u"])
    semi.close()


def test_device_igr_rhs_reproduces_golden():
    import mft_b200 as m

    semi, domain = _semi(m, sources=lambda m, solver, eq, domain: dict(
        igr=m.SourceIGR(solver, eq, domain, alpha=float(GF["igr_alpha"]), linear_solver=m.cg_, maxiter=mg.IGR_MAXITER)))
    u = cases.ic_smooth_euler(domain.pd.points, 0.0)
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    assert cases.relerr(du, GF["igr_rhs_du"]) <= 1e-9
    sigma = semi.source_terms.igr.cache.sigma
    assert np.abs(sigma - GF["igr_sigma"]).max() <= 1e-9 * np.abs(GF["igr_sigma"]).max()
    assert semi.source_terms.igr.cache.igr_status[0] == int(GF["igr_iters"])
    semi.close()


def test_device_advection_config_reproduces_golden():
    """BASELINE configs[0] on the CUDA path from the stored r^5 tables: rhs! bit-identical, 100 SSPRK33 steps to 1e-9"""
    import mft_b200 as m

    nb = G["neighbors"].astype(np.int64)
    D, Hf, Ht = mg.advection_matrices(nb, GF)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(5)))
    solver = m.PointCloudSolver(basis)
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.LinearScalarAdvectionEquation2D(1.0, 0.5)
    ic = cases.ic_bump_advection
    bc = dict(inlet=m.BoundaryConditionDirichlet(ic), outlet=m.BoundaryConditionDoNothing(),
              top=m.BoundaryConditionDoNothing(), bottom=m.BoundaryConditionDoNothing(), cyl=m.BoundaryConditionDoNothing())
    hv, hv2 = object.__new__(m.SourceHyperviscosityFlyer), object.__new__(m.SourceHyperviscosityTominec)
    hv.hv_differentiation_matrix, hv.gamma, hv.c = Hf, float(GF["adv_gamma_flyer"]), 1.0      # matrices from the golden tables,
    hv2.hv_differentiation_matrix, hv2.gamma, hv2.c = Ht, float(GF["adv_gamma_tominec"]), 1.0  # not regenerated here
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc,
                                          source_terms=m.SourceTerms(hv=hv, hv2=hv2), operators=D)
    u = ic(domain.pd.points, 0.0)
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    assert np.array_equal(u, GF["adv_rhs_u"]) and np.array_equal(du, GF["adv_rhs_du"])
    sol = m.solve(m.semidiscretize(semi, (0.0, 100 * float(GF["adv_dt"]))), m.SSPRK33(), dt=float(GF["adv_dt"]), nsteps=100)
    assert cases.relerr(sol.u, GF["adv_steps100_u"]) <= 1e-9
    semi.close()
