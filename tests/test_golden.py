"""Committed golden vectors (tests/golden/fixture_golden.npz, made by tests/golden/make_golden.py).
CPU: the oracle reproduces them.  GPU: the CUDA path reproduces them WITHOUT the oracle in the loop."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import cases

G = np.load(os.path.join(cases.GOLDEN, "fixture_golden.npz"))


def _ops_from_golden():
    nb = G["neighbors"].astype(np.int64)
    n, k = nb.shape
    rows = np.repeat(np.arange(n), k)
    out = []
    for w in (G["wx"], G["wy"]):
        A = sp.coo_matrix((w.reshape(-1), (rows, nb.reshape(-1))), shape=(n, n)).tocsc()
        A.sort_indices()
        out.append(A)
    return nb, out


def test_oracle_reproduces_golden():
    from cases import orc

    fx = cases.fixture_setup(p=3, N=3)
    assert np.array_equal(fx["nb"], G["neighbors"]) and fx["dx_min"] == float(G["dx_min"]) and fx["dx_avg"] == float(G["dx_avg"])
    nb, ops = _ops_from_golden()
    mk = lambda sources=(), ic=cases.ic_smooth_euler: orc.OracleProblem(
        fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, ic), list(sources))
    u0 = cases.ic_gradient(fx["points"], 0.0)
    du = np.zeros_like(u0)
    mk(ic=cases.ic_gradient).calc_fluxes(u0, du)
    assert np.array_equal(du, G["calc_fluxes_gradient_du"])
    for name, srcs in (("none", []), ("upwind", [orc.source_upwind(fx["dx_avg"])]),
                       ("residual", [orc.source_residual(fx["dx_avg"], polydeg=3)])):
        u = cases.ic_smooth_euler(fx["points"], 0.0) * 1.01
        du = mk(srcs).rhs(u, 0.0)
        assert np.array_equal(u, G[f"rhs_{name}_u"])
        assert cases.relerr(du, G[f"rhs_{name}_du"]) < 1e-13
    # regenerated operators agree with the stored ones (LAPACK build differences stay below 1e-9 of the row scale)
    Dx, Dy = orc.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    assert np.abs(Dx.toarray() - ops[0].toarray()).max() < 1e-9 * np.abs(G["wx"]).max()


@pytest.mark.gpu
def test_cuda_path_reproduces_golden():
    import mft_b200 as m

    nb, ops = _ops_from_golden()
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(diagnostics=True))
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    # kNN indices bit-exact, spacing constants identical
    assert np.array_equal(domain.pd.neighbors, nb)
    assert domain.pd.dx_min == float(G["dx_min"]) and domain.pd.dx_avg == float(G["dx_avg"])
    # operator sparsity pattern bit-exact, weights to 1e-8 of the row scale (different dense solver)
    mine = m.setup_ops.compute_flux_operator(domain.pd.points, domain.pd.neighbors, 3, 3)
    for A, B in zip(mine, ops):
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        assert np.abs(A.data - B.data).max() < 1e-8 * np.abs(B.data).max()
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)

    def semi_for(srcs, ic):
        bc = dict(inlet=m.BoundaryConditionDirichlet(ic), outlet=m.BoundaryConditionDoNothing(),
                  top=m.boundary_condition_slip_wall, bottom=m.boundary_condition_slip_wall, cyl=m.boundary_condition_slip_wall)
        return m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=m.SourceTerms(**srcs),
                                              operators=ops)

    semi = semi_for({}, cases.ic_gradient)
    u0 = cases.ic_gradient(domain.pd.points, 0.0)
    du = np.zeros_like(u0)
    m.calc_fluxes_(du, u0, semi)
    assert np.array_equal(du, G["calc_fluxes_gradient_du"])          # test/divergence_test.jl, bit for bit
    semi.close()
    semi = semi_for(dict(rv=m.SourceUpwindViscosityTominec(solver, eq, domain)), cases.ic_gradient)
    du = np.zeros_like(u0)
    semi.source_terms.rv(du, u0, 0.0)
    assert np.array_equal(du, G["upwind_source_gradient_du"])        # test/upwind_viscosity_test.jl
    assert np.array_equal(semi.source_terms.rv.cache.eps, G["upwind_source_gradient_eps"])
    semi.close()
    for name, srcs in (("none", {}), ("upwind", dict(rv=m.SourceUpwindViscosityTominec(solver, eq, domain))),
                       ("residual", dict(rv=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3)))):
        semi = semi_for(srcs, cases.ic_smooth_euler)
        u = cases.ic_smooth_euler(domain.pd.points, 0.0) * 1.01
        du = np.empty_like(u)
        m.rhs_(du, u, semi, 0.0)
        assert np.array_equal(u, G[f"rhs_{name}_u"])
        assert cases.relerr(du, G[f"rhs_{name}_du"]) <= 1e-12
        if name != "residual":
            assert np.array_equal(du, G[f"rhs_{name}_du"])
        semi.close()
    semi = semi_for(dict(rv=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3)), cases.ic_smooth_euler)
    sol = m.solve(m.semidiscretize(semi, (0.0, 1.0)), m.SSPRK33(), dt=float(G["steps30_dt"]),
                  callback=m.HistoryCallback(approx_order=3), nsteps=30)
    assert cases.relerr(sol.u, G["steps30_u"]) <= 1e-9
    semi.close()
