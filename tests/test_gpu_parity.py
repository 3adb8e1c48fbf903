"""GPU parity: the CUDA path (libmft_b200.so through its C ABI, driven by the host mirror of the reference API)
against the CPU oracle on the same inputs.  Tolerances (BASELINE.json north_star): 1e-12 normwise per rhs!,
1e-9 after N steps; in exact-order mode the Dx/Dy sums are required to be BIT-IDENTICAL to the oracle."""
import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-12
STEP_TOL = 1e-9


def _mft():
    import mft_b200

    return mft_b200


@pytest.fixture(scope="module")
def fx():
    m = _mft()
    s = cases.fixture_setup(p=3, N=3)
    s["ops"] = m.setup_ops.compute_flux_operator(s["points"], s["nb"], 3, 3)
    return s


def _semi(fx, sources=None, bcs=cases.DIVERGENCE_TEST_BCS, ic=cases.ic_gradient, time_dependent=False, **engine_kw):
    """the product side, built exactly like test/divergence_test.jl builds the reference side"""
    m = _mft()
    basis = m.PointCloudBasis(m.Point2D(), fx["N"], approximation_type=m.RBF(m.PolyharmonicSpline(fx["p"])), nv=fx["nv"])
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(**engine_kw))
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    assert np.array_equal(domain.pd.neighbors, fx["nb"])      # kNN indices bit-exact vs the oracle's
    equations = m.CompressibleEulerEquations2D(cases.GAMMA)
    kinds = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic, time_dependent=time_dependent),
                 slip=lambda: m.boundary_condition_slip_wall, nothing=lambda: m.BoundaryConditionDoNothing())
    bc = {name: kinds[k]() for name, k in bcs.items()}
    srcs = m.SourceTerms(**(sources(m, solver, equations, domain) if sources else {}))
    semi = m.SemidiscretizationHyperbolic(domain, equations, ic, solver, boundary_conditions=bc, source_terms=srcs,
                                          operators=fx["ops"])
    return m, semi


def _oracle(fx, sources=(), bcs=cases.DIVERGENCE_TEST_BCS, ic=cases.ic_gradient):
    return orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], fx["ops"][0], fx["ops"][1],
                             cases.oracle_bcs(fx, bcs, ic), sources)


@pytest.mark.parametrize("ic", [cases.ic_gradient, cases.ic_oscillatory, cases.ic_smooth_euler])
def test_calc_fluxes_bit_exact(fx, ic):
    """test/divergence_test.jl: calc_fluxes!(du0,u0); exact-order mode must reproduce the CSC SpMV sums bit for bit"""
    m, semi = _semi(fx)
    u0 = ic(fx["points"], 0.0)
    du_ref = np.zeros_like(u0)
    _oracle(fx).calc_fluxes(u0, du_ref)
    du = np.zeros_like(u0)
    m.calc_fluxes_(du, u0, semi)
    assert cases.relerr(du, du_ref) <= RHS_TOL
    assert np.array_equal(du, du_ref), f"not bit-identical: max abs diff {np.abs(du - du_ref).max():.3e}"
    # accumulate semantics (du += ...)
    du2 = np.full_like(u0, 0.25)
    du2_ref = du2.copy()
    _oracle(fx).calc_fluxes(u0, du2_ref)
    m.calc_fluxes_(du2, u0, semi)
    assert np.array_equal(du2, du2_ref)
    semi.close()


def test_calc_fluxes_fma_mode_within_tolerance(fx):
    m, semi = _semi(fx, exact_order=False)
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    du_ref = np.zeros_like(u0)
    _oracle(fx).calc_fluxes(u0, du_ref)
    du = np.zeros_like(u0)
    m.calc_fluxes_(du, u0, semi)
    assert cases.relerr(du, du_ref) <= RHS_TOL
    semi.close()


def test_no_reorder_equals_hilbert(fx):
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    out = []
    for reorder in (None, "hilbert"):
        m, semi = _semi(fx, reorder=reorder)
        du = np.zeros_like(u0)
        m.calc_fluxes_(du, u0, semi)
        out.append(du)
        semi.close()
    assert np.array_equal(out[0], out[1])


def test_upwind_viscosity_source(fx):
    """test/upwind_viscosity_test.jl:69-78 through the CUDA path, checked against the oracle and the identity"""
    mk = lambda m, solver, eq, dom: dict(rv=m.SourceUpwindViscosityTominec(solver, eq, dom))
    m, semi = _semi(fx, sources=mk, diagnostics=True)
    src_o = orc.source_upwind(fx["dx_avg"])
    P = _oracle(fx, [src_o])
    u0 = cases.ic_gradient(fx["points"], 0.0)
    du_ref = np.zeros_like(u0)
    P.apply_source(0, u0, du_ref)
    du = np.zeros_like(u0)
    semi.source_terms.rv(du, u0, 0.0)
    assert cases.relerr(du, du_ref) <= RHS_TOL
    assert np.array_equal(du, du_ref)
    eps = semi.source_terms.rv.cache.eps
    assert np.array_equal(eps, src_o.arrays["eps"])
    Dx, Dy = fx["ops"]
    du1 = np.stack([-(Dx.T @ (eps * (Dx @ u0[v]))) - (Dy.T @ (eps * (Dy @ u0[v]))) for v in range(4)])
    np.testing.assert_allclose(du, du1, rtol=1e-9, atol=1e-9)
    semi.close()


def test_hyperviscosity_sources(fx):
    fx5 = cases.fixture_setup(p=5, N=3)
    m = _mft()
    fx5["ops"] = m.setup_ops.compute_flux_operator(fx5["points"], fx5["nb"], 5, 3)
    mk = lambda m, solver, eq, dom: dict(hv=m.SourceHyperviscosityFlyer(solver, eq, dom, k=2, c=1.0),
                                         hv2=m.SourceHyperviscosityTominec(solver, eq, dom, c=1.0))
    m, semi = _semi(fx5, sources=mk, ic=cases.ic_oscillatory)
    srcs = semi.source_terms
    o1 = orc.OracleSource(kind=orc.SRC_HV_FLYER, hv=orc.JuliaCSC(srcs.hv.hv_differentiation_matrix), gamma=srcs.hv.gamma)
    o2 = orc.OracleSource(kind=orc.SRC_HV_TOMINEC, hv=orc.JuliaCSC(srcs.hv2.hv_differentiation_matrix), gamma=srcs.hv2.gamma)
    P = _oracle(fx5, [o1, o2], ic=cases.ic_oscillatory)
    u0 = cases.ic_oscillatory(fx5["points"], 0.0)
    for i, s in enumerate((srcs.hv, srcs.hv2)):
        du_ref = np.full_like(u0, 0.5)
        P.apply_source(i, u0, du_ref)
        du = np.full_like(u0, 0.5)
        s(du, u0, 0.0)
        assert cases.relerr(du, du_ref) <= RHS_TOL
        assert np.array_equal(du, du_ref)
    semi.close()


def test_boundary_pass(fx):
    m, semi = _semi(fx, ic=cases.ic_smooth_euler, time_dependent=True)
    P = _oracle(fx, ic=cases.ic_smooth_euler)
    u = cases.ic_smooth_euler(fx["points"], 0.0) * 1.03
    du = np.full_like(u, 0.7)
    u_ref, du_ref = u.copy(), du.copy()
    P.boundary_pass(u_ref, du_ref, 0.4)
    m.calc_boundary_flux_(du, u, semi, 0.4)
    assert np.array_equal(u, u_ref) and np.array_equal(du, du_ref)
    semi.close()


SOURCE_SETS = {
    "none": (lambda m, s, e, d: {}, lambda fx: []),
    "upwind": (lambda m, s, e, d: dict(rv=m.SourceUpwindViscosityTominec(s, e, d)),
               lambda fx: [orc.source_upwind(fx["dx_avg"])]),
    "residual": (lambda m, s, e, d: dict(rv=m.SourceResidualViscosityTominec(s, e, d, polydeg=3)),
                 lambda fx: [orc.source_residual(fx["dx_avg"], polydeg=3)]),
}


@pytest.mark.parametrize("which", ["none", "upwind", "residual"])
def test_full_rhs(fx, which):
    """whole rhs! (BC pass, fluxes, source, BC pass); u in/out"""
    mk, mko = SOURCE_SETS[which]
    m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler, diagnostics=True)
    P = _oracle(fx, mko(fx), ic=cases.ic_smooth_euler)
    u = cases.ic_smooth_euler(fx["points"], 0.0) * 1.01
    u_ref = u.copy()
    du_ref = P.rhs(u_ref, 0.0)
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    assert np.array_equal(u, u_ref)
    err = cases.relerr(du, du_ref)
    assert err <= RHS_TOL, err
    if which != "residual":
        assert np.array_equal(du, du_ref)
    semi.close()


LAYOUTS = [dict(tile=0, pair_rows=0), dict(tile=0, pair_rows=1), dict(tile=0, pair_rows=3), dict(tile=3, tile_rows=11),
           dict(tile=7, tile_rows=11), dict(tile=15, tile_rows=11), dict(tile=15, tile_rows=11, stage_weights=4),
           dict(tile=11, tile_rows=11, stage_weights=15), dict(tile=15, tile_rows=22), dict(tile=15, tile_rows=22, stage_weights=15),
           dict(tile=7, tile_rows=22, stage_weights=4), dict(tile=3, tile_rows=44), dict(tile=15, tile_rows=44, stage_weights=15),
           dict(tile=15, tile_rows=42), dict(tile=13, tile_rows=12), dict(tile=14, tile_rows=41, pair_rows=0),
           dict(tile=15, tile_rows=11, prefetch_distance=0),
           dict(tile=15, tile_rows=22, prefetch_distance=8), dict(tile=15, tile_rows=11, prefetch_distance=4)]


@pytest.mark.parametrize("layout", LAYOUTS, ids=lambda d: "-".join(f"{k}{v}" for k, v in d.items()))
def test_operator_layouts_bit_identical(fx, layout):
    """Every device layout of the operators (sliced ELL, row pairs, union tiles with 1/2/4 rows per thread) applies the
    same sums in the same order: calc_fluxes! and the upwind-viscosity source are bit-identical to the oracle, the
    residual-viscosity rhs! agrees to 1e-12 (its global norms are reduced in a different association)."""
    u0 = cases.ic_smooth_euler(fx["points"], 0.0) * 1.01
    m, semi = _semi(fx, sources=SOURCE_SETS["upwind"][0], ic=cases.ic_smooth_euler, **layout)
    du_ref = np.full_like(u0, 0.125)
    _oracle(fx).calc_fluxes(u0, du_ref)
    du = np.full_like(u0, 0.125)
    m.calc_fluxes_(du, u0, semi)
    assert np.array_equal(du, du_ref)
    P = _oracle(fx, SOURCE_SETS["upwind"][1](fx), ic=cases.ic_smooth_euler)
    u, u_ref = u0.copy(), u0.copy()
    du_ref = P.rhs(u_ref, 0.0)
    m.rhs_(du, u, semi, 0.0)
    assert np.array_equal(u, u_ref) and np.array_equal(du, du_ref)
    du_s = np.zeros_like(u0)
    du_s_ref = np.zeros_like(u0)
    P.apply_source(0, u0, du_s_ref)
    semi.source_terms.rv(du_s, u0, 0.0)          # the source on its own (pass A without the flux part)
    assert np.array_equal(du_s, du_s_ref)
    semi.close()
    m, semi = _semi(fx, sources=SOURCE_SETS["residual"][0], ic=cases.ic_smooth_euler, **layout)
    P = _oracle(fx, SOURCE_SETS["residual"][1](fx), ic=cases.ic_smooth_euler)
    u, u_ref = u0.copy(), u0.copy()
    du_ref = P.rhs(u_ref, 0.0)
    m.rhs_(du, u, semi, 0.0)
    assert cases.relerr(du, du_ref) <= RHS_TOL
    semi.close()
    m, semi = _semi(fx, exact_order=False, sources=SOURCE_SETS["upwind"][0], ic=cases.ic_smooth_euler, **layout)
    u = u0.copy()
    m.rhs_(du, u, semi, 0.0)
    assert cases.relerr(du, _oracle(fx, SOURCE_SETS["upwind"][1](fx), ic=cases.ic_smooth_euler).rhs(u0.copy(), 0.0)) <= 1e-11
    semi.close()


def test_hv_then_upwind_order_of_sources(fx):
    """SourceTerms(hv=..., rv=...) as in test/upwind_viscosity_test.jl:52 -- sources see the du left by earlier ones"""
    mk = lambda m, s, e, d: dict(hv=m.SourceHyperviscosityTominec(s, e, d, c=1.0),
                                 rv=m.SourceResidualViscosityTominec(s, e, d, polydeg=3))
    m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler)
    hv = semi.source_terms.hv
    o_hv = orc.OracleSource(kind=orc.SRC_HV_TOMINEC, hv=orc.JuliaCSC(hv.hv_differentiation_matrix), gamma=hv.gamma)
    P = _oracle(fx, [o_hv, orc.source_residual(fx["dx_avg"], polydeg=3)], ic=cases.ic_smooth_euler)
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    u_ref = u.copy()
    du_ref = P.rhs(u_ref, 0.0)
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    assert cases.relerr(du, du_ref) <= RHS_TOL
    semi.close()


def test_residual_viscosity_with_history_and_time_integration(fx):
    """config-2 shape on the fixture: Euler + residual viscosity + HistoryCallback + SSPRK33, 30 steps"""
    mk, mko = SOURCE_SETS["residual"]
    m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler, diagnostics=True)
    src_o = mko(fx)[0]
    P = _oracle(fx, [src_o], ic=cases.ic_smooth_euler)
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    dt = 0.1 * fx["dx_min"] / 3.0
    nsteps = 30
    u_ref, t_ref = P.solve_ssprk33(u0, 0.0, dt, nsteps, approx_order=3)
    ode = m.semidiscretize(semi, (0.0, nsteps * dt))
    assert np.array_equal(ode.u0, u0)
    sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=m.HistoryCallback(approx_order=3), nsteps=nsteps)
    assert abs(sol.t - t_ref) < 1e-15
    err = cases.relerr(sol.u, u_ref)
    assert err <= STEP_TOL, err
    # the limiter really was active (eps_rv < eps_uw somewhere) so the residual path is exercised
    c = semi.source_terms.rv.cache
    assert (c.eps_c == 0).any() and (c.eps_c == 1).any()
    np.testing.assert_allclose(c.approx_du, src_o.arrays["approx_du"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(c.norms, P.residual_norms(0, u_ref, np.zeros_like(u_ref)), rtol=1e-12)
    semi.close()


def test_residual_norm_conventions(fx):
    """documented switches: ode_mean divisor V*N vs N, lexicographic vs per-component maximum"""
    mk, _ = SOURCE_SETS["residual"]
    for vn, lex in [(True, True), (False, False), (True, False), (False, True)]:
        m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler, diagnostics=True, mean_divisor_vn=vn,
                        max_lexicographic=lex)
        P = _oracle(fx, [orc.source_residual(fx["dx_avg"], polydeg=3, mean_divisor_vn=vn, max_lexicographic=lex)],
                    ic=cases.ic_smooth_euler)
        u = cases.ic_smooth_euler(fx["points"], 0.0)
        u_ref = u.copy()
        du_ref = P.rhs(u_ref, 0.0)
        du = np.empty_like(u)
        m.rhs_(du, u, semi, 0.0)
        np.testing.assert_allclose(semi.source_terms.rv.cache.norms, P.residual_norms(0, u_ref, du_ref), rtol=1e-13)
        assert cases.relerr(du, du_ref) <= RHS_TOL
        semi.close()


def test_advection_config1(fx):
    """BASELINE config 1: 2-D linear advection on the fixture cloud, PHS 5 / degree 3, hyperviscosity, SSPRK33"""
    m = _mft()
    fx5 = cases.fixture_setup(p=5, N=3)
    ops = m.setup_ops.compute_flux_operator(fx5["points"], fx5["nb"], 5, 3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(5)))
    solver = m.PointCloudSolver(basis)
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.LinearScalarAdvectionEquation2D(1.0, 0.5)
    ic = cases.ic_bump_advection
    bc = dict(inlet=m.BoundaryConditionDirichlet(ic), outlet=m.BoundaryConditionDoNothing(),
              top=m.BoundaryConditionDoNothing(), bottom=m.BoundaryConditionDoNothing(), cyl=m.BoundaryConditionDoNothing())
    srcs = m.SourceTerms(hv=m.SourceHyperviscosityFlyer(solver, eq, domain, k=2, c=1.0),
                         hv2=m.SourceHyperviscosityTominec(solver, eq, domain, c=1.0))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs, operators=ops)
    o1 = orc.OracleSource(kind=orc.SRC_HV_FLYER, hv=orc.JuliaCSC(srcs.hv.hv_differentiation_matrix), gamma=srcs.hv.gamma)
    o2 = orc.OracleSource(kind=orc.SRC_HV_TOMINEC, hv=orc.JuliaCSC(srcs.hv2.hv_differentiation_matrix), gamma=srcs.hv2.gamma)
    obc = [orc.OracleBC(orc.BC_DIRICHLET, fx5["bidx"][0], fx5["bnrm"][0], value_fn=lambda x, t: ic(x, t))]
    obc += [orc.OracleBC(orc.BC_DO_NOTHING, fx5["bidx"][g], fx5["bnrm"][g]) for g in (1, 3, 2, 4)]
    P = orc.OracleProblem(fx5["points"], 1, orc.EQ_ADVECTION2D, [1.0, 0.5], ops[0], ops[1], obc, [o1, o2])
    u = ic(fx5["points"], 0.0)
    u_ref = u.copy()
    du_ref = P.rhs(u_ref, 0.0)
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    assert np.array_equal(du, du_ref) and np.array_equal(u, u_ref)
    dt = 0.1 * fx5["dx_min"]
    u_end, _ = P.solve_ssprk33(ic(fx5["points"], 0.0), 0.0, dt, 100)
    sol = m.solve(m.semidiscretize(semi, (0.0, 100 * dt)), m.SSPRK33(), dt=dt, nsteps=100)
    err = cases.relerr(sol.u, u_end)
    assert err <= STEP_TOL, err
    semi.close()


def test_synthetic_cloud_medium(fx):
    """jittered-lattice cloud (the generator of BASELINE configs 2-4) at a size the oracle finishes in seconds"""
    m = _mft()
    cl = m.cloud.jittered_lattice(96, 80, 10.0, 10.0 * 80 / 96, seed=0)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(diagnostics=True))
    names = dict(left=1, right=2, bottom=3, top=4)
    domain = m.PointCloudDomain(solver, cl, names)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, cases.GAMMA, center=(5.0, 4.0))
    bc = {k: m.BoundaryConditionDirichlet(ic) for k in names}
    srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, c_rv=1.0, c_uw=1.0, polydeg=3))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs)
    ops = semi.cache.rbf_differentiation_matrices
    pd = domain.pd
    nb_o, dxmin_o, dxavg_o = orc.point_data(pd.points, pd.num_neighbors)
    assert np.array_equal(nb_o, pd.neighbors) and dxmin_o == pd.dx_min and dxavg_o == pd.dx_avg
    obc = [orc.OracleBC(orc.BC_DIRICHLET, domain.boundary_tags[k].idx, domain.boundary_tags[k].normals,
                        value_fn=lambda x, t: ic(x, t)) for k in names]
    src_o = orc.source_residual(pd.dx_avg, polydeg=3)
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], obc, [src_o])
    u0 = ic(pd.points, 0.0)
    dt = 0.1 * pd.dx_min / 8.0
    u_ref, _ = P.solve_ssprk33(u0, 0.0, dt, 12, approx_order=3)
    sol = m.solve(m.semidiscretize(semi, (0.0, 12 * dt)), m.SSPRK33(), dt=dt, callback=m.HistoryCallback(3), nsteps=12)
    err = cases.relerr(sol.u, u_ref)
    assert err <= STEP_TOL, err
    # single rhs on the evolved state
    u = sol.u.copy()
    ur = sol.u.copy()
    du_ref = P.rhs(ur, 12 * dt)
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 12 * dt)
    assert cases.relerr(du, du_ref) <= RHS_TOL
    semi.close()


def test_error_paths():
    import ctypes as C
    m = _mft()
    lib = m._lib.load()
    ctx = C.c_void_p()
    assert lib.mft_ctx_create(C.byref(ctx), 0, 10, 0, 3, 2, 5) == -4     # nvars 3 unsupported
    assert lib.mft_ctx_create(C.byref(ctx), 0, 0, 0, 4, 2, 5) == -1
    assert lib.mft_ctx_create(C.byref(ctx), 0, 10, 0, 1, 2, 5) == 0
    prm = np.array([1.0, 0.5])
    assert lib.mft_set_equation(ctx, m._lib.EQ_ADVECTION2D, m._lib.ptr(prm), 2) == 0
    # residual viscosity is Euler-only in the reference (hyperviscosity.jl:289-291)
    p4 = np.array([1.0, 1.0, 0.1, 3.0])
    assert lib.mft_add_source(ctx, m._lib.SRC_RESIDUAL, m._lib.ptr(p4), 4, None, None, None) == -4
    assert b"Euler" in lib.mft_last_error()
    # compute without operators -> EINVAL, never a silent fallback
    u = np.zeros((1, 10))
    assert lib.mft_rhs(ctx, 0.0, m._lib.soa_ptrs(u), m._lib.soa_ptrs(u.copy()), 0) == -1
    assert lib.mft_ctx_destroy(ctx) == 0


def test_ell_operator_input_equals_csc_input(fx):
    """mft_set_operator_ell (neighbour / weight tables) and mft_set_operator_csc (Julia CSC fields) build the same device
    operator: identical rhs! bit for bit (Euler + upwind viscosity exercises D and D')."""
    import ctypes as C

    m = _mft()
    L = m._lib
    lib = m.load()
    mk = lambda m_, s, e, d: dict(rv=m_.SourceUpwindViscosityTominec(s, e, d))
    m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler)
    u = cases.ic_smooth_euler(fx["points"], 0.0) * 1.01
    du_csc = np.empty_like(u)
    m.rhs_(du_csc, u.copy(), semi, 0.0)
    # same problem through the ELL entry point
    nb = fx["nb"]
    n, k = nb.shape
    Dx, Dy = (A.tocsr() for A in fx["ops"])
    wx = np.empty((n, k))
    wy = np.empty((n, k))
    for i in range(n):
        lx = dict(zip(Dx.indices[Dx.indptr[i]:Dx.indptr[i + 1]], Dx.data[Dx.indptr[i]:Dx.indptr[i + 1]]))
        ly = dict(zip(Dy.indices[Dy.indptr[i]:Dy.indptr[i + 1]], Dy.data[Dy.indptr[i]:Dy.indptr[i + 1]]))
        wx[i] = [lx[j] for j in nb[i]]
        wy[i] = [ly[j] for j in nb[i]]
    ctx = C.c_void_p()
    L.check(lib.mft_ctx_create(C.byref(ctx), 0, n, 0, 4, 2, k))
    g = np.array([cases.GAMMA])
    L.check(lib.mft_set_equation(ctx, L.EQ_EULER2D, L.ptr(g), 1))
    nbr1 = np.ascontiguousarray(nb + 1)
    L.check(lib.mft_set_operator_ell(ctx, L.ptr(nbr1), L.ptr(wx), L.ptr(wy)))
    for name, bc, tag in semi._bc_groups:
        idx1 = np.ascontiguousarray(tag.idx + 1)
        nrm = np.ascontiguousarray(tag.normals)
        vals = None
        if bc.kind == L.BC_DIRICHLET:
            vals = np.ascontiguousarray(cases.ic_smooth_euler(fx["points"][tag.idx], 0.0))
        L.check(lib.mft_add_boundary(ctx, bc.kind, len(idx1), L.ptr(idx1), L.ptr(nrm), L.ptr(vals)))
    prm = np.array([1.0, fx["dx_avg"]])
    L.check(lib.mft_add_source(ctx, L.SRC_UPWIND, L.ptr(prm), 2, None, None, None))
    uu = u.copy()
    du_ell = np.empty_like(u)
    L.check(lib.mft_rhs(ctx, 0.0, L.soa_ptrs(uu), L.soa_ptrs(du_ell), L.MEM_HOST))
    L.check(lib.mft_ctx_destroy(ctx))
    assert np.array_equal(du_ell, du_csc)
    semi.close()


def test_ssprk43_step_and_adaptive_run(fx):
    """SSPRK43 (the integrator the reference names, rbfsolver_test.jl:104-107): one step (state + embedded error estimate)
    and an adaptive run with accepted AND rejected steps, residual viscosity + history callback, against the oracle"""
    mk, mko = SOURCE_SETS["residual"]
    m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler)
    P = _oracle(fx, mko(fx), ic=cases.ic_smooth_euler)
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    # single step
    k = P.rhs(u0.copy(), 0.0)
    ub = u0.copy()
    P.boundary_pass(ub, np.zeros_like(ub), 0.0)
    un, kn, eest_ref = orc._ssprk43_step(P, ub, k, 0.0, 1e-3, 1e-6, 1e-6)
    import ctypes as C
    lib, L = m.load(), m._lib
    L.check(lib.mft_upload_state(semi.ctx, L.soa_ptrs(u0)))
    ss, cnt = C.c_double(), C.c_int64()
    L.check(lib.mft_ssprk43_step(semi.ctx, 0.0, 1e-3, 1e-6, 1e-6, C.byref(ss), C.byref(cnt)))
    assert cnt.value == u0.size
    eest = np.sqrt(ss.value / cnt.value)
    assert abs(eest - eest_ref) <= 1e-10 * eest_ref
    L.check(lib.mft_step_commit(semi.ctx, 1))
    u1 = np.empty_like(u0)
    L.check(lib.mft_download_state(semi.ctx, L.soa_ptrs(u1)))
    assert cases.relerr(u1, un) <= 1e-12
    # a rejected step leaves no trace
    L.check(lib.mft_ssprk43_step(semi.ctx, 1e-3, 5e-2, 1e-6, 1e-6, C.byref(ss), C.byref(cnt)))
    L.check(lib.mft_step_commit(semi.ctx, 0))
    u2 = np.empty_like(u0)
    L.check(lib.mft_download_state(semi.ctx, L.soa_ptrs(u2)))
    assert np.array_equal(u2, u1)
    semi.close()
    # adaptive run (same controller on both sides)
    m, semi = _semi(fx, sources=mk, ic=cases.ic_smooth_euler)
    P = _oracle(fx, mko(fx), ic=cases.ic_smooth_euler)
    u_ref, t_ref, log_ref = orc.solve_ssprk43(P, u0, 0.0, 0.004, 1e-3, abstol=1e-6, reltol=1e-6, approx_order=3)
    sol = m.solve(m.semidiscretize(semi, (0.0, 0.004)), m.SSPRK43(), dt=1e-3, abstol=1e-6, reltol=1e-6,
                  callback=m.HistoryCallback(approx_order=3))
    assert [a[3] for a in sol.log] == [a[3] for a in log_ref], "accept/reject sequence differs"
    assert any(not a[3] for a in sol.log) and sum(a[3] for a in sol.log) >= 5
    np.testing.assert_allclose([a[1] for a in sol.log], [a[1] for a in log_ref], rtol=1e-9)
    assert abs(sol.t - t_ref) < 1e-15
    assert cases.relerr(sol.u, u_ref) <= STEP_TOL
    semi.close()
