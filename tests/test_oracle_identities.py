"""Pins the CPU oracle (oracle/) against the reference's own test identities -- the only fixtures the reference holds
for this path:
  test/divergence_test.jl:66-82, test/upwind_viscosity_test.jl:69-78, test/history_test.jl:14-32,
plus polynomial-reproduction known answers standing in for the two operator tests that need the un-vendored
RadialBasisFiniteDifferences package (test/first_order_test.jl:67-73, test/hyperviscosity_test.jl:72-82).
Runs on CPU."""
import numpy as np
import pytest

import cases
from cases import orc


@pytest.fixture(scope="module")
def fx():
    s = cases.fixture_setup(p=3, N=3)
    s["ops"] = orc.compute_flux_operator(s["points"], s["nb"], 3, 3)
    return s


def test_fixture_cloud_facts(fx):
    # 2154 points, groups 40/40/60/60/26, k = 20 for degree 3, self is the first neighbour
    assert fx["points"].shape == (2154, 2)
    assert [len(b) for b in fx["bidx"]] == [40, 40, 60, 60, 26]
    assert fx["nv"] == 20
    assert (fx["nb"][:, 0] == np.arange(2154)).all()
    assert abs(fx["dx_min"] - 0.03490) < 1e-4 and abs(fx["dx_avg"] - 0.05290) < 1e-4


def test_num_neighbors_formula():
    # geometry_primatives.jl:197-198: degree 2->15, 3->20, 4->30, 5->42, 6->56
    assert [orc.num_neighbors(N) for N in (2, 3, 4, 5, 6)] == [15, 20, 30, 42, 56]


def test_operator_sparsity_is_the_neighbor_table(fx):
    Dx, Dy = fx["ops"]
    for A in (Dx, Dy):
        R = A.tocsr()
        R.sort_indices()
        assert (np.diff(R.indptr) == 20).all()          # explicit zeros are kept by sparse(I,J,V)
        got = R.indices.reshape(-1, 20)
        assert (got == np.sort(fx["nb"], axis=1)).all()


def test_polynomial_reproduction(fx):
    Dx, Dy = fx["ops"]
    x, y = fx["points"][:, 0], fx["points"][:, 1]
    for f, fx_, fy_ in [(np.ones_like(x), 0 * x, 0 * x), (x, 1 + 0 * x, 0 * x), (y, 0 * x, 1 + 0 * x),
                        (x * y, y, x), (x ** 3, 3 * x * x, 0 * x), (y ** 3 + x * x * y, 2 * x * y, 3 * y * y + x * x)]:
        assert np.abs(Dx @ f - fx_).max() < 5e-12
        assert np.abs(Dy @ f - fy_).max() < 5e-12


def test_laplacian_known_answer(fx):
    L = orc.compute_flux_operator(fx["points"], fx["nb"], 5, 3, 2)
    lap = L[0] + L[1]
    x, y = fx["points"][:, 0], fx["points"][:, 1]
    assert np.abs(lap @ (x * x + y * y) - 4.0).max() < 1e-9


def _problem(fx, sources=(), bcs=None, ic=cases.ic_gradient):
    bc = cases.oracle_bcs(fx, bcs or cases.DIVERGENCE_TEST_BCS, ic)
    return orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], fx["ops"][0], fx["ops"][1], bc, sources)


def test_divergence_identity(fx):
    """test/divergence_test.jl:66-82: calc_fluxes!(du0,u0) == -Dx*F(u) - Dy*G(u)"""
    P = _problem(fx)
    u0 = cases.ic_gradient(fx["points"], 0.0)
    du0 = np.zeros_like(u0)
    P.calc_fluxes(u0, du0)
    F, G = P.flux(u0, 0), P.flux(u0, 1)
    Dx, Dy = fx["ops"]
    du1 = np.stack([-(Dx @ F[v]) - (Dy @ G[v]) for v in range(4)])
    np.testing.assert_allclose(du0, du1, rtol=1e-10, atol=1e-10)
    # the Trixi flux formula against a plain-numpy statement of the Euler fluxes
    rho, m1, m2, E = u0
    v1, v2 = m1 / rho, m2 / rho
    p = (cases.GAMMA - 1) * (E - 0.5 * (m1 * v1 + m2 * v2))
    np.testing.assert_allclose(F, np.stack([m1, m1 * v1 + p, m1 * v2, (E + p) * v1]), rtol=1e-14)
    np.testing.assert_allclose(G, np.stack([m2, m2 * v1, m2 * v2 + p, (E + p) * v2]), rtol=1e-14)


def test_upwind_viscosity_identity(fx):
    """test/upwind_viscosity_test.jl:69-78: source(du0,u0) == -Dx'(eps.*(Dx u)) - Dy'(eps.*(Dy u))"""
    src = orc.source_upwind(fx["dx_avg"])
    P = _problem(fx, sources=[src])
    u0 = cases.ic_gradient(fx["points"], 0.0)
    du0 = np.zeros_like(u0)
    P.apply_source(0, u0, du0)
    eps = src.arrays["eps"]
    Dx, Dy = fx["ops"]
    du1 = np.stack([-(Dx.T @ (eps * (Dx @ u0[v]))) - (Dy.T @ (eps * (Dy @ u0[v]))) for v in range(4)])
    np.testing.assert_allclose(du0, du1, rtol=1e-9, atol=1e-9)
    # eps itself: c_uw * 0.5 * dx_avg * (|v| + c)
    rho, m1, m2, E = u0
    v1, v2 = m1 / rho, m2 / rho
    p = (cases.GAMMA - 1) * (E - 0.5 * (m1 * v1 + m2 * v2))
    np.testing.assert_allclose(eps, 0.5 * fx["dx_avg"] * (np.hypot(v1, v2) + np.sqrt(cases.GAMMA * p / rho)), rtol=1e-13)
    assert (src.arrays["eps_c"] == 1).all()


def test_history_known_answer():
    """test/history_test.jl:14-32: six samples of u = t*1 -> reconstructed du/dt == 1"""
    order = 5
    n = order + 1
    src = orc.source_residual(1.0, polydeg=order)
    P = orc.OracleProblem(np.zeros((n, 2)), 1, orc.EQ_ADVECTION2D, [1.0, 0.5], np.eye(n), np.eye(n), [], [src])
    # drive the C functions the way the test drives shift_soln_history! / time_deriv_weights!
    u = np.zeros((1, n))
    t = 0.0
    for i in range(order + 1):
        t += 1.0
        u += 1.0
        P.history_callback(u, t, i + 1, order)
    np.testing.assert_allclose(src.arrays["approx_du"], np.ones((1, n)), rtol=1e-9)
    np.testing.assert_allclose(src.arrays["time_history"], np.arange(6.0, 0.0, -1.0))


def test_time_deriv_weights_vs_numpy():
    t = np.array([0.5, 0.47, 0.45, 0.41])
    ts = t / np.abs(t).max()
    A = np.stack([ts ** k for k in range(4)], axis=1)
    b = np.array([k * ts[0] ** (k - 1) if k > 0 else 0.0 for k in range(4)])
    w = np.linalg.solve(A.T, b) / np.abs(t).max()
    np.testing.assert_allclose(orc.time_deriv_weights(t), w, rtol=1e-9)
    # derivative of a cubic is reproduced exactly
    f = lambda s: 2.0 - s + 3.0 * s ** 2 - 0.5 * s ** 3
    assert abs(orc.time_deriv_weights(t) @ f(t) - (-1 + 6 * t[0] - 1.5 * t[0] ** 2)) < 1e-9


def test_rhs_composition_and_u_mutation(fx):
    """rhs! = reset, BC pass, fluxes, sources (in order), BC pass; u is mutated at boundary points only."""
    src_hv = orc.source_hyperviscosity_tominec(fx["points"], fx["nb"], 3, 3, fx["dx_min"])
    src_uw = orc.source_upwind(fx["dx_avg"])
    P = _problem(fx, sources=[src_hv, src_uw], ic=cases.ic_smooth_euler)
    u = cases.ic_smooth_euler(fx["points"], 0.0) * 1.01   # off the Dirichlet data so that the BC visibly acts
    u_in = u.copy()
    du = P.rhs(u, 0.3)
    # manual composition out of the primitives
    u2 = u_in.copy()
    du2 = np.zeros_like(u2)
    P.boundary_pass(u2, du2, 0.3)
    P.calc_fluxes(u2, du2)
    P.apply_source(0, u2, du2)
    P.apply_source(1, u2, du2)
    P.boundary_pass(u2, du2, 0.3)
    assert np.array_equal(du, du2) and np.array_equal(u, u2)
    bpts = np.concatenate([fx["bidx"][g] for g in (0, 2, 3, 4)])   # inlet + slip walls
    interior = np.setdiff1d(np.arange(u.shape[1]), bpts)
    assert np.array_equal(u[:, interior], u_in[:, interior])
    inlet = fx["bidx"][0]
    np.testing.assert_array_equal(u[:, inlet], cases.ic_smooth_euler(fx["points"][inlet], 0.3))
    assert (du[:, inlet] == 0).all()
    walls = np.concatenate([fx["bidx"][g] for g in (2, 3, 4)])
    assert (du[1:3, walls] == 0).all()
    # slip wall: normal momentum removed
    for g in (2, 3, 4):
        nrm = fx["bnrm"][g] / np.linalg.norm(fx["bnrm"][g], axis=1, keepdims=True)
        mn = u[1, fx["bidx"][g]] * nrm[:, 0] + u[2, fx["bidx"][g]] * nrm[:, 1]
        assert np.abs(mn).max() < 1e-14


def test_residual_viscosity_limiter_semantics(fx):
    src = orc.source_residual(fx["dx_avg"], polydeg=3)
    P = _problem(fx, sources=[src], ic=cases.ic_smooth_euler)
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    # success_iter == 0 -> eps == eps_uw everywhere (hyperviscosity.jl:333)
    P.rhs(u.copy(), 0.0)
    assert np.array_equal(src.arrays["eps"], src.arrays["eps_uw"]) and (src.arrays["eps_c"] == 1).all()
    # after two pushes approx_du is a backward difference and eps = min(eps_rv, eps_uw)
    P.history_callback(u, 0.0, 0, 3)
    P.history_callback(u * 1.001, 0.01, 1, 3)
    np.testing.assert_allclose(src.arrays["approx_du"], (u * 1.001 - u) / 0.01, rtol=1e-6, atol=1e-9)
    P.rhs(u.copy(), 0.01)
    a = src.arrays
    assert np.array_equal(a["eps"], np.minimum(a["eps_rv"], a["eps_uw"]))
    assert set(np.unique(a["eps_c"])) <= {0, 1}
    # norm conventions: reference defaults (divide by V*N, lexicographic maximum) vs the "intended" ones
    du = np.zeros_like(u)
    nrm_ref = P.residual_norms(0, u, du)
    mean_ref = u.sum(axis=1) / (4 * u.shape[1])
    dev = np.abs(u - mean_ref[:, None])
    istar = np.lexsort(dev[::-1])[-1]            # lexicographic argmax, first component most significant
    np.testing.assert_allclose(nrm_ref, dev[:, istar], rtol=1e-12)
    src2 = orc.source_residual(fx["dx_avg"], polydeg=3, mean_divisor_vn=False, max_lexicographic=False)
    P2 = _problem(fx, sources=[src2], ic=cases.ic_smooth_euler)
    nrm2 = P2.residual_norms(0, u, du)
    np.testing.assert_allclose(nrm2, np.abs(u - u.mean(axis=1, keepdims=True)).max(axis=1), rtol=1e-12)


def test_pairwise_sum_matches_fsum():
    rng = np.random.default_rng(5)
    a = rng.standard_normal(100003)
    import math
    assert abs(orc.sum_pairwise(a) - math.fsum(a)) < 1e-10


def test_ssprk43_error_estimate_is_third_order(fx):
    """known answer for the SSPRK43 restatement (OrdinaryDiffEq is unpinned): the embedded estimate of a 3rd-order pair
    scales as dt^3 on a smooth, BC-consistent state, and a state with rhs == 0 gives EEst == 0"""
    bcs = dict(inlet="dirichlet", outlet="dirichlet", top="dirichlet", bottom="dirichlet", cyl="dirichlet")
    P = _problem(fx, bcs=bcs, ic=cases.ic_smooth_euler)
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    k = P.rhs(u0.copy(), 0.0)
    e = [orc._ssprk43_step(P, u0.copy(), k, 0.0, dt, 1e-8, 1e-8)[2] for dt in (4e-3, 2e-3, 1e-3)]
    assert 7.5 < e[0] / e[1] < 8.5 and 7.5 < e[1] / e[2] < 8.5
    uc = np.ascontiguousarray(np.tile(np.array([[1.0], [0.3], [-0.2], [2.5]]), (1, u0.shape[1])))
    Pc = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], fx["ops"][0], fx["ops"][1], [], [])
    kc = Pc.rhs(uc.copy(), 0.0)
    un, kn, ee = orc._ssprk43_step(Pc, uc.copy(), kc, 0.0, 1e-2, 1e-8, 1e-8)
    assert ee < 1e-3 and np.abs(un - uc).max() < 1e-10     # constants are annihilated up to roundoff of D*1
    # adaptive run: the controller accepts/rejects and ends exactly at t1
    u, t, log = orc.solve_ssprk43(P, u0, 0.0, 0.03, 1e-3, abstol=1e-6, reltol=1e-6)
    assert abs(t - 0.03) < 1e-15 and len(log) >= 5 and np.isfinite(u).all()


def test_best_effort_cpu_baseline_matches_the_oracle():
    """oracle/mft_cpu_fast.c (CPU baseline (ii) of SURVEY.md 8d: fused, row-parallel, AoS) computes the same Euler +
    residual-viscosity rhs! as the reference-structured oracle: u identical, du to 1e-12 (only the association of the
    global mean differs) -- with 1 thread and with all host cores."""
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    ic = cases.ic_smooth_euler
    bcs = dict(inlet="dirichlet", outlet="dirichlet", top="dirichlet", bottom="dirichlet", cyl="dirichlet")
    for si in (0, 5):
        src = orc.source_residual(fx["dx_avg"], polydeg=3)
        src.success_iter = si
        P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, bcs, ic), [src])
        u0 = ic(fx["points"], 0.0) * 1.01
        u_ref = u0.copy()
        du_ref = P.rhs(u_ref, 0.0)
        bidx = np.concatenate(fx["bidx"])
        F = orc.FastCpuProblem(ops[0], ops[1], cases.GAMMA, fx["dx_avg"], bidx, ic(fx["points"][bidx], 0.0), success_iter=si)
        for threads in (1, 3, orc.fast_lib().fast_max_threads()):
            orc.fast_lib().fast_set_threads(int(threads))
            u = u0.copy()
            du = F.rhs(u)
            assert np.array_equal(u, u_ref)
            assert cases.relerr(du, du_ref) <= 1e-12
