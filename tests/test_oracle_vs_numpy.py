"""Two independent restatements of the reference agree: the C oracle (oracle/mft_oracle.c, the checker of every parity
test) against the numpy/scipy restatement written separately from the Julia sources (oracle/mft_ref_numpy.py).

This covers exactly what the reference's own tests do NOT pin ("parity unpinned"): both BC passes
and the mutation of u, whole-rhs! with sources in NamedTuple order, update_upwind_visc! with its p<0 / rho<0 clamp,
update_residual_visc! incl. the ode_mean divisor and the lexicographic maximum, update_visc! in all three branches, the
history callback (ring shift, time_deriv_weights!, update_approx_du!) and the SSPRK33 loop around them.  Agreement is to
rounding (1e-12 normwise per rhs!, 1e-10 after the time loop): the two differ in summation order inside a mat-vec."""
import numpy as np
import pytest

import cases
from cases import orc

import mft_ref_numpy as ref


@pytest.fixture(scope="module")
def fx():
    s = cases.fixture_setup(p=3, N=3)
    s["ops"] = orc.compute_flux_operator(s["points"], s["nb"], 3, 3)
    return s


def _ref_bcs(fx, spec, ic):
    mk = dict(dirichlet=lambda: ref.Dirichlet(lambda x, t: ic(x, t)), slip=ref.SlipWall, nothing=ref.DoNothing)
    return [(mk[kind](), fx["bidx"][cases.BOUNDARY_NAMES[name] - 1], fx["bnrm"][cases.BOUNDARY_NAMES[name] - 1])
            for name, kind in spec.items()]


def _pair(fx, osources, rsources, spec=cases.DIVERGENCE_TEST_BCS, ic=cases.ic_smooth_euler):
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], fx["ops"][0], fx["ops"][1],
                          cases.oracle_bcs(fx, spec, ic), osources)
    R = ref.RefProblem(fx["points"], ref.Euler2D(cases.GAMMA), fx["ops"][0], fx["ops"][1], _ref_bcs(fx, spec, ic), rsources)
    return P, R


def test_flux_divergence_and_both_bc_passes(fx):
    P, R = _pair(fx, [], [])
    u0 = cases.ic_smooth_euler(fx["points"], 0.0) * (1.0 + 0.02 * np.sin(5 * fx["points"][:, 1]))   # off the Dirichlet data
    ua, ub = u0.copy(), u0.copy()
    da, db = P.rhs(ua, 0.3), R.rhs(ub, 0.3)
    assert cases.relerr(ua, ub) <= 1e-15          # BC-imposed u (slip-wall projection: one fused multiply-add apart)
    assert (ua != u0).any()
    assert cases.relerr(da, db) <= 1e-12


def test_sources_in_order_hyperviscosity_then_upwind(fx):
    hv = orc.source_hyperviscosity_tominec(fx["points"], fx["nb"], 3, 3, fx["dx_min"])
    ops2 = orc.compute_flux_operator(fx["points"], fx["nb"], 3, 3, 2)
    lap = (ops2[0] + ops2[1]).tocsc()
    n = len(fx["points"])
    rs = [ref.Hyperviscosity(lap.T @ lap, fx["dx_min"] ** 4.5), ref.TominecViscosity(n, 4, fx["dx_avg"], residual=False)]
    P, R = _pair(fx, [hv, orc.source_upwind(fx["dx_avg"])], rs)
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    u0[3, 400:420] *= 0.05                         # negative pressure pocket: the p < 0 clamp of update_upwind_visc!
    u0[0, 1000:1003] *= -1.0                       # and the rho < 0 one
    da, db = P.rhs(u0.copy(), 0.0), R.rhs(u0.copy(), 0.0)
    a = P.sources[1].arrays
    np.testing.assert_allclose(a["eps_uw"], rs[1].eps_uw, rtol=1e-14, atol=0)
    assert (a["eps_c"] == 1).all() and np.array_equal(a["eps"], a["eps_uw"])
    assert cases.relerr(da, db) <= 1e-12


def test_residual_viscosity_with_history_all_branches(fx):
    n = len(fx["points"])
    osrc = orc.source_residual(fx["dx_avg"], polydeg=3)
    rsrc = ref.TominecViscosity(n, 4, fx["dx_avg"], residual=True, polydeg=3)
    P, R = _pair(fx, [osrc], [rsrc])
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    # success_iter == 0: upwind branch everywhere
    da, db = P.rhs(u.copy(), 0.0), R.rhs(u.copy(), 0.0)
    assert (osrc.arrays["eps_c"] == 1).all() and (rsrc.eps_c == 1).all() and cases.relerr(da, db) <= 1e-12
    # three pushes of a drifting state -> approx_du from time_deriv_weights!, min(eps_rv, eps_uw) with both outcomes
    for it, t in enumerate((0.0, 0.013, 0.024, 0.04)):
        w = u * (1.0 + 0.5 * t * np.cos(fx["points"][:, 0]))
        P.history_callback(w, t, it, 3)
        R.history_callback(w, t, it, 3)
    np.testing.assert_allclose(osrc.arrays["time_weights"], rsrc.time_weights, rtol=1e-10)
    assert cases.relerr(osrc.arrays["approx_du"], rsrc.approx_du) <= 1e-10
    u2 = u * (1.0 + 0.02 * np.cos(fx["points"][:, 0]))
    u2[3, 77] = np.inf                             # infinite energy at one point: its eps_uw and eps_rv are both non-finite,
                                                   # its stencil neighbours get a NaN eps_rv with a finite eps_uw
    ua, ub = u2.copy(), u2.copy()
    with np.errstate(all="ignore"):
        da, db = P.rhs(ua, 0.04), R.rhs(ub, 0.04)
    a = osrc.arrays
    finite = np.isfinite(rsrc.eps_rv)
    np.testing.assert_allclose(a["eps_rv"][finite], rsrc.eps_rv[finite], rtol=1e-9)
    assert np.array_equal(np.isfinite(a["eps_rv"]), finite)
    assert np.array_equal(a["eps_c"], rsrc.eps_c)
    assert 0 < (a["eps_c"] == 0).sum() and 0 < (a["eps_c"] == 1).sum()
    assert a["eps_c"][77] == 2 and a["eps"][77] == ref.EPS      # eps_rv and eps_uw both non-finite -> Base.eps()
    np.testing.assert_allclose(a["eps"][np.isfinite(a["eps"])], rsrc.eps[np.isfinite(rsrc.eps)], rtol=1e-9)
    np.testing.assert_allclose(P.residual_norms(0, u, np.zeros_like(u)),
                               _norms(rsrc, u), rtol=1e-13)


def _norms(rsrc, u):
    rsrc.update_residual_visc(np.zeros_like(u), u)
    return rsrc.n_inf_norms


def test_update_visc_semantics_of_the_restatement():
    """all three branches of update_visc! (hyperviscosity.jl:331-349) incl. Julia's NaN-propagating min"""
    s = ref.TominecViscosity(4, 4, 0.1, residual=True)
    s.success_iter = 3
    s.eps_rv[:] = [np.nan, 1.0, np.inf, 0.5]
    s.eps_uw[:] = [np.inf, 2.0, 1.0, np.nan]
    s.update_visc()
    assert list(s.eps_c) == [2, 0, 1, 1] and s.eps[0] == ref.EPS and s.eps[1] == 1.0 and s.eps[2] == 1.0 and np.isnan(s.eps[3])


def test_ssprk33_with_residual_viscosity_and_history(fx):
    n = len(fx["points"])
    osrc = orc.source_residual(fx["dx_avg"], polydeg=3)
    rsrc = ref.TominecViscosity(n, 4, fx["dx_avg"], residual=True, polydeg=3)
    P, R = _pair(fx, [osrc], [rsrc])
    u0 = cases.ic_smooth_euler(fx["points"], 0.0)
    dt = 0.1 * fx["dx_min"] / 3.0
    ua, ta = P.solve_ssprk33(u0, 0.0, dt, 8, approx_order=3)
    ub, tb = R.solve_ssprk33(u0, 0.0, dt, 8, approx_order=3)
    assert abs(ta - tb) <= 1e-15 and np.isfinite(ua).all()
    assert cases.relerr(ua, ub) <= 1e-10
    assert np.array_equal(osrc.arrays["eps_c"], rsrc.eps_c) or (osrc.arrays["eps_c"] != rsrc.eps_c).sum() <= 2
    assert cases.relerr(osrc.arrays["approx_du"], rsrc.approx_du) <= 1e-8


def test_advection_with_flyer_hyperviscosity():
    """BASELINE configs[0]: linear advection a = (1, 0.5), PHS r^5, Flyer hyperviscosity k = 2, inlet Dirichlet"""
    s = cases.fixture_setup(p=5, N=3)
    ops = orc.compute_flux_operator(s["points"], s["nb"], 5, 3)
    ops4 = orc.compute_flux_operator(s["points"], s["nb"], 5, 3, 4)
    H = (ops4[0] + ops4[1]).tocsc()
    ic = cases.ic_bump_advection
    spec = dict(inlet="dirichlet", outlet="nothing", top="nothing", bottom="nothing", cyl="nothing")
    osrc = orc.source_hyperviscosity_flyer(s["points"], s["nb"], 5, 3, s["dx_min"], k=2, c=1.0)
    P = orc.OracleProblem(s["points"], 1, orc.EQ_ADVECTION2D, [1.0, 0.5], ops[0], ops[1], cases.oracle_bcs(s, spec, ic), [osrc])
    R = ref.RefProblem(s["points"], ref.Advection2D((1.0, 0.5)), ops[0], ops[1], _ref_bcs(s, spec, ic),
                       [ref.Hyperviscosity(H, s["dx_min"] ** 4)])
    u0 = ic(s["points"], 0.0)
    assert cases.relerr(P.rhs(u0.copy(), 0.0), R.rhs(u0.copy(), 0.0)) <= 1e-12
    dt = 0.1 * s["dx_min"]
    ua, _ = P.solve_ssprk33(u0, 0.0, dt, 20)
    ub, _ = R.solve_ssprk33(u0, 0.0, dt, 20)
    assert cases.relerr(ua, ub) <= 1e-11


def test_batched_setup_of_the_cpu_bench_arm():
    """oracle/bench_setup.py (setup of `bench.py --impl reference`): the numpy Hilbert numbering equals the product's
    mft_sfc_order, and the batched weights agree with the oracle's point-by-point compute_flux_operator to rounding"""
    import bench_setup as bs
    import mft_oracle as orc

    cm = bs.cloud_module()
    cl = cm.jittered_lattice(40, 32, 10.0, 8.0, seed=0)
    perm = bs.hilbert_order(cl.points)
    assert sorted(perm.tolist()) == list(range(len(cl.points)))
    try:
        import mft_b200 as m

        assert np.array_equal(perm, m._lib.sfc_order(cl.points))
    except OSError:
        pass   # (the product library is not needed for this test)
    cl = cm.reorder(cl, perm)
    nb, dx_min, dx_avg = orc.point_data(cl.points, 20)
    fast = bs.flux_operator_batched(cl.points, nb, 3, 3)
    ref = orc.compute_flux_operator(cl.points, nb, 3, 3)
    for a, b in zip(fast, ref):
        assert a.nnz == b.nnz and np.array_equal(a.indices, b.indices) and np.array_equal(a.indptr, b.indptr)
        assert np.abs(a.data - b.data).max() <= 1e-11 * np.abs(b.data).max()
