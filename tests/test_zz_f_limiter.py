"""Zhang-Shu positivity limiter (SURVEY.md section 8 row f4;
reference: src/callbacks_stage/positivity_zhang_shu_point2d.jl:22-82, positivity_zhang_shu.jl:29-72).
CPU: the oracle's C restatement against a plain-Python restatement and the limiter's defining properties.
GPU: the device kernels against the oracle (written after this round's GPU budget was spent: first hardware run is the
round-end test pass)."""
import numpy as np
import pytest

import cases
from cases import orc


def _state(fx):
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    u[0, 100:130] *= 0.01          # a pocket of very low density
    u[3, 500:520] *= 0.2           # a pocket of low (even negative) pressure
    u[3, 900] = 0.01
    return u


def _python_restatement(u, nbrs, thresholds, variables, gamma):
    def var(v, w):
        return w[0] if v == orc.VAR_DENSITY else (gamma - 1.0) * (w[3] - 0.5 * (w[1] * w[1] + w[2] * w[2]) / w[0])

    w = u.copy()
    for thr, v in zip(thresholds, variables):
        loc = np.zeros_like(w)
        for e in range(w.shape[1]):
            nb = nbrs[e]
            vmin = min(var(v, w[:, i]) for i in nb)
            if not vmin < thr:
                continue
            mean = np.zeros(4)
            for i in nb:
                mean = mean + w[:, i]
            mean = mean / len(nb)
            vm = var(v, mean)
            th = (vm - thr) / (vm - vmin)
            loc[:, e] = th * w[:, e] + (1 - th) * mean
        m = (loc != 0).any(axis=0)
        w[:, m] = loc[:, m]
    return w


def test_oracle_limiter_against_python_restatement_and_properties():
    fx = cases.fixture_setup(p=3, N=3)
    u0 = _state(fx)
    thr, var = (0.05, 0.02), (orc.VAR_DENSITY, orc.VAR_PRESSURE)
    u = orc.limiter_zhang_shu(u0.copy(), fx["nb"], thr, var, cases.GAMMA)
    w = _python_restatement(u0, fx["nb"], thr, var, cases.GAMMA)
    assert np.abs(w - u).max() <= 1e-14 * np.abs(u).max()          # fma vs separate rounding only
    changed = (u != u0).any(axis=0)
    assert 50 < changed.sum() < 400                                 # only stencils that see a bad value are touched
    # untouched rows are bit-identical; a second application with the same thresholds leaves density >= threshold rows alone
    assert np.array_equal(u[:, ~changed], u0[:, ~changed])
    assert u[0].min() > 0.0 and u0[0].min() < 0.05
    # nothing to do -> identity
    v = cases.ic_smooth_euler(fx["points"], 0.0)
    assert np.array_equal(orc.limiter_zhang_shu(v.copy(), fx["nb"], (1e-6,), (orc.VAR_DENSITY,), cases.GAMMA), v)


def test_emulated_device_kernels_match_oracle_bit_for_bit():
    """the product's kernel thread bodies (csrc/mft_limiter_kernels.cuh) run on the host by tests/emu, device data layout"""
    import emu

    fx = cases.fixture_setup(p=3, N=3)
    for thr, var in (((0.05, 0.02), (orc.VAR_DENSITY, orc.VAR_PRESSURE)), ((0.9,), (orc.VAR_DENSITY,)),
                     ((1e-6,), (orc.VAR_PRESSURE,))):
        u0 = _state(fx)
        ref = orc.limiter_zhang_shu(u0.copy(), fx["nb"], thr, var, cases.GAMMA)
        got = emu.limiter_zhang_shu(u0.copy(), fx["nb"], thr, var, cases.GAMMA)
        assert np.array_equal(got, ref, equal_nan=True)
    u0 = _state(fx)
    u0[0, 40] = np.nan                                              # NaN propagates through Julia's min: stencils of 40 are limited
    ref = orc.limiter_zhang_shu(u0.copy(), fx["nb"], (0.05,), (orc.VAR_DENSITY,), cases.GAMMA)
    got = emu.limiter_zhang_shu(u0.copy(), fx["nb"], (0.05,), (orc.VAR_DENSITY,), cases.GAMMA)
    assert np.array_equal(got, ref, equal_nan=True)


@pytest.mark.gpu
def test_device_limiter_matches_oracle():
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    for reorder in ("hilbert", None):
        basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=fx["nv"])
        solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(reorder=reorder))
        domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
        eq = m.CompressibleEulerEquations2D(cases.GAMMA)
        semi = m.SemidiscretizationHyperbolic(domain, eq, cases.ic_smooth_euler, solver,
                                              boundary_conditions=dict(inlet=m.BoundaryConditionDoNothing()), operators=ops)
        lim = m.PositivityPreservingLimiterZhangShu(thresholds=(0.05, 0.02), variables=(m.density, m.pressure))
        u0 = _state(fx)
        ref = orc.limiter_zhang_shu(u0.copy(), fx["nb"], (0.05, 0.02), (orc.VAR_DENSITY, orc.VAR_PRESSURE), cases.GAMMA)
        u = lim(u0.copy(), semi)
        assert np.array_equal(u, ref), np.abs(u - ref).max()
        semi.close()


@pytest.mark.gpu
def test_stage_limiter_inside_ssprk33_matches_oracle():
    """SSPRK33(stage_limiter!): limiter after every stage update, on the device, inside the (graph-replayed) step"""
    import ctypes as C

    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=fx["nv"])
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA())
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = cases.ic_smooth_euler
    bcs = cases.DIVERGENCE_TEST_BCS
    kinds = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic), slip=lambda: m.boundary_condition_slip_wall,
                 nothing=lambda: m.BoundaryConditionDoNothing())
    srcs = m.SourceTerms(rv=m.SourceUpwindViscosityTominec(solver, eq, domain))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions={k: kinds[v]() for k, v in bcs.items()},
                                          source_terms=srcs, operators=ops)
    thr, var = (0.9,), (orc.VAR_DENSITY,)     # the smooth state dips to rho = 0.8: the limiter is active every stage
    lim = m.PositivityPreservingLimiterZhangShu(thresholds=thr, variables=(m.density,))
    dt, nsteps = 0.1 * fx["dx_min"] / 3.0, 5
    ode = m.semidiscretize(semi, (0.0, nsteps * dt))
    sol = m.solve(ode, m.SSPRK33(stage_limiter=lim), dt=dt, nsteps=nsteps)
    # oracle: the same Shu-Osher stages with the limiter after each stage update
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, bcs, ic),
                          [orc.source_upwind(fx["dx_avg"])])
    lib = orc.lib()
    u = ode.u0.copy()
    k = P.rhs(u, 0.0)
    for _ in range(nsteps):
        uprev = u.copy()
        for s in (1, 2, 3):
            lib.orc_ssprk33_stage(C.c_int64(u.size), s, C.c_double(dt), C.c_void_p(uprev.ctypes.data),
                                  C.c_void_p(k.ctypes.data), C.c_void_p(u.ctypes.data))
            orc.limiter_zhang_shu(u, fx["nb"], thr, var, cases.GAMMA)
            k = P.rhs(u, 0.0)
    assert cases.relerr(sol.u, u) <= 1e-9
    plain = m.solve(ode, m.SSPRK33(), dt=dt, nsteps=nsteps)       # and the limiter really changed the trajectory
    assert cases.relerr(plain.u, u) > 1e-6
    semi.close()
