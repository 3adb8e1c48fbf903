"""The fused device-resident step (MFT_OPT_FUSED_STEP, csrc/mft_fused_kernels.cuh): ONE kernel per stage for BC pass 2 + SSPRK33
stage update + BC pass 1 + ode_mean / ode_maximum, against the separate kernels of round 1 (same arithmetic, same summation
trees => the states must agree BIT FOR BIT) and against the oracle (1e-9 after N steps, north_star tolerance).

The one-pass ode_maximum statistic (16 lexicographic "leaves") is exercised on states built to tie: constant density with a few
momentum levels (every level of the lexicographic order is decided among exact ties), duplicated extreme points, a uniform state,
Sod data (two constant plateaus).  With dt = 0 the stages reproduce the uploaded state exactly, so the norms each stage computes
are the norms of that state; pass A re-checks every row against them (MFT_FIELD_NORM_MISSES must stay 0)."""
import ctypes as C

import numpy as np
import pytest

import cases
from cases import orc

pytestmark = pytest.mark.gpu

NAMES = dict(left=1, right=2, bottom=3, top=4)


def _semi(m, cl, ic, bc_kinds, source, fused, **eng):
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(fused_step=fused, **eng))
    domain = m.PointCloudDomain(solver, cl, NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    mk = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic), slip=lambda: m.boundary_condition_slip_wall,
              nothing=lambda: m.BoundaryConditionDoNothing())
    if source == "residual":
        srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, c_rv=1.0, c_uw=1.0, polydeg=3))
    else:
        srcs = m.SourceTerms(uw=m.SourceUpwindViscosityTominec(solver, eq, domain, c_uw=1.0))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions={k: mk[v]() for k, v in bc_kinds.items()},
                                          source_terms=srcs)
    return semi, domain


def _run(m, semi, u0, dt, nsteps, residual):
    """device-resident steps through the C ABI; returns the final state, the norms of the last rhs! and the miss counter"""
    L = m._lib
    lib = m.load()
    ctx = semi.ctx
    L.check(lib.mft_upload_state(ctx, L.soa_ptrs(u0)))
    if residual:
        L.check(lib.mft_history_push(ctx, 0.0, 0, 3))
    t = 0.0
    for i in range(nsteps):
        L.check(lib.mft_ssprk_step(ctx, L.SSPRK33, t, dt))
        t += dt
        if residual and dt > 0.0:   # (dt = 0: the history times would coincide; success_iter stays 0)
            L.check(lib.mft_history_push(ctx, t, i + 1, 3))
    u = np.empty_like(u0)
    L.check(lib.mft_download_state(ctx, L.soa_ptrs(u)))
    du = np.empty_like(u0)
    L.check(lib.mft_download_du(ctx, L.soa_ptrs(du)))
    norms, miss = np.zeros(4), np.zeros(1)
    if residual:
        L.check(lib.mft_get_field(ctx, L.FIELD_NORMS, L.ptr(norms)))
    L.check(lib.mft_get_field(ctx, L.FIELD_NORM_MISSES, L.ptr(miss)))
    return u, du, norms, int(miss[0])


def _both(m, cl, ic, bc_kinds, source, u0, dt, nsteps, **eng):
    out = []
    for fused in (True, False):
        semi, domain = _semi(m, cl, ic, bc_kinds, source, fused, **eng)
        out.append(_run(m, semi, u0, dt, nsteps, source == "residual"))
        pd = domain.pd
        semi.close()
    (uf, duf, nf, mf), (uc, duc, nc, _) = out
    assert mf == 0, f"{mf} rows exceeded the one-pass norms"
    assert np.array_equal(uf, uc), f"fused step differs from the separate kernels: {np.abs(uf - uc).max():.3e}"
    assert np.array_equal(duf, duc)
    assert np.array_equal(nf, nc), (nf, nc)
    return uf, pd


ALL_DIRICHLET = dict(left="dirichlet", right="dirichlet", bottom="dirichlet", top="dirichlet")


@pytest.mark.parametrize("source", ["residual", "upwind"])
def test_fused_equals_separate_kernels_and_oracle_vortex(source):
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(96, 80, 10.0, 10.0 * 80 / 96, seed=1)
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, cases.GAMMA, center=(5.0, 4.0))   # noqa: E731
    u0 = np.ascontiguousarray(ic(cl.points, 0.0))
    semi, domain = _semi(m, cl, ic, ALL_DIRICHLET, source, True)
    pd = domain.pd
    dt, nsteps = 0.1 * pd.dx_min / 8.0, 12
    ops = semi.cache.rbf_differentiation_matrices
    obc = [orc.OracleBC(orc.BC_DIRICHLET, domain.boundary_tags[k].idx, domain.boundary_tags[k].normals, value_fn=lambda x, t: ic(x, t))
           for k in NAMES]
    src = orc.source_residual(pd.dx_avg, polydeg=3) if source == "residual" else orc.source_upwind(pd.dx_avg)
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], obc, [src])
    ur, _ = P.solve_ssprk33(u0, 0.0, dt, nsteps, approx_order=3 if source == "residual" else None)
    semi.close()
    uf, _ = _both(m, cl, ic, ALL_DIRICHLET, source, u0, dt, nsteps)
    assert cases.relerr(uf, ur) <= 1e-9


@pytest.mark.parametrize("graph", [1, 0])
def test_fused_sod_slip_walls(graph):
    """configs[3] physics: discontinuous data (two plateaus of identical states), slip walls (BC pass 2 changes u and du)"""
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(96, 48, 2.0, 1.0, seed=2)
    ic = lambda x, t, e=None: m.cloud.sod(x, cases.GAMMA, x_mid=1.0)   # noqa: E731
    kinds = dict(left="dirichlet", right="dirichlet", bottom="slip", top="slip")
    u0 = np.ascontiguousarray(ic(cl.points, 0.0))
    semi, domain = _semi(m, cl, ic, kinds, "residual", True)
    pd = domain.pd
    dt, nsteps = 0.1 * pd.dx_min / 3.0, 15
    ops = semi.cache.rbf_differentiation_matrices
    okind = dict(dirichlet=orc.BC_DIRICHLET, slip=orc.BC_SLIP_WALL)
    obc = [orc.OracleBC(okind[v], domain.boundary_tags[k].idx, domain.boundary_tags[k].normals,
                        value_fn=(lambda x, t: ic(x, t)) if v == "dirichlet" else None) for k, v in kinds.items()]
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], obc, [orc.source_residual(pd.dx_avg, polydeg=3)])
    ur, _ = P.solve_ssprk33(u0, 0.0, dt, nsteps, approx_order=3)
    semi.close()
    uf, _ = _both(m, cl, ic, kinds, "residual", u0, dt, nsteps, cuda_graph=graph)
    assert cases.relerr(uf, ur) <= 1e-9


def _tie_states(pts, rng):
    n = len(pts)
    X, Y = pts[:, 0], pts[:, 1]
    E = 30.0 + np.sin(3.1 * X) * np.cos(2.3 * Y)
    out = {}
    # constant density, momenta on a few exact levels: every key of the lexicographic order is decided among ties
    out["levels"] = np.stack([np.ones(n), np.round(2.0 * np.sin(X)) / 2.0, np.round(2.0 * np.cos(1.7 * Y)) / 2.0, E])
    # the extreme-density point duplicated with different momenta
    u = np.stack([1.0 + 0.1 * np.sin(X + Y), 0.3 * np.cos(X), 0.2 * np.sin(Y), E])
    hi, lo = np.argsort(u[0])[-1], np.argsort(u[0])[0]
    for k, j in enumerate(rng.choice(n, 40, replace=False)):
        u[:, j] = u[:, hi if k % 2 else lo]
        u[1 + k % 3, j] += 0.01 * (k - 20)
    out["duplicates"] = u
    # a completely uniform state (all norms 0 -> eps), and one with a single perturbed point
    out["uniform"] = np.stack([np.full(n, 1.0), np.full(n, 0.5), np.full(n, -0.25), np.full(n, 20.0)])
    v = out["uniform"].copy()
    v[:, n // 3] = [1.0, 0.5, -0.25, 20.5]
    out["one_point"] = v
    return out


@pytest.mark.parametrize("lex,vn", [(True, True), (True, False), (False, True)])
def test_one_pass_norms_on_tied_states(lex, vn):
    """dt = 0: every stage recomputes the norms of the uploaded state (after the BC pass) with the one-pass statistic; pass A
    counts the rows that exceed them; the separate kernels (two passes over u) must give the same bits"""
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(80, 72, 10.0, 9.0, seed=7)
    rng = np.random.default_rng(5)
    kinds = dict(left="nothing", right="dirichlet", bottom="slip", top="slip")
    for name, u0 in _tie_states(cl.points, rng).items():
        u0 = np.ascontiguousarray(u0)
        ic = lambda x, t, e=None, u0=u0: _table(cl, u0, x)   # noqa: E731
        for dt in (0.0, 1e-4):
            _both(m, cl, ic, kinds, "residual", u0, dt, 2, max_lexicographic=lex, mean_divisor_vn=vn)


def test_adjacent_densities_that_tie_after_rounding():
    """The 2-GPU bench cloud's first stage (profiles/r2y_norm_miss_diagnostic_2ranks.log): 17 points share the largest density T,
    a few of them sit one ulp above / below it, and with the V*N divisor of ode_mean the deviations |T - m0| and |T +- ulp - m0|
    round to the same double for half of the means -- the order is then decided on |m1 - mean| among BOTH density groups.  The
    statistic keeps a second density level for exactly this; the means' last bits are swept by nudging one far-away row."""
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(80, 72, 10.0, 9.0, seed=7)
    n = len(cl.points)
    X, Y = cl.points[:, 0], cl.points[:, 1]
    rng = np.random.default_rng(11)
    kinds = dict(left="nothing", right="nothing", bottom="nothing", top="nothing")
    T = 1.0 - 5 * 2.0 ** -53
    base = np.stack([0.5 + 0.3 * np.sin(X) * np.cos(Y), 1.0 + 0.01 * np.cos(X), 0.01 * np.sin(Y), 30.0 + np.sin(3.1 * X)])
    tie = rng.choice(n, 17, replace=False)
    base[0, tie] = T
    base[1, tie] = 1.0 + 5e-7 * np.linspace(-1.0, 1.0, 17)
    far = int(np.argsort(base[0])[n // 2])                        # a row in the middle of the density range: only moves the mean
    ic = lambda x, t, e=None: _table(cl, base, x)                 # noqa: E731  (no Dirichlet boundary: never evaluated on the device)
    semi_f, _ = _semi(m, cl, ic, kinds, "residual", True)
    semi_c, _ = _semi(m, cl, ic, kinds, "residual", False)
    decided_on_m1 = 0
    for up in (np.nextafter(T, 2.0), np.nextafter(T, 0.0)):
        for k in range(64):
            u0 = base.copy()
            u0[0, tie[[2, 9]]] = up                               # not the rows with the extreme momenta
            u0[0, far] += k * 4.0 * n * 2.0 ** -56                # shifts ode_mean(rho) = sum / (4 n) by about k * 2^-56
            u0 = np.ascontiguousarray(u0)
            uf, duf, nf, mf = _run(m, semi_f, u0, 0.0, 1, True)
            uc, duc, nc, _ = _run(m, semi_c, u0, 0.0, 1, True)
            assert mf == 0, f"{mf} rows exceeded the one-pass norms (k = {k})"
            assert np.array_equal(nf, nc), (k, nf, nc)
            assert np.array_equal(uf, uc) and np.array_equal(duf, duc)
            # did the two density groups tie after rounding?  then |m1 - mean| of the norms belongs to a row of the OTHER group
            dev1 = np.abs(u0[1, tie] - u0[1].sum() / (4.0 * n))
            j = int(np.argmin(np.abs(dev1 - nc[1])))
            top = up if up > T else T
            decided_on_m1 += int(u0[0, tie[j]] != top)
    semi_f.close()
    semi_c.close()
    assert decided_on_m1 > 0, "none of the swept means made the two densities tie: the sweep does not exercise the second level"


def _table(cl, u0, x):
    """Dirichlet data = the state itself at the queried boundary points (nearest point of the cloud)"""
    from scipy.spatial import cKDTree

    _, j = cKDTree(cl.points).query(x)
    return np.ascontiguousarray(u0[:, j])


def test_fused_on_the_fixture_cloud():
    """the reference's own point cloud: inlet Dirichlet, outlet do-nothing, slip walls on top / bottom / cylinder"""
    import mft_b200 as m

    kinds = dict(inlet="dirichlet", outlet="nothing", bottom="slip", top="slip", cyl="slip")
    names = cases.BOUNDARY_NAMES
    out = []
    for fused in (True, False):
        basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
        solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(fused_step=fused))
        domain = m.PointCloudDomain(solver, cases.FIXTURE, names)
        eq = m.CompressibleEulerEquations2D(cases.GAMMA)
        mk = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(cases.ic_smooth_euler), slip=lambda: m.boundary_condition_slip_wall,
                  nothing=lambda: m.BoundaryConditionDoNothing())
        srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3))
        semi = m.SemidiscretizationHyperbolic(domain, eq, cases.ic_smooth_euler, solver,
                                              boundary_conditions={k: mk[v]() for k, v in kinds.items()}, source_terms=srcs)
        u0 = m.semidiscretize(semi, (0.0, 1.0)).u0
        out.append(_run(m, semi, u0, 0.1 * domain.pd.dx_min / 3.0, 10, True))
        semi.close()
    assert out[0][3] == 0
    for a, b in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(a, b)


def test_a_rounding_tie_is_reported_loudly():
    """The documented limit of the one-pass statistic (tests/test_norm_leaves_cpu.py): a plateau that ties EXACTLY on
    the density, carries rounding noise in m1 while the mean of m1 is O(1) (so |m1 - mean| rounds to one double for the whole
    plateau) and varies materially in m2.  The reference's lexicographic maximum is then decided on m2 among ALL plateau points;
    the leaves only hold the points with extreme m1.  Pass A checks every row against the norms it was given, so the fused step
    must not return silently: mft_synchronize reports MFT_ENORMS; the separate kernels (two passes) are unaffected."""
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(64, 48, 8.0, 6.0, seed=3)
    pts = cl.points
    rng = np.random.default_rng(9)
    n = len(pts)
    plateau = pts[:, 0] < 4.0
    u0 = np.stack([np.where(plateau, 2.0, 1.0), np.where(plateau, 1e-20 * rng.standard_normal(n), 1.0),
                   np.where(plateau, np.sin(2.0 * pts[:, 1]), 0.0), np.full(n, 30.0)])
    u0 = np.ascontiguousarray(u0)
    kinds = dict(left="nothing", right="nothing", bottom="nothing", top="nothing")
    ic = lambda x, t, e=None: _table(cl, u0, x)   # noqa: E731
    L = m._lib
    lib = m.load()
    norms = {}
    for fused in (True, False):
        semi, _ = _semi(m, cl, ic, kinds, "residual", fused)
        L.check(lib.mft_upload_state(semi.ctx, L.soa_ptrs(u0)))
        L.check(lib.mft_history_push(semi.ctx, 0.0, 0, 3))
        L.check(lib.mft_ssprk_step(semi.ctx, L.SSPRK33, 0.0, 0.0))      # dt = 0: every stage sees exactly u0
        rc = lib.mft_synchronize(semi.ctx)
        out = np.zeros(4)
        L.check(lib.mft_get_field(semi.ctx, L.FIELD_NORMS, L.ptr(out)))
        norms[fused] = out
        if fused:
            assert rc == -6, rc                                          # MFT_ENORMS
            assert b"one-pass" in lib.mft_last_error()
            miss = np.zeros(1)
            L.check(lib.mft_get_field(semi.ctx, L.FIELD_NORM_MISSES, L.ptr(miss)))
            assert miss[0] > 0
            assert lib.mft_synchronize(semi.ctx) == 0                    # reported once per occurrence
        else:
            assert rc == 0
        semi.close()
    # the first two keys agree (same maximal |rho - mean| and the rounded |m1 - mean|); the third is where the tie is decided
    assert np.array_equal(norms[True][:2], norms[False][:2]) and norms[True][2] < norms[False][2]
