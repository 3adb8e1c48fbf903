"""Row f1 (device setup pipeline), GPU tier: mft_setup_knn / mft_setup_rbf_weights through the C ABI against the oracle,
brute force and the host emulation of the same kernel source.  (Written after this round's GPU budget was spent: first
hardware run is the round-end test pass; the kernel bodies and the orchestration are covered on the CPU tier by
tests/test_emu_setup.py.)"""
import numpy as np
import pytest

import cases
import emu
from cases import orc
from test_emu_setup import adversarial_clouds, brute_knn

pytestmark = pytest.mark.gpu


def _m():
    import mft_b200

    return mft_b200


def _knn_dev(pts, k):
    m = _m()
    L = m._lib
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    n = len(pts)
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    nb = np.empty((n, k), np.int64)
    d = np.empty((n, k))
    L.check(L.load().mft_setup_knn(0, n, L.ptr(x), L.ptr(y), k, L.ptr(nb), L.ptr(d)))
    return nb - 1, d


@pytest.mark.parametrize("name", sorted(adversarial_clouds()))
def test_device_knn_bit_exact_vs_brute_force(name):
    pts = np.ascontiguousarray(adversarial_clouds()[name])
    for k in (1, 20, 42):
        if k > len(pts):
            continue
        nb, d = _knn_dev(pts, k)
        rb, rd = brute_knn(pts, k)
        assert np.array_equal(nb, rb), (name, k)
        assert np.array_equal(d, rd), (name, k)


def test_device_knn_matches_the_oracle_on_the_fixture():
    m = _m()
    fx = cases.fixture_setup()
    nb, dx_min, dx_avg = m.setup_ops.knn_device(fx["points"], fx["nv"])
    assert np.array_equal(nb, fx["nb"]) and dx_min == fx["dx_min"] and dx_avg == fx["dx_avg"]


@pytest.mark.parametrize("p,N,k", [(3, 3, None), (5, 3, 2), (5, 3, 4), (7, 5, None)])
def test_device_weights_match_oracle_and_emulation(p, N, k):
    m = _m()
    s = cases.fixture_setup(p=p, N=N)
    wx, wy = m.setup_ops.rbf_fd_weights_device(s["points"], s["nb"], p, N, k)
    ref = orc.compute_flux_operator(s["points"], s["nb"], p, N, k)
    ex, ey = emu.setup_rbf_weights(s["points"], s["nb"], p, N, k or 1)
    for w, e, B in zip((wx, wy), (ex, ey), ref):
        A = m.setup_ops.assemble_csc(s["nb"], w)
        assert np.array_equal(A.indices, B.indices)
        tol = 1e-8 if (k or 1) <= 2 else 1e-5
        assert np.abs(A.data - B.data).max() <= tol * np.abs(B.data).max()
        # same source, same arithmetic (no FMA contraction, IEEE sqrt / division): expected bit-identical; the bar
        # asserted is 1e-12 of the row scale
        assert np.abs(w - e).max() <= 1e-12 * np.abs(e).max()


def test_device_setup_feeds_rhs():
    """RBFFDEngineCUDA(setup="device"): domain + operators from the GPU pipeline, rhs! against the oracle that is handed
    the same operators (1e-12), and against the oracle's own operators (1e-9: weights differ by rounding)."""
    m = _m()
    fx = cases.fixture_setup(p=3, N=3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(setup="device"))
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    assert np.array_equal(domain.pd.neighbors, fx["nb"]) and domain.pd.dx_min == fx["dx_min"]
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    ic = cases.ic_smooth_euler
    kinds = dict(dirichlet=lambda: m.BoundaryConditionDirichlet(ic), slip=lambda: m.boundary_condition_slip_wall,
                 nothing=lambda: m.BoundaryConditionDoNothing())
    bcs = cases.DIVERGENCE_TEST_BCS
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions={k: kinds[v]() for k, v in bcs.items()},
                                          source_terms=m.SourceTerms(uw=m.SourceUpwindViscosityTominec(solver, eq, domain)))
    ops = semi.cache.rbf_differentiation_matrices
    ode = m.semidiscretize(semi, (0.0, 1.0))
    u = ode.u0.copy()
    du = np.empty_like(u)
    m.rhs_(du, u, semi, 0.0)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], cases.oracle_bcs(fx, bcs, ic),
                          [orc.source_upwind(fx["dx_avg"])])
    assert cases.relerr(du, P.rhs(ode.u0.copy(), 0.0)) <= 1e-12
    oops = orc.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    P2 = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], oops[0], oops[1], cases.oracle_bcs(fx, bcs, ic),
                           [orc.source_upwind(fx["dx_avg"])])
    assert cases.relerr(du, P2.rhs(ode.u0.copy(), 0.0)) <= 1e-9
    semi.close()


def test_device_setup_at_one_million_points():
    """BASELINE configs[1] size: structural properties of the neighbour tables, a sampled comparison with a KD-tree, and
    polynomial reproduction of the weights"""
    from scipy.spatial import cKDTree

    m = _m()
    cl = m.cloud.jittered_lattice(1024, 1024, 10.0, 10.0, seed=0).points
    n = len(cl)
    nb, d = _knn_dev(cl, 20)
    assert np.array_equal(nb[:, 0], np.arange(n)) and (d[:, 0] == 0).all()
    assert (np.diff(d, axis=1) >= 0).all() and nb.min() >= 0 and nb.max() < n
    sample = np.random.default_rng(0).choice(n, 20000, replace=False)
    td, ti = cKDTree(cl).query(cl[sample], k=24)
    ti, td = m.setup_ops.canonical_ties(ti.astype(np.int64), td)
    assert np.array_equal(nb[sample], ti[:, :20]) and np.array_equal(d[sample], td[:, :20])
    wx, wy = m.setup_ops.rbf_fd_weights_device(cl, nb, 3, 3)
    X, Y = cl[nb, 0], cl[nb, 1]
    scale = np.abs(wx).sum(axis=1).max()
    assert np.abs(wx.sum(axis=1)).max() <= 1e-9 * scale and np.abs(wy.sum(axis=1)).max() <= 1e-9 * scale
    assert np.abs((wx * X).sum(axis=1) - 1.0).max() <= 1e-7 and np.abs((wy * Y).sum(axis=1) - 1.0).max() <= 1e-7
    # and the host mirror (batched LAPACK LU) agrees on a slice
    hx, hy = m.setup_ops.rbf_fd_weights(cl, nb[:50000], 3, 3)
    assert np.abs(wx[:50000] - hx).max() <= 1e-8 * np.abs(hx).max() and np.abs(wy[:50000] - hy).max() <= 1e-8 * np.abs(hy).max()


def test_device_knn_queries_and_weight_rows():
    """the per-rank entry points (mft_setup_knn_queries / mft_setup_rbf_weights_rows) against the whole-cloud calls"""
    m = _m()
    rng = np.random.default_rng(11)
    pts = rng.random((4000, 2)) * [2.0, 1.0]
    full_nb, full_d = _knn_dev(pts, 20)
    q = rng.integers(0, len(pts), 700)
    nb, d = m.setup_ops.knn_queries_device(pts, q, 20)
    assert np.array_equal(nb, full_nb[q]) and np.array_equal(d, full_d[q])
    enb, ed = emu.setup_knn(pts, 20, queries=q)
    assert np.array_equal(nb, enb) and np.array_equal(d, ed)
    wx, wy = m.setup_ops.rbf_fd_weights_device(pts, full_nb, 3, 3)
    rx, ry = m.setup_ops.rbf_fd_weights_rows_device(pts, full_nb[q], 3, 3)
    assert np.array_equal(rx, wx[q]) and np.array_equal(ry, wy[q])
    with pytest.raises(m._lib.MftError, match="query index"):
        m.setup_ops.knn_queries_device(pts, np.array([4000]), 20)


def test_device_hybrid_gaussian_phs_weights():
    """mft_setup_rbf_weights_hybrid against the oracle and the host emulation (exp differs by an ulp between libm and CUDA:
    1e-10 of the row scale instead of 1e-12)"""
    m = _m()
    hyb = (0.5, 2.0, 0.7)
    s = cases.fixture_setup(p=5, N=3)
    wx, wy = m.setup_ops.rbf_fd_weights_rows_device(s["points"], s["nb"], 5, 3, None, 0, hyb)
    ref = orc.compute_flux_operator(s["points"], s["nb"], 5, 3, None, hybrid=hyb)
    ex, ey = emu.setup_rbf_weights(s["points"], s["nb"], 5, 3, 1, hybrid=hyb)
    for w, e, B in zip((wx, wy), (ex, ey), ref):
        A = m.setup_ops.assemble_csc(s["nb"], w)
        assert np.abs(A.data - B.data).max() <= 1e-8 * np.abs(B.data).max()
        assert np.abs(w - e).max() <= 1e-10 * np.abs(e).max()
