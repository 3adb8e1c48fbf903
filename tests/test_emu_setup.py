"""Row f1 (device setup pipeline), CPU tier: the product's kernel thread bodies and host orchestration
(csrc/mft_setup_kernels.cuh, csrc/mft_setup_host.inl), compiled for the host by tests/emu and checked against the oracle
and brute force.  The GPU tier (tests/test_zz_g_setup_gpu.py) runs the same source as CUDA kernels through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import cases
import emu
from cases import orc


def brute_knn(pts, k):
    """all-pairs reference: ascending (distance, index)"""
    dx = pts[:, None, 0] - pts[None, :, 0]
    dy = pts[:, None, 1] - pts[None, :, 1]
    d = np.sqrt(dx * dx + dy * dy)
    idx = np.broadcast_to(np.arange(len(pts)), d.shape)
    o = np.lexsort((idx, d), axis=1)[:, :k]
    return o, np.take_along_axis(d, o, 1)


def adversarial_clouds():
    rng = np.random.default_rng(5)
    out = {}
    gx, gy = np.meshgrid(np.arange(40.0), np.arange(25.0))
    out["lattice"] = np.stack([gx.ravel(), gy.ravel()], 1) * 0.1          # exact distance ties everywhere
    out["uniform"] = rng.random((3000, 2)) * [3.0, 1.0]
    c = rng.normal(size=(2500, 2)) * 0.01
    c[:500] += [5, 5]
    c[500:1000] *= 100
    out["clustered"] = c                                                    # cell occupancy from 0 to hundreds
    out["thin"] = np.stack([rng.random(2000), rng.random(2000) * 1e-9], 1)  # extreme aspect ratio
    out["line"] = np.stack([rng.random(500), np.zeros(500)], 1)             # zero extent in y
    d = rng.random((1000, 2))
    d[::7] = d[3]
    out["duplicates"] = d                                                   # coincident points (distance 0 ties)
    out["tiny"] = rng.random((20, 2))                                       # k == n
    out["offset"] = rng.random((2000, 2)) * 1e-3 + [1e5, -3e4]              # far from the origin
    out["single_cell"] = np.zeros((30, 2)) + rng.random((30, 2)) * 1e-300   # degenerate extent
    return out


@pytest.mark.parametrize("name", sorted(adversarial_clouds()))
def test_knn_bit_exact_vs_brute_force(name):
    pts = np.ascontiguousarray(adversarial_clouds()[name])
    for k in (1, 7, 20, 42):
        if k > len(pts):
            continue
        nb, d = emu.setup_knn(pts, k)
        rb, rd = brute_knn(pts, k)
        assert np.array_equal(nb, rb), (name, k)
        assert np.array_equal(d, rd), (name, k)


def test_knn_matches_the_oracle_on_the_fixture_and_a_synthetic_cloud():
    fx = cases.fixture_setup()
    nb, d = emu.setup_knn(fx["points"], fx["nv"])
    assert np.array_equal(nb, fx["nb"])
    assert d[:, 1].min() == fx["dx_min"] and d[:, 1].mean() == fx["dx_avg"]
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(256, 128, 10.0, 5.0, seed=0).points
    nb, d = emu.setup_knn(cl, 20)
    onb, omin, oavg = orc.point_data(cl, 20)
    assert np.array_equal(nb, onb) and d[:, 1].min() == omin and d[:, 1].mean() == oavg


def test_knn_argument_errors():
    pts = np.random.default_rng(0).random((10, 2))
    with pytest.raises(emu.EmuError, match="exceeds the number of points"):
        emu.setup_knn(pts, 11)
    with pytest.raises(emu.EmuError, match="outside 1"):
        emu.setup_knn(np.random.default_rng(0).random((100, 2)), 65)
    bad = pts.copy()
    bad[3, 1] = np.nan
    with pytest.raises(emu.EmuError, match="non-finite"):
        emu.setup_knn(bad, 3)


@pytest.mark.parametrize("p,N,k", [(3, 3, None), (5, 3, None), (5, 3, 2), (5, 3, 4), (3, 2, None), (5, 4, None), (7, 5, None)])
def test_weights_match_oracle(p, N, k):
    """one-thread-per-point LU (product kernel body) vs the oracle's per-point Bunch-Kaufman solve: 1e-8 of the row scale
    (same bar as the host mirror, tests/test_setup_cpu.py); 1e-5 for 4th derivatives"""
    s = cases.fixture_setup(p=p, N=N)
    import mft_b200 as m

    wx, wy = emu.setup_rbf_weights(s["points"], s["nb"], p, N, k or 1)
    ref = orc.compute_flux_operator(s["points"], s["nb"], p, N, k)
    for w, B in zip((wx, wy), ref):
        A = m.setup_ops.assemble_csc(s["nb"], w)
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        tol = 1e-8 if (k or 1) <= 2 else 1e-5
        assert np.abs(A.data - B.data).max() <= tol * np.abs(B.data).max()


def test_weights_reproduce_polynomials():
    """size-independent property: D x^a y^b is exact for a + b <= N (here at the 1e-9 level of the operator scale)"""
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(64, 48, 4.0, 3.0, seed=2).points
    nb, _ = emu.setup_knn(cl, 20)
    wx, wy = emu.setup_rbf_weights(cl, nb, 3, 3, 1)
    X, Y = cl[nb, 0], cl[nb, 1]
    scale = np.abs(wx).sum(axis=1).max()
    assert np.abs(wx.sum(axis=1)).max() <= 1e-9 * scale and np.abs(wy.sum(axis=1)).max() <= 1e-9 * scale
    assert np.abs((wx * X).sum(axis=1) - 1.0).max() <= 1e-8 and np.abs((wy * Y).sum(axis=1) - 1.0).max() <= 1e-8
    assert np.abs((wx * Y).sum(axis=1)).max() <= 1e-8 and np.abs((wy * X).sum(axis=1)).max() <= 1e-8
    f = X ** 2 * Y
    assert np.abs((wx * f).sum(axis=1) - 2 * cl[:, 0] * cl[:, 1]).max() <= 1e-7
    # second derivatives: L (x^2 + y^2) = 4
    wxx, wyy = emu.setup_rbf_weights(cl, nb, 5, 3, 2)
    assert np.abs(((wxx + wyy) * (X ** 2 + Y ** 2)).sum(axis=1) - 4.0).max() <= 1e-6


def test_weights_do_not_depend_on_the_launch_chunking():
    s = cases.fixture_setup(p=3, N=3)
    a = emu.setup_rbf_weights(s["points"], s["nb"], 3, 3, 1)
    b = emu.setup_rbf_weights(s["points"], s["nb"], 3, 3, 1, scratch_bytes=1)      # 256 points per launch
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_weights_argument_and_degenerate_stencil_errors():
    rng = np.random.default_rng(1)
    pts = rng.random((200, 2))
    nb, _ = emu.setup_knn(pts, 20)
    with pytest.raises(emu.EmuError, match="derivative order"):
        emu.setup_rbf_weights(pts, nb, 3, 3, 5)
    with pytest.raises(emu.EmuError, match="odd"):
        emu.setup_rbf_weights(pts, nb, 4, 3, 1)
    with pytest.raises(emu.EmuError, match="monomials"):
        emu.setup_rbf_weights(pts, nb[:, :8], 3, 3, 1)
    bad = nb.copy()
    bad[5, 3] = 200
    with pytest.raises(emu.EmuError, match="out of range"):
        emu.setup_rbf_weights(pts, bad, 3, 3, 1)
    line = np.stack([rng.random(100), np.zeros(100)], 1)          # collinear stencil: the reference's `\\` would throw
    nbl, _ = emu.setup_knn(line, 20)
    with pytest.raises(emu.EmuError, match="singular"):
        emu.setup_rbf_weights(line, nbl, 3, 3, 1)


def test_product_library_refuses_setup_without_a_device():
    import mft_b200 as m

    lib = m._lib.load()
    if lib.mft_device_count() > 0:
        pytest.skip("GPU box: covered by tests/test_zz_g_setup_gpu.py")
    pts = np.random.default_rng(0).random((50, 2))
    with pytest.raises(m._lib.MftError, match="no CPU fallback"):
        m.setup_ops.knn_device(pts, 5)
    nb, _ = emu.setup_knn(pts, 20)
    with pytest.raises(m._lib.MftError, match="no CPU fallback"):
        m.setup_ops.rbf_fd_weights_device(pts, nb, 3, 3)


def test_knn_query_subsets_and_weight_row_subsets():
    """mft_setup_knn_queries / mft_setup_rbf_weights_rows (what a rank of a partitioned cloud calls): the listed points
    query against all points; any order, duplicates allowed"""
    rng = np.random.default_rng(11)
    pts = rng.random((4000, 2)) * [2.0, 1.0]
    full_nb, full_d = emu.setup_knn(pts, 20)
    q = rng.integers(0, len(pts), 700)
    q[5] = q[6]
    nb, d = emu.setup_knn(pts, 20, queries=q)
    assert np.array_equal(nb, full_nb[q]) and np.array_equal(d, full_d[q])
    nb0, _ = emu.setup_knn(pts, 20, queries=np.zeros(0, np.int64))
    assert nb0.shape == (0, 20)
    with pytest.raises(emu.EmuError, match="query index"):
        emu.setup_knn(pts, 20, queries=np.array([4000]))
    wx, wy = emu.setup_rbf_weights(pts, full_nb, 3, 3, 1)
    rx, ry = emu.setup_rbf_weights(pts, full_nb[q], 3, 3, 1)
    assert np.array_equal(rx, wx[q]) and np.array_equal(ry, wy[q])


def test_partition_planner_with_device_setup_functions_matches_host_planner():
    """build_rank_partition with the (emulated) device kNN / weight functions injected gives every rank the same partition,
    halo plan and neighbour tables as the host KD-tree / LAPACK path; operator values agree to rounding"""
    import threading

    import mft_b200 as m
    from mft_b200 import partition

    cl = m.cloud.jittered_lattice(72, 60, 6.0, 5.0, seed=4)
    R = 3

    class ThreadComm:
        def __init__(self):
            self.barrier, self.slots = threading.Barrier(R), [None] * R

        def make(self, rank):
            def allgather(obj):
                self.slots[rank] = obj
                self.barrier.wait()
                out = list(self.slots)
                self.barrier.wait()
                return out
            return allgather

    def run(device_like):
        comm, parts, errs = ThreadComm(), [None] * R, []

        def work(r):
            try:
                kw = {}
                if device_like:
                    kw = dict(knn_queries=lambda pts, q, nv: emu.setup_knn(pts, nv, queries=q),
                              weights_rows=lambda pts, rows, p, N: emu.setup_rbf_weights(pts, rows, p, N, 1))
                parts[r] = partition.build_rank_partition(cl.points, [np.asarray(b) for b in cl.boundary_idxs],
                                                          cl.boundary_normals, r, R, 3, 3, 20, comm.make(r), **kw)
            except Exception as e:   # noqa: BLE001
                errs.append(e)
                comm.barrier.abort()
        th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
        [t.start() for t in th]
        [t.join() for t in th]
        assert not errs, errs
        return parts

    host, dev = run(False), run(True)
    for a, b in zip(host, dev):
        assert np.array_equal(a.owned_gid, b.owned_gid) and np.array_equal(a.halo_gid, b.halo_gid)
        assert np.array_equal(a.neighbors_owned, b.neighbors_owned)
        assert a.dx_min == b.dx_min and a.dx_avg == b.dx_avg and a.peers == b.peers
        assert all(np.array_equal(x, y) for x, y in zip(a.send_idx, b.send_idx)) and a.recv_count == b.recv_count
        for A, B in zip(a.ops, b.ops):
            assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
            assert np.abs(A.data - B.data).max() <= 1e-8 * np.abs(A.data).max()


def test_knn_property_random_clouds():
    """hypothesis: any finite cloud (random size, anisotropy, clustering, duplicated points, lattice snapping) -> the exact
    (distance, index)-ordered tables of the all-pairs reference"""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(1, 2 ** 31 - 1), st.integers(1, 260), st.integers(1, 24), st.sampled_from([1.0, 1e-3, 1e3]),
           st.sampled_from([1.0, 1e-6, 50.0]), st.sampled_from(["uniform", "normal", "snapped", "dups"]))
    def check(seed, n, k, scale, aspect, kind):
        rng = np.random.default_rng(seed)
        k = min(k, n)
        if kind == "normal":
            pts = rng.normal(size=(n, 2))
        else:
            pts = rng.random((n, 2))
        if kind == "snapped":
            pts = np.round(pts * 8) / 8          # many exact ties and coincident points
        if kind == "dups" and n > 3:
            pts[rng.integers(0, n, n // 3)] = pts[0]
        pts = np.ascontiguousarray(pts * [scale, scale * aspect] + [rng.normal() * scale, 0.0])
        nb, d = emu.setup_knn(pts, k)
        rb, rd = brute_knn(pts, k)
        assert np.array_equal(nb, rb) and np.array_equal(d, rd)

    check()


def test_weights_scale_exactly_with_a_power_of_two_dilation():
    """the stencil is normalised per axis before the solve, so doubling the coordinates halves first-derivative weights and
    quarters second-derivative weights bit for bit; an anisotropic dilation acts per axis"""
    s = cases.fixture_setup(p=5, N=3)
    pts, nb = s["points"], s["nb"]
    wx, wy = emu.setup_rbf_weights(pts, nb, 5, 3, 1)
    w2x, w2y = emu.setup_rbf_weights(pts * [2.0, 0.25], nb, 5, 3, 1)
    assert np.array_equal(w2x, wx / 2.0) and np.array_equal(w2y, wy * 4.0)
    vx, vy = emu.setup_rbf_weights(pts, nb, 5, 3, 2)
    v2x, v2y = emu.setup_rbf_weights(pts * 2.0, nb, 5, 3, 2)
    assert np.array_equal(v2x, vx / 4.0) and np.array_equal(v2y, vy / 4.0)


@pytest.mark.parametrize("hyb,p,N,k", [((1.0, 1.0, 1.0), 3, 3, None), ((0.5, 2.0, 0.7), 5, 3, 2), ((2.0, 1.0, 1.5), 3, 2, None)])
def test_hybrid_gaussian_phs_weights(hyb, p, N, k):
    """RBF(HybridGaussianPHS(Nrbf, alpha, beta, epsilon)) (geometry_primatives.jl:117-132, 238-262): the emulated device
    kernel and the host mirror against the oracle's symbolic restatement; polynomial reproduction still holds"""
    import mft_b200 as m

    s = cases.fixture_setup(p=p, N=N)
    ref = orc.compute_flux_operator(s["points"], s["nb"], p, N, k, hybrid=hyb)
    wx, wy = emu.setup_rbf_weights(s["points"], s["nb"], p, N, k or 1, hybrid=hyb)
    hx, hy = m.setup_ops.rbf_fd_weights(s["points"], s["nb"], p, N, k, hybrid=hyb)
    for w, h, B in zip((wx, wy), (hx, hy), ref):
        A = m.setup_ops.assemble_csc(s["nb"], w)
        H = m.setup_ops.assemble_csc(s["nb"], h)
        assert np.array_equal(A.indices, B.indices)
        assert np.abs(A.data - B.data).max() <= 1e-8 * np.abs(B.data).max()
        assert np.abs(H.data - B.data).max() <= 1e-8 * np.abs(B.data).max()
    phs = orc.compute_flux_operator(s["points"], s["nb"], p, N, k)
    assert np.abs(phs[0].data - ref[0].data).max() > 1e-6 * np.abs(ref[0].data).max()      # really a different basis
    if k is None:
        X = s["points"][s["nb"], 0]
        assert np.abs(wx.sum(axis=1)).max() <= 1e-9 * np.abs(wx).sum(axis=1).max() and np.abs((wx * X).sum(axis=1) - 1).max() <= 1e-8
    basis = m.PointCloudBasis(m.Point2D(), N, approximation_type=m.RBF(m.HybridGaussianPHS(p, *hyb)))
    assert m.api._hybrid_of(basis) == hyb and basis.approx_type.rbf_type.Nrbf == p


def test_weights_differentiate_a_smooth_function_with_the_expected_order():
    """independent of the oracle: Dx, Dy from the emulated device kernel applied to sin(x) cos(y) converge like h^N (interior
    points of jittered clouds, PHS r^5 + polynomials of degree N = 3: halving h cuts the error by ~2^3 or better)"""
    import mft_b200 as m

    errs = []
    for n_side in (24, 48, 96):
        cl = m.cloud.jittered_lattice(n_side, n_side, 2.0, 2.0, seed=7).points
        nb, _ = emu.setup_knn(cl, 20)
        wx, wy = emu.setup_rbf_weights(cl, nb, 5, 3, 1)
        f = np.sin(cl[:, 0]) * np.cos(cl[:, 1])
        fx = (wx * f[nb]).sum(axis=1)
        fy = (wy * f[nb]).sum(axis=1)
        inner = np.all((cl > 0.3) & (cl < 1.7), axis=1)          # one-sided boundary stencils converge a bit slower
        ex = np.abs(fx - np.cos(cl[:, 0]) * np.cos(cl[:, 1]))[inner].max()
        ey = np.abs(fy + np.sin(cl[:, 0]) * np.sin(cl[:, 1]))[inner].max()
        errs.append(max(ex, ey))
    assert errs[0] < 1e-3 and errs[2] < errs[1] < errs[0]
    assert errs[0] / errs[1] > 6.0 and errs[1] / errs[2] > 6.0, errs
    # second derivatives (k = 2): Laplacian of the same function, order N - 1
    cl = m.cloud.jittered_lattice(96, 96, 2.0, 2.0, seed=7).points
    nb, _ = emu.setup_knn(cl, 20)
    wxx, wyy = emu.setup_rbf_weights(cl, nb, 5, 3, 2)
    f = np.sin(cl[:, 0]) * np.cos(cl[:, 1])
    lap = ((wxx + wyy) * f[nb]).sum(axis=1)
    inner = np.all((cl > 0.3) & (cl < 1.7), axis=1)
    assert np.abs(lap + 2 * f)[inner].max() < 5e-3
