"""Setup-time operator generation (NOT on the timed path; north_star: "kNN search and stencil-weight generation
remain setup-time code").  Batched numpy implementation of what the reference does point by point:

  PointData ctor        src/domains/PointCloudDomain/geometry_primatives.jl:322-339  (kNN, dx_min, dx_avg)
  compute_flux_operator src/solvers/pointcloudsolver/compute_operators.jl:409-453 (1st derivatives), :549-594 (k-th)

In the drop-in deployment the Julia side keeps generating operators with its own code and hands the CSC arrays to
`mft_set_operator_csc`; this module exists so that the Python host mirror, the tests and bench.py can build the
same operators without the reference.
"""
from __future__ import annotations

import math
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree


def num_neighbors(N: int, dim: int = 2) -> int:
    """NV = max(2*binomial(N+d,d), [10,15,20][d])   geometry_primatives.jl:197-198"""
    return max(2 * math.comb(N + dim, dim), [10, 15, 20][dim - 1])


def knn(points: np.ndarray, nv: int, workers: int = -1):
    """Exact kNN sorted by distance, self first (NearestNeighbors.knn(tree, pts, nv, true)).
    Returns neighbors (N,nv) int64 0-based, dx_min, dx_avg (from the nearest-neighbour distance)."""
    tree = cKDTree(points)
    idx, d = knn_query(tree, points, nv, workers=workers)
    return idx, float(d[:, 1].min()), float(d[:, 1].mean())


TIE_MARGIN = 4


def knn_query(tree, queries, nv, workers=-1, index_map=None):
    """nv nearest neighbours with exact distance ties resolved by ascending point index -- also ties that straddle
    the k-th place (a few extra candidates are fetched, ordered canonically, then cut to nv)."""
    kq = min(nv + TIE_MARGIN, tree.n)
    d, idx = tree.query(queries, k=kq, workers=workers)
    idx = idx.astype(np.int64)
    if index_map is not None:
        idx = index_map[idx]
    idx, d = canonical_ties(idx, d)
    return np.ascontiguousarray(idx[:, :nv]), np.ascontiguousarray(d[:, :nv])


def canonical_ties(idx: np.ndarray, d: np.ndarray):
    """Order equidistant neighbours by ascending point index.  NearestNeighbors.jl leaves the order of exact distance
    ties implementation-defined; fixing it makes the neighbour tables -- and therefore the
    operator weights -- independent of which tree / which rank produced them."""
    o = np.argsort(idx, axis=1, kind="stable")
    idx, d = np.take_along_axis(idx, o, 1), np.take_along_axis(d, o, 1)
    o = np.argsort(d, axis=1, kind="stable")
    return np.take_along_axis(idx, o, 1), np.take_along_axis(d, o, 1)


def monomial_exponents(N: int):
    return [(a, d - a) for d in range(N + 1) for a in range(d, -1, -1)]


def _phs_axis_derivative(x, r2, p: int, k: int, hybrid=None):
    """k-th derivative with respect to x of phi = f(x^2+y^2): f(s) = s^q, q = p/2 (PolyharmonicSpline) or, hybrid =
    (alpha, beta, epsilon), f(s) = alpha exp(-epsilon^2 s) + beta s^q (HybridGaussianPHS, geometry_primatives.jl:238-262)."""
    q = p / 2.0

    def fd(m):  # m-th derivative of f
        c = 1.0
        for i in range(m):
            c *= q - i
        phs = c * np.power(r2, q - m)
        if hybrid is None:
            return phs
        a, b, e = hybrid
        return a * (-(e * e)) ** m * np.exp(-(e * e) * r2) + b * phs

    if k == 0:
        return fd(0)
    if k == 1:
        return 2.0 * x * fd(1)
    if k == 2:
        return 2.0 * fd(1) + 4.0 * x * x * fd(2)
    if k == 3:
        return 12.0 * x * fd(2) + 8.0 * x ** 3 * fd(3)
    if k == 4:
        return 12.0 * fd(2) + 48.0 * x * x * fd(3) + 16.0 * x ** 4 * fd(4)
    raise NotImplementedError("derivative order %d" % k)


def rbf_fd_weights(points: np.ndarray, neighbors: np.ndarray, p: int, N: int, k: int | None = None,
                   chunk: int = 4096, workers: int | None = None, hybrid=None):
    """Per-point weights Dx_loc, Dy_loc (N,nv): shift_stencil / [R P; P' 0] \\ rhs / rescale, batched.
    k=None: first derivatives; k: pure k-th derivatives d^k/dx^k, d^k/dy^k (as the reference builds them)."""
    npts, nv = neighbors.shape
    exps = monomial_exponents(N)
    npoly = len(exps)
    kk = 1 if k is None else k
    m = nv + npoly
    eps = np.finfo(np.float64).eps
    wx = np.empty((npts, nv))
    wy = np.empty((npts, nv))
    pr_x = np.array([math.factorial(kk) if (a, b) == (kk, 0) else 0.0 for (a, b) in exps])
    pr_y = np.array([math.factorial(kk) if (a, b) == (0, kk) else 0.0 for (a, b) in exps])
    ea = np.array([e[0] for e in exps], dtype=np.float64)
    eb = np.array([e[1] for e in exps], dtype=np.float64)
    def work(s0):
        s1 = min(npts, s0 + chunk)
        B = s1 - s0
        X = points[neighbors[s0:s1]]                      # (B,nv,2)
        Xs = X - X[:, :1, :]
        sc = 1.0 / np.abs(Xs).max(axis=1)                 # (B,2) per-axis scaling, compute_operators.jl:239-242
        Xs = Xs * sc[:, None, :]
        dx = Xs[:, :, None, 0] - Xs[:, None, :, 0]
        dy = Xs[:, :, None, 1] - Xs[:, None, :, 1]
        M = np.zeros((B, m, m))
        M[:, :nv, :nv] = _phs_axis_derivative(None, dx * dx + dy * dy, p, 0, hybrid)
        P = np.power(Xs[:, :, None, 0], ea[None, None, :]) * np.power(Xs[:, :, None, 1], eb[None, None, :])
        M[:, :nv, nv:] = P
        M[:, nv:, :nv] = np.transpose(P, (0, 2, 1))
        # rhs at mirrored stencil x_c - x_j with the centre at (eps,eps)   compute_operators.jl:248-263
        mx = -Xs[:, :, 0].copy()
        my = -Xs[:, :, 1].copy()
        mx[:, 0] = eps
        my[:, 0] = eps
        r2 = mx * mx + my * my
        rhs = np.zeros((B, m, 2))
        rhs[:, :nv, 0] = _phs_axis_derivative(mx, r2, p, kk, hybrid)
        rhs[:, :nv, 1] = _phs_axis_derivative(my, r2, p, kk, hybrid)
        rhs[:, nv:, 0] = pr_x
        rhs[:, nv:, 1] = pr_y
        W = np.linalg.solve(M, rhs)
        wx[s0:s1] = (sc[:, 0] ** kk)[:, None] * W[:, :nv, 0]
        wy[s0:s1] = (sc[:, 1] ** kk)[:, None] * W[:, :nv, 1]

    starts = list(range(0, npts, chunk))
    nthreads = min(len(starts), workers or os.cpu_count() or 1)
    if nthreads > 1:
        with ThreadPoolExecutor(nthreads) as ex:   # numpy releases the GIL inside solve / power
            list(ex.map(work, starts))
    else:
        for s0 in starts:
            work(s0)
    return wx, wy


def assemble_csc(neighbors: np.ndarray, w: np.ndarray, ncols: int | None = None):
    """sparse(vec(rows), vec(cols), vec(vals)) -> CSC with explicit zeros kept (compute_operators.jl:443-452)."""
    npts, nv = neighbors.shape
    rows = np.repeat(np.arange(npts, dtype=np.int64), nv)
    A = sp.coo_matrix((w.reshape(-1), (rows, neighbors.reshape(-1))), shape=(npts, ncols or npts)).tocsc()
    A.sort_indices()
    return A


def compute_flux_operator(points, neighbors, p: int, N: int, k: int | None = None, hybrid=None):
    wx, wy = rbf_fd_weights(points, neighbors, p, N, k, hybrid=hybrid)
    return [assemble_csc(neighbors, wx), assemble_csc(neighbors, wy)]


def julia_csc(A):
    """(colptr, rowval, nzval) exactly as Julia's SparseMatrixCSC{Float64,Int64} holds them (1-based)."""
    A = sp.csc_matrix(A)
    A.sort_indices()
    return (np.ascontiguousarray(A.indptr, dtype=np.int64) + 1, np.ascontiguousarray(A.indices, dtype=np.int64) + 1,
            np.ascontiguousarray(A.data, dtype=np.float64))


# ---- the same two setup steps on the device (libmft_b200.so, no CPU path) -----------
def knn_device(points: np.ndarray, nv: int, device: int = 0):
    """mft_setup_knn: cell-grid kNN on the GPU.  Same result contract as knn() (neighbour tables bit-identical:
    distance-sorted, self first, exact ties by ascending index)."""
    from . import _lib as L

    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = pts.shape[0]
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    nbr1 = np.empty((n, nv), dtype=np.int64)
    dist = np.empty((n, nv), dtype=np.float64)
    L.check(L.load().mft_setup_knn(device, n, L.ptr(x), L.ptr(y), nv, L.ptr(nbr1), L.ptr(dist)))
    nbr1 -= 1
    return nbr1, float(dist[:, 1].min()), float(dist[:, 1].mean())


def rbf_fd_weights_device(points: np.ndarray, neighbors: np.ndarray, p: int, N: int, k: int | None = None, device: int = 0):
    """mft_setup_rbf_weights: one GPU thread per point builds and solves its (nv + npoly)^2 system.  Same contract as
    rbf_fd_weights() (weights agree to rounding: different elimination order)."""
    from . import _lib as L

    pts = np.ascontiguousarray(points, dtype=np.float64)
    n, nv = neighbors.shape
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    nbr1 = np.ascontiguousarray(neighbors, dtype=np.int64) + 1
    wx = np.empty((n, nv), dtype=np.float64)
    wy = np.empty((n, nv), dtype=np.float64)
    L.check(L.load().mft_setup_rbf_weights(device, n, L.ptr(x), L.ptr(y), nv, L.ptr(nbr1), p, N, 1 if k is None else k,
                                           L.ptr(wx), L.ptr(wy)))
    return wx, wy


def knn_queries_device(points: np.ndarray, queries: np.ndarray, nv: int, device: int = 0):
    """mft_setup_knn_queries: the nv nearest of `points` for the points listed in `queries` (0-based positions);
    returns (neighbours (nq,nv) 0-based, distances (nq,nv)).  One rank of a partitioned cloud calls this on its padded box."""
    from . import _lib as L

    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = pts.shape[0]
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    q1 = np.ascontiguousarray(np.asarray(queries, dtype=np.int64) + 1)
    nq = len(q1)
    nbr1 = np.empty((nq, nv), dtype=np.int64)
    dist = np.empty((nq, nv), dtype=np.float64)
    if nq:
        L.check(L.load().mft_setup_knn_queries(device, n, L.ptr(x), L.ptr(y), nv, nq, L.ptr(q1), L.ptr(nbr1), L.ptr(dist)))
    nbr1 -= 1
    return nbr1, dist


def rbf_fd_weights_rows_device(points: np.ndarray, rows_nb: np.ndarray, p: int, N: int, k: int | None = None, device: int = 0,
                               hybrid=None):
    """mft_setup_rbf_weights_rows / _hybrid: weights for the stencils given as rows of rows_nb (indices into points)"""
    from . import _lib as L

    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = pts.shape[0]
    n_rows, nv = rows_nb.shape
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    nbr1 = np.ascontiguousarray(rows_nb, dtype=np.int64) + 1
    wx = np.empty((n_rows, nv), dtype=np.float64)
    wy = np.empty((n_rows, nv), dtype=np.float64)
    if n_rows and hybrid is not None:
        a, b, e = (float(v) for v in hybrid)
        L.check(L.load().mft_setup_rbf_weights_hybrid(device, n, L.ptr(x), L.ptr(y), n_rows, nv, L.ptr(nbr1), p, a, b, e, N,
                                                      1 if k is None else k, L.ptr(wx), L.ptr(wy)))
    elif n_rows:
        L.check(L.load().mft_setup_rbf_weights_rows(device, n, L.ptr(x), L.ptr(y), n_rows, nv, L.ptr(nbr1), p, N,
                                                    1 if k is None else k, L.ptr(wx), L.ptr(wy)))
    return wx, wy


def compute_flux_operator_device(points, neighbors, p: int, N: int, k: int | None = None, device: int = 0, hybrid=None):
    if hybrid is not None:
        wx, wy = rbf_fd_weights_rows_device(points, neighbors, p, N, k, device, hybrid)
    else:
        wx, wy = rbf_fd_weights_device(points, neighbors, p, N, k, device)
    return [assemble_csc(neighbors, wx), assemble_csc(neighbors, wy)]


def knn_with(engine, points, nv):
    """PointData ctor for an engine: RBFFDEngineCUDA(setup="device") searches on the GPU."""
    if getattr(engine, "setup", "host") == "device":
        return knn_device(points, nv, engine.device)
    return knn(points, nv)


def flux_operator_with(engine, points, neighbors, p, N, k=None, hybrid=None):
    if getattr(engine, "setup", "host") == "device":
        return compute_flux_operator_device(points, neighbors, p, N, k, engine.device, hybrid)
    return compute_flux_operator(points, neighbors, p, N, k, hybrid)
