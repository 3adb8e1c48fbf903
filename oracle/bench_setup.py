"""oracle/bench_setup.py -- TEST / MEASUREMENT INFRASTRUCTURE (never imported by the product).

Setup for the CPU arm of bench.py (`--impl reference`) that needs nothing of the product's shared library: the same
synthetic cloud in the same numbering as the GPU arm (Hilbert order restated in numpy), kNN through the oracle's
`point_data`, and the RBF-FD weights as a batched restatement of the reference's per-point solve
(compute_operators.jl:409-453: shift_stencil :225-246, rbf_block / poly_block :191-223, mirrored rhs :248-263) so that a
million points take seconds, not the minutes of the oracle's point-by-point `compute_flux_operator` (which stays the
parity reference for the small cases; the two agree to rounding, tests/test_oracle_vs_numpy.py::test_batched_setup).
"""
from __future__ import annotations

import importlib.util
import math
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def cloud_module():
    """the synthetic cloud generator / initial conditions of the bench (pure numpy source file of the product package, loaded
    by path: importing it loads no shared library)"""
    spec = importlib.util.spec_from_file_location("_mft_cloud_only", os.path.join(HERE, "..", "meshfreetrixi.jl_b200", "cloud.py"))
    mod = importlib.util.module_from_spec(spec)
    import sys

    sys.modules["_mft_cloud_only"] = mod   # (dataclasses resolve the module by name)
    spec.loader.exec_module(mod)
    return mod


def hilbert_order(points: np.ndarray, bits: int = 20) -> np.ndarray:
    """numbering along a Hilbert curve through the bounding square, 2^bits cells per side, ties by index: perm[d] = old index
    of new point d.  Same quantisation and curve as the product's mft_sfc_order, so both arms number the cloud alike."""
    x, y = points[:, 0], points[:, 1]
    xmin, ymin = x.min(), y.min()
    ext = max(x.max() - xmin, y.max() - ymin, 1e-300)
    scale = float((1 << bits) - 1) / ext
    qx = ((x - xmin) * scale).astype(np.uint64)
    qy = ((y - ymin) * scale).astype(np.uint64)
    d = np.zeros(len(x), dtype=np.uint64)
    full = np.uint64((1 << bits) - 1)
    s = 1 << (bits - 1)
    while s > 0:
        su = np.uint64(s)
        rx = ((qx & su) != 0).astype(np.uint64)
        ry = ((qy & su) != 0).astype(np.uint64)
        d += su * su * ((np.uint64(3) * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        qx = np.where(flip, full - qx, qx)
        qy = np.where(flip, full - qy, qy)
        swap = ry == 0
        qx, qy = np.where(swap, qy, qx), np.where(swap, qx, qy)
        s >>= 1
    return np.argsort(d, kind="stable").astype(np.int64)


def _monomials(N: int):
    return [(a, d - a) for d in range(N + 1) for a in range(d, -1, -1)]


def flux_operator_batched(points: np.ndarray, neighbors: np.ndarray, p: int = 3, N: int = 3, chunk: int = 8192):
    """[Dx, Dy] (scipy CSC, explicit zeros kept, sorted rows = Julia's sparse(I, J, V)) for the PHS r^p basis with
    polynomials up to degree N, first derivatives; batched LU (np.linalg.solve) instead of the per-point Bunch-Kaufman."""
    if p % 2 == 0:
        raise ValueError("PHS exponent must be odd")
    npts, nv = neighbors.shape
    exps = _monomials(N)
    npoly = len(exps)
    m = nv + npoly
    ea = np.array([e[0] for e in exps], dtype=np.float64)
    eb = np.array([e[1] for e in exps], dtype=np.float64)
    pr_x = np.array([1.0 if e == (1, 0) else 0.0 for e in exps])
    pr_y = np.array([1.0 if e == (0, 1) else 0.0 for e in exps])
    eps = np.finfo(np.float64).eps
    wx = np.empty((npts, nv))
    wy = np.empty((npts, nv))
    for s0 in range(0, npts, chunk):
        s1 = min(npts, s0 + chunk)
        B = s1 - s0
        Xs = points[neighbors[s0:s1]]
        Xs = Xs - Xs[:, :1, :]
        sc = 1.0 / np.abs(Xs).max(axis=1)
        Xs = Xs * sc[:, None, :]
        dx = Xs[:, :, None, 0] - Xs[:, None, :, 0]
        dy = Xs[:, :, None, 1] - Xs[:, None, :, 1]
        M = np.zeros((B, m, m))
        M[:, :nv, :nv] = np.sqrt(dx * dx + dy * dy) ** p
        P = np.power(Xs[:, :, None, 0], ea[None, None, :]) * np.power(Xs[:, :, None, 1], eb[None, None, :])
        M[:, :nv, nv:] = P
        M[:, nv:, :nv] = np.transpose(P, (0, 2, 1))
        mx, my = -Xs[:, :, 0].copy(), -Xs[:, :, 1].copy()
        mx[:, 0] = eps
        my[:, 0] = eps
        rpm2 = np.sqrt(mx * mx + my * my) ** (p - 2)
        rhs = np.zeros((B, m, 2))
        rhs[:, :nv, 0] = p * mx * rpm2          # d/dx r^p = p x r^(p-2)
        rhs[:, :nv, 1] = p * my * rpm2
        rhs[:, nv:, 0] = pr_x
        rhs[:, nv:, 1] = pr_y
        W = np.linalg.solve(M, rhs)
        wx[s0:s1] = sc[:, :1] * W[:, :nv, 0]
        wy[s0:s1] = sc[:, 1:2] * W[:, :nv, 1]
    rows = np.repeat(np.arange(npts, dtype=np.int64), nv)
    out = []
    for w in (wx, wy):
        A = sp.coo_matrix((w.reshape(-1), (rows, neighbors.reshape(-1))), shape=(npts, npts)).tocsc()
        A.sort_indices()
        out.append(A)
    return out


def num_neighbors(N: int, dim: int = 2) -> int:
    return max(2 * math.comb(N + dim, dim), [10, 15, 20][dim - 1])
