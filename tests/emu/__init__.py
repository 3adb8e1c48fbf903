"""Host emulation of the product's sync-free device kernels (TEST INFRASTRUCTURE).  The product's per-thread device
functions (meshfreetrixi.jl_b200/csrc/*_kernels.cuh, `MFT_HD` bodies) and their host orchestration are compiled with g++
and driven by a host loop, so the CPU-only test tier can check kernel logic against the oracle.  The product never loads
this library."""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        _lib.emu_last_error.restype = C.c_char_p
    return _lib


class EmuError(RuntimeError):
    pass


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise EmuError(lib().emu_last_error().decode())


def setup_knn(points, k, queries=None):
    """emulated mft_setup_knn / mft_setup_knn_queries -> (neighbors 0-based (nq,k), distances (nq,k));
    queries: 0-based indices of the points that query (None: all)"""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = len(pts)
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    q1 = None if queries is None else np.ascontiguousarray(np.asarray(queries, dtype=np.int64) + 1)
    nq = n if q1 is None else len(q1)
    nb = np.empty((nq, k), np.int64)
    d = np.empty((nq, k))
    _check(lib().emu_setup_knn(C.c_int64(n), _p(x), _p(y), C.c_int(k), C.c_int64(nq), _p(q1), _p(nb), _p(d)))
    return nb - 1, d


def setup_rbf_weights(points, neighbors, p, N, kk=1, scratch_bytes=0, hybrid=None):
    """emulated mft_setup_rbf_weights[_rows|_hybrid] -> (wx, wy) (n,k); scratch_bytes > 0 forces small launches (chunking);
    hybrid = (alpha, beta, epsilon) selects the HybridGaussianPHS basis"""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    n = len(pts)
    n_rows, k = neighbors.shape          # rows may be any subset / multiset of the points (their indices stay global)
    x, y = np.ascontiguousarray(pts[:, 0]), np.ascontiguousarray(pts[:, 1])
    nb1 = np.ascontiguousarray(neighbors, dtype=np.int64) + 1
    wx = np.empty((n_rows, k))
    wy = np.empty((n_rows, k))
    if hybrid is not None:
        a, b, e = (float(v) for v in hybrid)
        _check(lib().emu_setup_rbf_weights_hybrid(C.c_int64(n), _p(x), _p(y), C.c_int64(n_rows), C.c_int(k), _p(nb1), C.c_int(p),
                                                  C.c_double(a), C.c_double(b), C.c_double(e), C.c_int(N), C.c_int(kk), _p(wx), _p(wy)))
        return wx, wy
    _check(lib().emu_setup_rbf_weights(C.c_int64(n), _p(x), _p(y), C.c_int64(n_rows), C.c_int(k), _p(nb1), C.c_int(p), C.c_int(N),
                                       C.c_int(kk), _p(wx), _p(wy), C.c_int64(scratch_bytes)))
    return wx, wy


def limiter_zhang_shu(u, neighbors, thresholds, variables, gamma):
    """emulated k_zs_detect / k_zs_apply passes, in place on u (4,N); neighbors (N,k) 0-based in kNN list order"""
    assert u.dtype == np.float64 and u.flags.c_contiguous
    nb = np.ascontiguousarray(neighbors, dtype=np.int64)
    n, k = nb.shape
    thr = np.asarray(thresholds, dtype=np.float64)
    var = np.asarray(variables, dtype=np.int32)
    rc = lib().emu_limiter_zhang_shu(C.c_int64(n), C.c_int(k), _p(nb), C.c_double(gamma), C.c_int(len(thr)), _p(thr), _p(var), _p(u))
    if rc != 0:
        raise EmuError("emu_limiter_zhang_shu failed")
    return u


def igr_apply(neighbors, wx, wy, alpha, maxiter, u, du, b_out=None):
    """emulated launch_igr(): neighbors/wx/wy (n,k) with each row already in ascending-column order; u, du (4,n);
    du is updated in place.  Returns (sigma, (iterations, |r|, |r0|))"""
    nb = np.ascontiguousarray(neighbors, dtype=np.int64)
    n, k = nb.shape
    wx = np.ascontiguousarray(wx, dtype=np.float64)
    wy = np.ascontiguousarray(wy, dtype=np.float64)
    assert u.flags.c_contiguous and du.flags.c_contiguous and u.dtype == np.float64 and du.dtype == np.float64
    sigma = np.empty(n)
    st = np.empty(3)
    rc = lib().emu_igr_apply(C.c_int64(n), C.c_int(k), _p(nb), _p(wx), _p(wy), C.c_double(alpha), C.c_int(maxiter), _p(u), _p(du),
                             _p(sigma), _p(st), _p(b_out))
    if rc != 0:
        raise EmuError("emu_igr_apply failed")
    return sigma, (int(st[0]), float(st[1]), float(st[2]))


def count_nonfinite(u):
    """emulated k_count_nonfinite on a (V,N) state (laid out AoS like on the device)"""
    aos = np.ascontiguousarray(np.asarray(u, dtype=np.float64).T)
    f = lib().emu_count_nonfinite
    f.restype = C.c_int64
    return int(f(C.c_int64(aos.shape[0]), C.c_int(aos.shape[1]), _p(aos)))
