"""
mft_oracle.py -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).

numpy/scipy restatement of the setup side of MeshfreeTrixi.jl (Medusa reader, kNN, dx_min/dx_avg,
RBF-FD weight generation) plus a ctypes driver for the C restatement of `rhs!` in mft_oracle.c.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  Citations are file:line relative to /root/reference/.

Parity pins: the reference has no golden vectors.  This oracle is pinned by the
reference's own test identities (divergence, upwind viscosity, history known answer) and by polynomial
reproduction of the generated operators; see tests/test_oracle_identities.py.  PARITY UNPINNED for:
update_residual_visc!, update_visc!, the BC passes, whole-rhs!, time integration, kNN tie order.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg
import scipy.sparse as sp
from scipy.spatial import cKDTree

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EQ_EULER2D, EQ_ADVECTION2D = 0, 1
BC_DIRICHLET, BC_SLIP_WALL, BC_DO_NOTHING = 0, 1, 2
SRC_HV_FLYER, SRC_HV_TOMINEC, SRC_UPWIND, SRC_RESIDUAL, SRC_IGR = 0, 1, 2, 3, 4


def build(force: bool = False) -> str:
    """Compile mft_oracle.c -> libmft_oracle.so (gcc, see oracle/Makefile)."""
    so = os.path.join(_HERE, "libmft_oracle.so")
    src = os.path.join(_HERE, "mft_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


_FAST = None


def fast_lib():
    """libmft_cpu_fast.so: CPU baseline (ii) of SURVEY.md 8(d) (fused row-parallel multi-threaded rhs!, mft_cpu_fast.c)."""
    global _FAST
    if _FAST is None:
        so = os.path.join(_HERE, "libmft_cpu_fast.so")
        src = os.path.join(_HERE, "mft_cpu_fast.c")
        if not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
            subprocess.check_call(["make", "-C", _HERE, "-s", "libmft_cpu_fast.so"])
        _FAST = C.CDLL(so)
        _FAST.fast_max_threads.restype = C.c_int
    return _FAST


class _FASTP(C.Structure):
    _fields_ = [("n", C.c_int64), ("ptr", C.c_void_p), ("col", C.c_void_p), ("wx", C.c_void_p), ("wy", C.c_void_p),
                ("tptr", C.c_void_p), ("tcol", C.c_void_p), ("twx", C.c_void_p), ("twy", C.c_void_p),
                ("nb", C.c_int64), ("bidx", C.c_void_p), ("bval", C.c_void_p),
                ("gamma", C.c_double), ("c_rv", C.c_double), ("c_uw", C.c_double), ("dx_avg", C.c_double),
                ("success_iter_zero", C.c_int), ("mean_divisor_vn", C.c_int), ("max_lexicographic", C.c_int)]


class FastCpuProblem:
    """Euler + residual-viscosity rhs! with Dirichlet tables on the fused multi-threaded CPU kernel (AoS state)."""

    def __init__(self, Dx, Dy, gamma, dx_avg, bidx, bvals, c_rv=1.0, c_uw=1.0, success_iter=5,
                 mean_divisor_vn=True, max_lexicographic=True):
        X = sp.csr_matrix(Dx)
        Y = sp.csr_matrix(Dy)
        X.sort_indices()
        Y.sort_indices()
        assert np.array_equal(X.indptr, Y.indptr) and np.array_equal(X.indices, Y.indices)
        XT = sp.csr_matrix(sp.csc_matrix(Dx).T)     # rows of D' = columns of D, ascending source row
        YT = sp.csr_matrix(sp.csc_matrix(Dy).T)
        XT.sort_indices()
        YT.sort_indices()
        self.n = X.shape[0]
        self._keep = [np.ascontiguousarray(X.indptr, dtype=np.int64), np.ascontiguousarray(X.indices, dtype=np.int32),
                      np.ascontiguousarray(X.data), np.ascontiguousarray(Y.data),
                      np.ascontiguousarray(XT.indptr, dtype=np.int64), np.ascontiguousarray(XT.indices, dtype=np.int32),
                      np.ascontiguousarray(XT.data), np.ascontiguousarray(YT.data),
                      np.ascontiguousarray(bidx, dtype=np.int32), np.ascontiguousarray(np.asarray(bvals).T.copy())]
        k = self._keep
        self.p = _FASTP(self.n, *[a.ctypes.data for a in k[:8]], len(k[8]), k[8].ctypes.data, k[9].ctypes.data,
                        float(gamma), float(c_rv), float(c_uw), float(dx_avg), int(success_iter == 0),
                        int(mean_divisor_vn), int(max_lexicographic))
        self.g = np.zeros((self.n, 8))

    def rhs(self, u_soa, approx_du_soa=None, reps=1):
        """u_soa (4,N) in/out semantics as rhs!; returns du (4,N)"""
        u = np.ascontiguousarray(u_soa.T.copy())
        ad = np.zeros_like(u) if approx_du_soa is None else np.ascontiguousarray(approx_du_soa.T.copy())
        du = np.empty_like(u)
        fast_lib().fast_rhs_rv_repeat(C.byref(self.p), C.c_void_p(u.ctypes.data), C.c_void_p(ad.ctypes.data),
                                      C.c_void_p(du.ctypes.data), C.c_void_p(self.g.ctypes.data), int(reps))
        u_soa[:] = u.T
        return np.ascontiguousarray(du.T)


# --------------------------------------------------------------------------------------------
# Setup side
# --------------------------------------------------------------------------------------------
def read_medusa_file(casename: str):
    """src/auxiliary/medusa/read_medusa_file.jl:2-64.  Returns points (N,2), interior_idx (0-based),
    boundary_idxs (list of 0-based arrays, group g = type -(g+1)), boundary_normals (list of (n,2))."""
    positions = np.loadtxt(casename + "_positions.txt", delimiter=",", dtype=np.float64, ndmin=2)
    types = np.loadtxt(casename + "_types.txt", dtype=np.int64, ndmin=1)
    boundary_idx = np.loadtxt(casename + "_boundary.txt", dtype=np.int64, ndmin=1)
    interior_idx = np.loadtxt(casename + "_interior.txt", dtype=np.int64, ndmin=1)
    normals = np.loadtxt(casename + "_normals.txt", delimiter=",", dtype=np.float64, ndmin=2)
    num_bound = int(-types.min())
    bidx = [[] for _ in range(num_bound)]
    bnrm = [[] for _ in range(num_bound)]
    for j, b in enumerate(boundary_idx):
        g = -int(types[b]) - 1
        if 0 <= g < num_bound:
            bidx[g].append(int(b))
            bnrm[g].append(normals[j])
    keep = [g for g in range(num_bound) if len(bidx[g]) > 0]
    boundary_idxs = [np.asarray(bidx[g], dtype=np.int64) for g in keep]
    # NB (quirk 15): the reference drops empty groups from idxs only.
    boundary_normals = [np.asarray(bnrm[g], dtype=np.float64).reshape(-1, 2) for g in range(num_bound)]
    return positions, interior_idx, boundary_idxs, boundary_normals


def num_neighbors(N: int, dim: int = 2) -> int:
    """RefPointData: NV = max(2*binomial(N+d,d), [10,15,20][d])  geometry_primatives.jl:197-198"""
    return max(2 * math.comb(N + dim, dim), [10, 15, 20][dim - 1])


def point_data(points: np.ndarray, nv: int):
    """PointData ctor, geometry_primatives.jl:322-339: knn(tree, pts, nv, sorted=true) + dx_min/dx_avg
    from the 2-NN distances.  Returns neighbors (N,nv) int64 0-based (self first), dx_min, dx_avg."""
    tree = cKDTree(points)
    d, idx = tree.query(points, k=min(nv + 4, len(points)))
    # exact distance ties: NearestNeighbors.jl's order is implementation-defined; canonicalise by
    # ascending index (index sort, then a stable distance sort; 4 extra candidates cover ties across the k-th place)
    idx = idx.astype(np.int64)
    o = np.argsort(idx, axis=1, kind="stable")
    idx, d = np.take_along_axis(idx, o, 1), np.take_along_axis(d, o, 1)
    o = np.argsort(d, axis=1, kind="stable")
    idx = np.ascontiguousarray(np.take_along_axis(idx, o, 1)[:, :nv])
    d2, _ = tree.query(points, k=2)
    return idx, float(d2[:, 1].min()), float(d2[:, 1].mean())


def _monomial_exponents(N: int):
    """monomials([x,y], 0:N) (geometry_primatives.jl:277-284): all x^a y^b with a+b <= N."""
    return [(a, d - a) for d in range(N + 1) for a in range(d, -1, -1)]


_RBF_CACHE = {}


def _phs_functions(p: int, k: int | None, hybrid=None):
    """phi = sqrt(x^2+y^2)^p -- or, hybrid = (alpha, beta, epsilon), alpha*exp(-(epsilon*r)^2) + beta*r^p
    (rbf_basis, geometry_primatives.jl:211-262) -- and its (k-th) x/y derivatives, built symbolically like
    concrete_rbf_flux_basis (compute_operators.jl:9-82 with Symbolics)."""
    key = (p, k, None if hybrid is None else tuple(float(v) for v in hybrid))
    if key not in _RBF_CACHE:
        import sympy

        x, y = sympy.symbols("x y", real=True)
        r = sympy.sqrt(x**2 + y**2)
        if hybrid is None:
            phi = r ** p
        else:
            a, b, e = (sympy.Float(float(v)) for v in hybrid)
            phi = a * sympy.exp(-(e * r) ** 2) + b * r ** p
        kk = 1 if k is None else k
        fx = sympy.simplify(sympy.diff(phi, x, kk))
        fy = sympy.simplify(sympy.diff(phi, y, kk))
        _RBF_CACHE[key] = tuple(sympy.lambdify((x, y), e, "numpy") for e in (phi, fx, fy))
    return _RBF_CACHE[key]


def _poly_eval(exps, X):
    return np.stack([X[:, 0] ** a * X[:, 1] ** b for (a, b) in exps], axis=1)


def _poly_deriv_at_origin(exps, axis: int, k: int):
    """k-th derivative of each monomial w.r.t. `axis`, evaluated at the (shifted) centre 0."""
    out = np.zeros(len(exps))
    for j, (a, b) in enumerate(exps):
        e_ax, e_other = (a, b) if axis == 0 else (b, a)
        if e_ax == k and e_other == 0:
            out[j] = math.factorial(k)
    return out


def compute_flux_operator(points, neighbors, p: int, N: int, k: int | None = None, hybrid=None):
    """compute_flux_operator, compute_operators.jl:409-453 (k=None) and :549-594 (k-th derivative).
    Per point: shift_stencil (:225-246), rbf_block/poly_block (:191-223), Symmetric [R P; P' 0] (:265-267),
    rhs from mirror_stencil (:248-263) with the centre at (eps,eps), `M \\ rhs` (Bunch-Kaufman; scipy
    assume_a='sym' uses the same LAPACK driver), rescale by the per-axis factor^k.
    Returns [Dx, Dy] as scipy CSC with explicit zeros kept and sorted row indices (= Julia sparse(I,J,V))."""
    npts, nv = neighbors.shape
    exps = _monomial_exponents(N)
    npoly = len(exps)
    phi, phi_x, phi_y = _phs_functions(p, k, hybrid)
    kk = 1 if k is None else k
    Dx_loc = np.zeros((npts, nv))
    Dy_loc = np.zeros((npts, nv))
    eps = np.finfo(np.float64).eps
    pr_x = _poly_deriv_at_origin(exps, 0, kk)
    pr_y = _poly_deriv_at_origin(exps, 1, kk)
    pr_0 = np.array([1.0 if (a, b) == (0, 0) else 0.0 for (a, b) in exps])
    for e in range(npts):
        X = points[neighbors[e]]
        Xs = X - X[0]
        s = 1.0 / np.abs(Xs).max(axis=0)
        Xs = Xs * s
        dX = Xs[:, None, :] - Xs[None, :, :]
        with np.errstate(invalid="ignore", divide="ignore"):
            R = phi(dX[..., 0], dX[..., 1])
        R = np.where(np.isfinite(R), R, 0.0)
        P = _poly_eval(exps, Xs)
        M = np.zeros((nv + npoly, nv + npoly))
        M[:nv, :nv] = R
        M[:nv, nv:] = P
        M[nv:, :nv] = P.T
        Xm = -Xs
        Xm[0] = eps
        rhs = np.zeros((nv + npoly, 3))
        rhs[:nv, 0] = phi_x(Xm[:, 0], Xm[:, 1])
        rhs[:nv, 1] = phi_y(Xm[:, 0], Xm[:, 1])
        rhs[:nv, 2] = phi(Xm[:, 0], Xm[:, 1])
        rhs[nv:, 0] = pr_x
        rhs[nv:, 1] = pr_y
        rhs[nv:, 2] = pr_0
        W = scipy.linalg.solve(np.triu(M) + np.triu(M, 1).T, rhs, assume_a="sym")
        Dx_loc[e] = s[0] ** kk * W[:nv, 0]
        Dy_loc[e] = s[1] ** kk * W[:nv, 1]
    rows = np.repeat(np.arange(npts), nv)
    cols = neighbors.reshape(-1)
    out = []
    for loc in (Dx_loc, Dy_loc):
        A = sp.coo_matrix((loc.reshape(-1), (rows, cols)), shape=(npts, npts)).tocsc()
        A.sort_indices()
        out.append(A)
    return out


# --------------------------------------------------------------------------------------------
# ctypes mirror of the C structs
# --------------------------------------------------------------------------------------------
class _CSC(C.Structure):
    _fields_ = [("colptr", C.c_void_p), ("rowval", C.c_void_p), ("nzval", C.c_void_p)]


class _BC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad_", C.c_int32), ("n", C.c_int64), ("idx", C.c_void_p),
                ("normals", C.c_void_p), ("values", C.c_void_p)]


class _SRC(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mean_divisor_vn", C.c_int32), ("max_lexicographic", C.c_int32),
                ("pad_", C.c_int32), ("hv", _CSC), ("gamma", C.c_double), ("c_rv", C.c_double),
                ("c_uw", C.c_double), ("dx_avg", C.c_double), ("success_iter", C.c_int64),
                ("eps_uw", C.c_void_p), ("eps_rv", C.c_void_p), ("eps", C.c_void_p), ("eps_c", C.c_void_p),
                ("residual", C.c_void_p), ("approx_du", C.c_void_p),
                ("igr_alpha", C.c_double), ("igr_maxiter", C.c_int64), ("igr_sigma", C.c_void_p), ("igr_work", C.c_void_p),
                ("igr_iters", C.c_int64), ("igr_res", C.c_double)]


class _PROBLEM(C.Structure):
    _fields_ = [("n", C.c_int64), ("nvars", C.c_int32), ("eq", C.c_int32), ("eqp", C.c_double * 2),
                ("D", _CSC * 2), ("nbc", C.c_int32), ("nsrc", C.c_int32), ("bcs", C.c_void_p),
                ("srcs", C.c_void_p), ("scratch_a", C.c_void_p), ("scratch_b", C.c_void_p)]


def _ptr(a):
    return a.ctypes.data if a is not None else None


class JuliaCSC:
    """Arrays exactly as Julia's SparseMatrixCSC{Float64,Int64} stores them (1-based Int64)."""

    def __init__(self, A):
        A = sp.csc_matrix(A)
        A.sort_indices()
        self.shape = A.shape
        self.colptr = (A.indptr.astype(np.int64) + 1)
        self.rowval = (A.indices.astype(np.int64) + 1)
        self.nzval = np.ascontiguousarray(A.data, dtype=np.float64)
        self.scipy = A

    def cstruct(self):
        return _CSC(_ptr(self.colptr), _ptr(self.rowval), _ptr(self.nzval))


@dataclass
class OracleBC:
    kind: int
    idx: np.ndarray                      # 0-based point indices
    normals: np.ndarray                  # (n,2)
    value_fn: object = None              # callable (x(n,2), t) -> (V,n) for Dirichlet
    values: np.ndarray | None = None


@dataclass
class OracleSource:
    kind: int
    hv: JuliaCSC | None = None
    gamma: float = 0.0
    c_rv: float = 1.0
    c_uw: float = 1.0
    dx_avg: float = 0.0
    polydeg: int = 4
    mean_divisor_vn: bool = True
    max_lexicographic: bool = True
    # caches (create_tominec_rv_cache, hyperviscosity.jl:202-244)
    arrays: dict = field(default_factory=dict)
    success_iter: int = 0
    igr_alpha: float = 1.0
    igr_maxiter: int = 20


class OracleProblem:
    """Holds everything `rhs!` needs and drives the C restatement."""

    def __init__(self, points, nvars, eq, eqp, Dx, Dy, bcs=(), sources=()):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.n = self.points.shape[0]
        self.nvars = nvars
        self.eq = eq
        self.eqp = list(eqp) + [0.0] * (2 - len(eqp))
        self.D = [Dx if isinstance(Dx, JuliaCSC) else JuliaCSC(Dx), Dy if isinstance(Dy, JuliaCSC) else JuliaCSC(Dy)]
        self.bcs = list(bcs)
        self.sources = list(sources)
        n, V = self.n, nvars
        self.scratch_a = np.zeros((V, n))
        self.scratch_b = np.zeros((V, n))
        for s in self.sources:
            if s.kind in (SRC_UPWIND, SRC_RESIDUAL) and not s.arrays:
                s.arrays = dict(eps_uw=np.zeros(n), eps_rv=np.zeros(n), eps=np.zeros(n),
                                eps_c=np.zeros(n, dtype=np.int64), residual=np.zeros((V, n)),
                                approx_du=np.zeros((V, n)), time_history=np.zeros(s.polydeg + 1),
                                time_weights=np.zeros(s.polydeg + 1),
                                sol_history=np.zeros((V, s.polydeg + 1, n)))
            if s.kind == SRC_IGR and not s.arrays:
                s.arrays = dict(sigma=np.zeros(n), igr_work=np.zeros(12 * n))
        self._keep = []

    # -- struct assembly -------------------------------------------------------------------
    def _bc_idx1(self, bc):
        return np.ascontiguousarray(bc.idx.astype(np.int64) + 1)

    def _cproblem(self, t):
        keep = []
        bcs = (_BC * max(1, len(self.bcs)))()
        for g, bc in enumerate(self.bcs):
            idx1 = self._bc_idx1(bc)
            nrm = np.ascontiguousarray(bc.normals, dtype=np.float64)
            vals = None
            if bc.kind == BC_DIRICHLET:
                if bc.value_fn is not None:
                    vals = np.ascontiguousarray(bc.value_fn(self.points[bc.idx], t), dtype=np.float64)
                else:
                    vals = np.ascontiguousarray(bc.values, dtype=np.float64)
                assert vals.shape == (self.nvars, len(bc.idx))
            keep += [idx1, nrm, vals]
            bcs[g] = _BC(bc.kind, 0, len(bc.idx), _ptr(idx1), _ptr(nrm), _ptr(vals))
        srcs = (_SRC * max(1, len(self.sources)))()
        for i, s in enumerate(self.sources):
            a = s.arrays
            hv = s.hv.cstruct() if s.hv is not None else _CSC(None, None, None)
            srcs[i] = _SRC(s.kind, int(s.mean_divisor_vn), int(s.max_lexicographic), 0, hv, s.gamma, s.c_rv,
                           s.c_uw, s.dx_avg, s.success_iter, _ptr(a.get("eps_uw")), _ptr(a.get("eps_rv")),
                           _ptr(a.get("eps")), _ptr(a.get("eps_c")), _ptr(a.get("residual")),
                           _ptr(a.get("approx_du")), s.igr_alpha, s.igr_maxiter, _ptr(a.get("sigma")), _ptr(a.get("igr_work")),
                           0, 0.0)
        P = _PROBLEM()
        P.n, P.nvars, P.eq = self.n, self.nvars, self.eq
        P.eqp[0], P.eqp[1] = self.eqp[0], self.eqp[1]
        P.D[0], P.D[1] = self.D[0].cstruct(), self.D[1].cstruct()
        P.nbc, P.nsrc = len(self.bcs), len(self.sources)
        P.bcs = C.cast(bcs, C.c_void_p)
        P.srcs = C.cast(srcs, C.c_void_p)
        P.scratch_a, P.scratch_b = _ptr(self.scratch_a), _ptr(self.scratch_b)
        self._keep = keep + [bcs, srcs]
        return P

    # -- reference entry points --------------------------------------------------------------
    def rhs(self, u, t=0.0):
        """Trixi.rhs! (rbfsolver.jl:397-428).  u: (V,N) float64 C-contiguous, MUTATED.  Returns du."""
        assert u.flags.c_contiguous and u.shape == (self.nvars, self.n)
        du = np.empty_like(u)
        P = self._cproblem(t)
        lib().orc_rhs(C.byref(P), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(du)))
        self._collect(P)
        return du

    def _collect(self, P):
        """scalar outputs the C side leaves in the source structs (IGR: CG iteration count, final residual)"""
        srcs = C.cast(P.srcs, C.POINTER(_SRC))
        for i, s in enumerate(self.sources):
            if s.kind == SRC_IGR:
                s.arrays["iters"], s.arrays["res"] = int(srcs[i].igr_iters), float(srcs[i].igr_res)

    def rhs_repeat(self, u, reps, t=0.0):
        du = np.empty_like(u)
        P = self._cproblem(t)
        lib().orc_rhs_repeat(C.byref(P), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(du)), C.c_int(reps))
        return du

    def calc_fluxes(self, u, du):
        """calc_fluxes! (rbfsolver.jl:247-265): du += -Dx F(u) - Dy G(u)"""
        P = self._cproblem(0.0)
        lib().orc_calc_fluxes(C.byref(P), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(du)))

    def apply_source(self, i, u, du, t=0.0):
        """source(du,u,t,...) functor call (hyperviscosity.jl:52-64,121-134,351-409)"""
        P = self._cproblem(t)
        srcs = C.cast(P.srcs, C.POINTER(_SRC))
        lib().orc_apply_source(C.byref(P), C.byref(srcs[i]), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(du)))
        self._collect(P)

    def residual_norms(self, i, u, du, t=0.0):
        """n_inf_norms of update_residual_visc! (hyperviscosity.jl:305-311) for source i; also refreshes eps_rv."""
        P = self._cproblem(t)
        srcs = C.cast(P.srcs, C.POINTER(_SRC))
        out = np.zeros(self.nvars)
        lib().orc_update_residual_visc(C.byref(P), C.byref(srcs[i]), C.c_void_p(_ptr(du)), C.c_void_p(_ptr(u)),
                                       C.c_void_p(_ptr(out)))
        return out

    def boundary_pass(self, u, du, t=0.0):
        P = self._cproblem(t)
        lib().orc_boundary_pass(C.byref(P), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(du)))

    def flux(self, u, dim):
        f = np.empty_like(u)
        P = self._cproblem(0.0)
        lib().orc_flux(C.byref(P), C.c_int(dim), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(f)))
        return f

    # -- HistoryCallback (history.jl:62-129) ---------------------------------------------------
    def history_callback(self, u, t, success_iter, approx_order):
        for s in self.sources:
            if s.kind != SRC_RESIDUAL:
                continue
            a = s.arrays
            s.success_iter = int(success_iter)
            nslots = s.polydeg + 1
            V, n = self.nvars, self.n
            L = lib()
            L.orc_shift_soln_history(C.c_int64(n), C.c_int(V), C.c_int(nslots), C.c_void_p(_ptr(a["time_history"])),
                                     C.c_void_p(_ptr(a["sol_history"])), C.c_double(t), C.c_void_p(_ptr(u)))
            L.orc_update_approx_du(C.c_int64(n), C.c_int(V), C.c_int(nslots), C.c_void_p(_ptr(a["approx_du"])),
                                   C.c_void_p(_ptr(a["time_weights"])), C.c_void_p(_ptr(a["time_history"])),
                                   C.c_void_p(_ptr(a["sol_history"])), C.c_int64(int(success_iter)),
                                   C.c_int(int(approx_order)))

    # -- SSPRK33 with FSAL, callbacks after each accepted step ---------
    def solve_ssprk33(self, u0, t0, dt, nsteps, approx_order=None):
        """Fixed-dt SSPRK33 in the OrdinaryDiffEq in-place/FSAL structure: k=f(u_n) is carried over from
        the end of the previous step; HistoryCallback runs at init and after every step."""
        L = lib()
        u = np.ascontiguousarray(u0, dtype=np.float64).copy()
        t = float(t0)
        success_iter = 0
        if approx_order is not None:
            self.history_callback(u, t, success_iter, approx_order)  # initialize! history.jl:51-54
        k = self.rhs(u, t)  # FSAL first
        n_el = u.size
        for _ in range(nsteps):
            uprev = u.copy()
            L.orc_ssprk33_stage(C.c_int64(n_el), 1, C.c_double(dt), C.c_void_p(_ptr(uprev)), C.c_void_p(_ptr(k)),
                                C.c_void_p(_ptr(u)))
            k = self.rhs(u, t + dt)
            L.orc_ssprk33_stage(C.c_int64(n_el), 2, C.c_double(dt), C.c_void_p(_ptr(uprev)), C.c_void_p(_ptr(k)),
                                C.c_void_p(_ptr(u)))
            k = self.rhs(u, t + dt / 2)
            L.orc_ssprk33_stage(C.c_int64(n_el), 3, C.c_double(dt), C.c_void_p(_ptr(uprev)), C.c_void_p(_ptr(k)),
                                C.c_void_p(_ptr(u)))
            k = self.rhs(u, t + dt)
            t = t + dt
            success_iter += 1
            if approx_order is not None:
                self.history_callback(u, t, success_iter, approx_order)
        return u, t


class PIController:
    """OrdinaryDiffEq's default PI step-size controller (third party, unpinned) for an order-3 method:
    beta1 = 7/(10k), beta2 = 2/(5k), gamma = 0.9, qmin = 0.2, qmax = 10, qsteady in [1, 1.2], qoldinit = 1e-4.
    Used on BOTH sides of the SSPRK43 parity tests; in the drop-in deployment the controller stays inside OrdinaryDiffEq
    and the library only supplies the error estimate."""

    def __init__(self, order=3, gamma=0.9, qmin=0.2, qmax=10.0, qsteady_min=1.0, qsteady_max=1.2, qoldinit=1e-4):
        self.beta1, self.beta2 = 7.0 / (10.0 * order), 2.0 / (5.0 * order)
        self.gamma, self.qmin, self.qmax = gamma, qmin, qmax
        self.qsteady_min, self.qsteady_max, self.qoldinit = qsteady_min, qsteady_max, qoldinit
        self.qold = qoldinit

    def propose(self, dt, eest):
        """returns (accept, dt_next)"""
        if eest == 0.0:
            q, q11 = 1.0 / self.qmax, 0.0
        else:
            q11 = eest ** self.beta1
            q = q11 / self.qold ** self.beta2
            q = max(1.0 / self.qmax, min(1.0 / self.qmin, q / self.gamma))
        if eest <= 1.0:                                   # step_accept_controller!
            if self.qsteady_min <= q <= self.qsteady_max:
                q = 1.0
            self.qold = max(eest, self.qoldinit)
            return True, dt / q
        return False, dt / min(1.0 / self.qmin, q11 / self.gamma)   # step_reject_controller!


def _ssprk43_step(P, u, k, t, dt, abstol, reltol):
    """one SSPRK43 step from (u, k=f(u,t)); returns (u_new, k_new=f(u_new,t+dt), EEst).  u, k are not modified."""
    L = lib()
    L.orc_error_sumsq.restype = C.c_double
    uprev = u
    un = u.copy()
    ut = np.zeros_like(u)
    n_el = u.size
    args = lambda st, kk: (C.c_int64(n_el), st, C.c_double(dt), C.c_void_p(_ptr(uprev)), C.c_void_p(_ptr(kk)),
                           C.c_void_p(_ptr(un)), C.c_void_p(_ptr(ut)))
    L.orc_ssprk43_stage(*args(1, k))
    kk = P.rhs(un, t + dt / 2)
    L.orc_ssprk43_stage(*args(2, kk))
    kk = P.rhs(un, t + dt)
    L.orc_ssprk43_stage(*args(3, kk))
    kk = P.rhs(un, t + dt / 2)
    L.orc_ssprk43_stage(*args(4, kk))
    ss = L.orc_error_sumsq(C.c_int64(n_el), C.c_void_p(_ptr(ut)), C.c_void_p(_ptr(uprev)), C.c_void_p(_ptr(un)),
                           C.c_double(abstol), C.c_double(reltol))
    eest = math.sqrt(ss / n_el)
    kk = P.rhs(un, t + dt)
    return un, kk, eest


def solve_ssprk43(P, u0, t0, t1, dt0, abstol=1e-8, reltol=1e-8, approx_order=None, max_steps=10000):
    """Adaptive SSPRK43 with FSAL, HistoryCallback after every ACCEPTED step (rbfsolver_test.jl:104-107)."""
    u = np.ascontiguousarray(u0, dtype=np.float64).copy()
    t, dt = float(t0), float(dt0)
    ctrl = PIController()
    success_iter = 0
    if approx_order is not None:
        P.history_callback(u, t, 0, approx_order)
    k = P.rhs(u, t)
    log = []
    for _ in range(max_steps):
        if t >= t1 - 1e-14 * max(1.0, abs(t1)):
            break
        dt = min(dt, t1 - t)
        un, kn, eest = _ssprk43_step(P, u, k, t, dt, abstol, reltol)
        accept, dt_next = ctrl.propose(dt, eest)
        log.append((t, dt, eest, accept))
        if accept:
            u, k, t = un, kn, t + dt
            success_iter += 1
            if approx_order is not None:
                P.history_callback(u, t, success_iter, approx_order)
        dt = dt_next
    return u, t, log


VAR_DENSITY, VAR_PRESSURE = 0, 1


def limiter_zhang_shu(u, neighbors, thresholds, variables, gamma):
    """PositivityPreservingLimiterZhangShu(thresholds, variables)(u, ...) (positivity_zhang_shu.jl:29-72 +
    positivity_zhang_shu_point2d.jl:22-82): in place on u (4,N); neighbors (N,k) 0-based in kNN list order"""
    u = np.ascontiguousarray(u)
    nb = np.ascontiguousarray(neighbors, dtype=np.int64)
    n, k = nb.shape
    s1, s2 = np.zeros_like(u), np.zeros_like(u)
    for thr, var in zip(thresholds, variables):
        lib().orc_limiter_zhang_shu(C.c_int64(n), C.c_int(k), C.c_void_p(_ptr(nb)), C.c_double(gamma), C.c_double(thr),
                                    C.c_int(var), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(s1)), C.c_void_p(_ptr(s2)))
    return u


def time_deriv_weights(t):
    t = np.ascontiguousarray(t, dtype=np.float64)
    w = np.zeros_like(t)
    lib().orc_time_deriv_weights(C.c_int(len(t)), C.c_void_p(_ptr(t)), C.c_void_p(_ptr(w)))
    return w


def sum_pairwise(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    f = lib().orc_sum
    f.restype = C.c_double
    return f(C.c_void_p(_ptr(a)), C.c_int64(a.size))


# --------------------------------------------------------------------------------------------
# Source constructors (argument meaning as in src/sources/hyperviscosity.jl)
# --------------------------------------------------------------------------------------------
def source_hyperviscosity_flyer(points, neighbors, p, N, dx_min, k=2, c=1.0):
    """create_flyer_hv_cache hyperviscosity.jl:34-50: H = sum(compute_flux_operator(.., 2k)), gamma = c dx_min^(2k)"""
    ops = compute_flux_operator(points, neighbors, p, N, 2 * k)
    H = sp.csc_matrix(ops[0] + ops[1])
    return OracleSource(kind=SRC_HV_FLYER, hv=JuliaCSC(H), gamma=c * dx_min ** (2 * k))


def source_hyperviscosity_tominec(points, neighbors, p, N, dx_min, c=1.0):
    """create_tominec_hv_cache hyperviscosity.jl:101-119: lap = sum(ops(2)); H = lap' * lap; gamma = c dx_min^4.5"""
    ops = compute_flux_operator(points, neighbors, p, N, 2)
    lap = sp.csc_matrix(ops[0] + ops[1])
    H = sp.csc_matrix(lap.T @ lap)
    return OracleSource(kind=SRC_HV_TOMINEC, hv=JuliaCSC(H), gamma=c * dx_min ** 4.5)


def source_igr(alpha=1.0, maxiter=20):
    """SourceIGR(solver, equations, domain; alpha, linear_solver = cg!) IGR.jl:32-36; maxiter = 20 is hard-wired (:190)"""
    return OracleSource(kind=SRC_IGR, igr_alpha=alpha, igr_maxiter=maxiter)


def source_upwind(dx_avg, c_uw=1.0, polydeg=4):
    return OracleSource(kind=SRC_UPWIND, c_rv=1.0, c_uw=c_uw, dx_avg=dx_avg, polydeg=polydeg)


def source_residual(dx_avg, c_rv=1.0, c_uw=1.0, polydeg=4, **kw):
    return OracleSource(kind=SRC_RESIDUAL, c_rv=c_rv, c_uw=c_uw, dx_avg=dx_avg, polydeg=polydeg, **kw)
