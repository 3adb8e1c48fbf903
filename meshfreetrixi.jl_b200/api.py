"""Host-side mirror of the reference's Trixi-style interface for the `rhs!` path.

Same names, argument meaning and call order as MeshfreeTrixi.jl so that tests read like the reference's own
(test/divergence_test.jl, test/upwind_viscosity_test.jl, test/history_test.jl):

    basis  = PointCloudBasis(Point2D(), 3, approximation_type=RBF(PolyharmonicSpline(3)))
    solver = PointCloudSolver(basis, engine=RBFFDEngineCUDA())
    domain = PointCloudDomain(solver, "data/cyl_0_05", dict(inlet=1, outlet=2, bottom=3, top=4, cyl=5))
    semi   = SemidiscretizationHyperbolic(domain, equations, ic, solver, boundary_conditions=..., source_terms=...)
    ode    = semidiscretize(semi, (0.0, 0.5));  rhs_(du, ode.u0, semi, t)

Everything numerical on the path is executed by libmft_b200.so through its C ABI (see _lib.py); this module only
marshals arrays.  State arrays are (V, N) float64 C-contiguous = the V component vectors of the reference's
StructArray (allocate_nested_array, src/solvers/pointcloudsolver/rbfsolver.jl:111-116).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from . import cloud as cloudmod
from . import setup_ops


# ---- basis / solver / engine (src/solvers/rbfsolver.jl:8-71, src/solvers/pointcloudsolver/types.jl:71-86) ----
class Point2D:
    ndims = 2


@dataclass
class PolyharmonicSpline:
    Nrbf: int = 3


@dataclass
class HybridGaussianPHS:
    """blended Gaussian and odd-order polyharmonic spline phi = alpha exp(-(epsilon r)^2) + beta r^Nrbf
    (geometry_primatives.jl:117-132, 238-262)"""
    Nrbf: int = 3
    alpha: float = 1.0
    beta: float = 1.0
    epsilon: float = 1.0


@dataclass
class RBF:
    rbf_type: object = field(default_factory=PolyharmonicSpline)   # PolyharmonicSpline | HybridGaussianPHS


def _hybrid_of(basis):
    """(alpha, beta, epsilon) of a HybridGaussianPHS basis, None for a pure polyharmonic spline"""
    t = basis.approx_type.rbf_type
    return (t.alpha, t.beta, t.epsilon) if isinstance(t, HybridGaussianPHS) else None


@dataclass
class RefPointData:
    elem: object
    approx_type: RBF
    N: int
    nv: int


def PointCloudBasis(element_type, polydeg, approximation_type=None, nv=None):
    """types.jl:83-86 -> RefPointData (geometry_primatives.jl:189-201).  `nv` overrides the stencil width
    (RefPointData is a plain struct in the reference; the stencil sweep of BASELINE.json overrides it)."""
    approximation_type = approximation_type or RBF()
    return RefPointData(element_type, approximation_type, polydeg, nv or setup_ops.num_neighbors(polydeg, 2))


@dataclass
class RBFFDEngineCUDA:
    """The reference's reserved, method-less plug-in point (src/solvers/rbfsolver.jl:60-61), given a body."""
    device: int = 0
    reorder: str | None = "hilbert"   # space-filling-curve device ordering
    exact_order: bool = True           # reference summation order, separate mul/add (bit-identical sums)
    diagnostics: bool = False          # keep eps/eps_uw/eps_rv/residual on device for inspection
    mean_divisor_vn: bool = True       # ode_mean divides by V*N (recursive_length)
    max_lexicographic: bool = True     # maximum(::StructArray{SVector}) is a lexicographic max
    stage_weights: int = 5             # bit0: pass A, bit1: pass B -- bulk-copy whole operator slices to smem
    cuda_graph: int = 1                # 1: graph replay of SSPRK steps on one GPU; 2: also multi-rank; 0: eager
    exchange: str = "p2p"              # multi-GPU halo exchange: "p2p" (CUDA-IPC peer memory over NVLink) or "nccl"
    pair_rows: int = 1                 # row-pair (union stencil) layout: bit0 transposed operator (pass B), bit1 forward (pass A)
    tile: int = 31                     # union-tile kernels: bit0 pass A, bit1 pass B, bit2 bank-coloured slots, bit3 two record copies,
                                       # bit4 (31, default since r2: -13 % shared wavefronts measured): second copy tuned by local search
    tile_rows: int = 11                # rows per thread of the union-tile kernels: units digit pass A, tens digit pass B (1, 2, 4)
    pdl: bool = True                   # programmatic dependent launch between the kernels of a fused stage
    fused_step: bool = True            # mft_ssprk_step: one fused stage kernel (stage update + BCs + norms + u halo puts), halo waits inside the pass kernels
    prefetch_distance: int | None = None  # slices ahead for the L2 prefetch of operator data (None: library default, 0: off)
    single_sweep_exact: bool = False   # k=20: exact-order pass A in one sweep (y-products parked in registers)
    refine_order: bool = False         # order the rows inside a tile by D' row length (fewer padding steps in pass B; opt-in)
    setup: str = "host"                # "device": kNN + RBF-FD weight solves on the GPU (mft_setup_knn / mft_setup_rbf_weights)
    layout_device: bool = True         # union-tile layouts built on the GPU at finalize (False: host threads; same bytes)


@dataclass
class RBFSolver:
    basis: RefPointData
    engine: object


def PointCloudSolver(basis, engine=None):
    return RBFSolver(basis, engine or RBFFDEngineCUDA())


# ---- domain (geometry_primatives.jl:297-364, SerialPointCloud.jl:12-54) ------------------------------------------
@dataclass
class PointData:
    points: np.ndarray
    neighbors: np.ndarray  # (N,nv) 0-based, self first
    num_points: int
    num_neighbors: int
    dx_min: float
    dx_avg: float


@dataclass
class BoundaryData:
    idx: np.ndarray      # 0-based
    normals: np.ndarray  # (n,2)


class PointCloudDomain:
    def __init__(self, solver, source, boundary_names):
        """source: Medusa case name (file-based ctor, types.jl:133-145) or a cloud.Cloud; boundary_names maps a
        name to the 1-based boundary group number, as in the reference tests."""
        cl = cloudmod.read_medusa_file(source) if isinstance(source, str) else source
        nv = solver.basis.nv
        nb, dx_min, dx_avg = setup_ops.knn_with(solver.engine, cl.points, nv)
        self.pd = PointData(np.ascontiguousarray(cl.points, dtype=np.float64), nb, cl.points.shape[0], nv, dx_min, dx_avg)
        self.boundary_tags = {name: BoundaryData(np.asarray(cl.boundary_idxs[g - 1], dtype=np.int64),
                                                 np.asarray(cl.boundary_normals[g - 1], dtype=np.float64))
                              for name, g in boundary_names.items()}
        self.cloud = cl


class ParallelPointCloudDomain(PointCloudDomain):
    """One rank's part of the cloud: [owned points ; halo points] (layout of the reference's
    ParallelPointCloudDomain, src/domains/PointCloudDomain/ParallelPointCloud.jl:85-162), but cut along a
    space-filling curve and with full global stencils for the owned rows (see partition.py)."""

    def __init__(self, solver, source, boundary_names, comm, wide_halo=False):
        """wide_halo=True keeps a second stencil ring of column-only halo points: needed by sources whose operator is a
        product of two stencil operators (SourceHyperviscosityTominec: L'L)"""
        from . import partition

        cl = cloudmod.read_medusa_file(source) if isinstance(source, str) else source
        basis = solver.basis
        eng = solver.engine
        on_device = getattr(eng, "setup", "host") == "device"
        hyb = _hybrid_of(basis)
        part = partition.build_rank_partition(
            cl.points, cl.boundary_idxs, cl.boundary_normals, comm.rank, comm.nranks, basis.approx_type.rbf_type.Nrbf,
            basis.N, basis.nv, comm.allgather,
            knn_queries=(lambda pts, q, nv: setup_ops.knn_queries_device(pts, q, nv, eng.device)) if on_device else None,
            weights_rows=(lambda pts, rows, p, N: setup_ops.rbf_fd_weights_rows_device(pts, rows, p, N, None, eng.device, hyb))
            if on_device else ((lambda pts, rows, p, N: setup_ops.rbf_fd_weights(pts, rows, p, N, hybrid=hyb)) if hyb else None),
            wide_halo=wide_halo)
        self.partition, self.comm, self.cloud = part, comm, cl
        self.pd = PointData(part.points, part.neighbors_owned, part.n_local + part.n_halo, basis.nv, part.dx_min, part.dx_avg)
        self.boundary_tags = {name: BoundaryData(part.boundary_idxs[g - 1], part.boundary_normals[g - 1])
                              for name, g in boundary_names.items()}


# ---- equations (Trixi, third party) ------------------------------------------------------------------------------
@dataclass
class CompressibleEulerEquations2D:
    gamma: float
    nvars = 4
    kind = L.EQ_EULER2D

    def params(self):
        return [self.gamma]


@dataclass
class LinearScalarAdvectionEquation2D:
    a1: float
    a2: float
    nvars = 1
    kind = L.EQ_ADVECTION2D

    def params(self):
        return [self.a1, self.a2]


# ---- boundary conditions (src/equations/PointCloudBCs.jl) ------------------------------------------------------
@dataclass
class BoundaryConditionDirichlet:
    boundary_value_function: object   # f(x (n,2), t, equations) -> (V,n)
    time_dependent: bool = False
    kind = L.BC_DIRICHLET


class _SlipWall:
    kind = L.BC_SLIP_WALL


boundary_condition_slip_wall = _SlipWall()


class BoundaryConditionDoNothing:
    kind = L.BC_DO_NOTHING


# ---- sources (src/sources/hyperviscosity.jl, generic_sources.jl) --------------------------------------------------
class _Source:
    index = None   # position in SourceTerms once attached to a semidiscretization
    semi = None

    def __call__(self, du, u, t, semi=None):
        """source(du,u,t,domain,equations,solver,cache): du is accumulated into."""
        semi = semi or self.semi
        L.check(L.load().mft_apply_source(semi.ctx, self.index, float(t), L.soa_ptrs(u), L.soa_ptrs(du)))

    @property
    def cache(self):
        return _SourceCacheView(self)


class _SourceCacheView:
    def __init__(self, src):
        self._s = src

    def _scalar(self, fld):
        semi = self._s.semi
        out = np.empty(semi.n)
        L.check(L.load().mft_get_field(semi.ctx, fld, L.ptr(out)))
        return out

    eps = property(lambda self: self._scalar(L.FIELD_EPS))
    eps_uw = property(lambda self: self._scalar(L.FIELD_EPS_UW))
    eps_rv = property(lambda self: self._scalar(L.FIELD_EPS_RV))
    eps_c = property(lambda self: self._scalar(L.FIELD_EPS_C))

    def _vec(self, fld):
        semi = self._s.semi
        out = np.empty((semi.V, semi.n))
        L.check(L.load().mft_get_field(semi.ctx, fld, L.ptr(out)))
        return out

    residual = property(lambda self: self._vec(L.FIELD_RESIDUAL))
    approx_du = property(lambda self: self._vec(L.FIELD_APPROX_DU))

    sigma = property(lambda self: self._scalar(L.FIELD_SIGMA))

    @property
    def igr_status(self):
        """(CG iterations of the last solve, final |r|, initial |r|)"""
        out = np.empty(3)
        L.check(L.load().mft_get_field(self._s.semi.ctx, L.FIELD_IGR_STATUS, L.ptr(out)))
        return int(out[0]), float(out[1]), float(out[2])

    @property
    def norms(self):
        out = np.empty(self._s.semi.V)
        L.check(L.load().mft_get_field(self._s.semi.ctx, L.FIELD_NORMS, L.ptr(out)))
        return out


class SourceHyperviscosityFlyer(_Source):
    """hyperviscosity.jl:14-64: H = sum_d d^{2k}/dx_d^{2k}, gamma = c*dx_min^{2k}."""
    kind = L.SRC_HV_FLYER

    def __init__(self, solver, equations, domain, k=2, c=1.0):
        p, N = solver.basis.approx_type.rbf_type.Nrbf, solver.basis.N
        part = getattr(domain, "partition", None)
        if part is None:
            ops = setup_ops.flux_operator_with(solver.engine, domain.pd.points, domain.pd.neighbors, p, N, 2 * k, _hybrid_of(solver.basis))
            self.hv_differentiation_matrix = (ops[0] + ops[1]).tocsc()
        else:
            # one rank of a partitioned cloud: H has the sparsity of D, so the owned rows need exactly the u halo that
            # rhs! already exchanges.  Weights from the GLOBAL stencils of the owned rows; columns
            # renumbered to the local [owned ; halo] layout; halo rows stay empty (they are not computed here).
            eng = solver.engine
            gpts = np.ascontiguousarray(domain.cloud.points, dtype=np.float64)
            hyb = _hybrid_of(solver.basis)
            if getattr(eng, "setup", "host") == "device":
                wx, wy = setup_ops.rbf_fd_weights_rows_device(gpts, part.neighbors_owned, p, N, 2 * k, eng.device, hyb)
            else:
                wx, wy = setup_ops.rbf_fd_weights(gpts, part.neighbors_owned, p, N, 2 * k, hybrid=hyb)
            lut = np.full(part.n_global, -1, dtype=np.int64)
            lut[part.local_gid] = np.arange(part.n_local + part.n_halo)
            cols = lut[part.neighbors_owned]
            assert (cols >= 0).all()
            n_tot = part.n_local + part.n_halo
            rows = np.repeat(np.arange(part.n_local, dtype=np.int64), part.neighbors_owned.shape[1])
            H = sp.coo_matrix(((wx + wy).reshape(-1), (rows, cols.reshape(-1))), shape=(n_tot, n_tot)).tocsc()
            H.sort_indices()
            self.hv_differentiation_matrix = H
        self.gamma = c * domain.pd.dx_min ** (2 * k)
        self.c = c


class SourceHyperviscosityTominec(_Source):
    """hyperviscosity.jl:80-134: H = L'L with L = dxx + dyy, gamma = c*dx_min^4.5."""
    kind = L.SRC_HV_TOMINEC

    def __init__(self, solver, equations, domain, c=1.0):
        p, N = solver.basis.approx_type.rbf_type.Nrbf, solver.basis.N
        part = getattr(domain, "partition", None)
        if part is None:
            ops = setup_ops.flux_operator_with(solver.engine, domain.pd.points, domain.pd.neighbors, p, N, 2, _hybrid_of(solver.basis))
            lap = (ops[0] + ops[1]).tocsc()
            self.hv_differentiation_matrix = (lap.T @ lap).tocsc()
        else:
            # one rank of a partitioned cloud.  Row i of L'L is sum_k L[k,i] L[k,:] over the rows k whose stencils contain i:
            # for an owned i these are owned rows and the foreign rows R_r of the halo, and the columns are their stencils --
            # local thanks to the wide halo.  L rows from the GLOBAL stencils; halo rows of the result are dropped.
            if not part.wide_halo:
                raise ValueError("SourceHyperviscosityTominec on a partitioned cloud needs ParallelPointCloudDomain(..., "
                                 "wide_halo=True): L'L reaches two stencil rings")
            eng = solver.engine
            gpts = np.ascontiguousarray(domain.cloud.points, dtype=np.float64)
            hyb = _hybrid_of(solver.basis)
            n_loc, n_tot = part.n_local, part.n_local + part.n_halo
            is_owned_g = np.zeros(part.n_global, dtype=bool)
            is_owned_g[part.owned_gid] = True
            hnb = part.neighbors_halo
            in_R = np.zeros(part.n_halo, dtype=bool)
            if part.n_halo:
                has = hnb[:, 0] >= 0
                in_R[has] = is_owned_g[hnb[has]].any(axis=1)
            rows_local = np.concatenate([np.arange(n_loc, dtype=np.int64), n_loc + np.nonzero(in_R)[0]])
            rows_nb = np.concatenate([part.neighbors_owned, hnb[in_R]]) if in_R.any() else part.neighbors_owned
            if getattr(eng, "setup", "host") == "device":
                wx, wy = setup_ops.rbf_fd_weights_rows_device(gpts, rows_nb, p, N, 2, eng.device, hyb)
            else:
                wx, wy = setup_ops.rbf_fd_weights(gpts, rows_nb, p, N, 2, hybrid=hyb)
            lut = np.full(part.n_global, -1, dtype=np.int64)
            lut[part.local_gid] = np.arange(n_tot)
            cols = lut[rows_nb]
            assert (cols >= 0).all(), "a stencil column of an owned / R row is not local (wide halo incomplete)"
            rr = np.repeat(rows_local, rows_nb.shape[1])
            lap = sp.coo_matrix(((wx + wy).reshape(-1), (rr, cols.reshape(-1))), shape=(n_tot, n_tot)).tocsc()
            H = (lap.T @ lap).tocsr()
            keep = sp.diags(np.concatenate([np.ones(n_loc), np.zeros(part.n_halo)]))
            H = (keep @ H).tocsc()
            H.eliminate_zeros()
            H.sort_indices()
            self.hv_differentiation_matrix = H
        self.gamma = c * domain.pd.dx_min ** 4.5
        self.c = c


class SourceUpwindViscosityTominec(_Source):
    """hyperviscosity.jl:148-167, apply :351-380."""
    kind = L.SRC_UPWIND

    def __init__(self, solver, equations, domain, c=1.0, c_uw=1.0, polydeg=4):
        self.c_rv, self.c_uw, self.polydeg, self.dx_avg = c, c_uw, polydeg, domain.pd.dx_avg


class SourceResidualViscosityTominec(_Source):
    """hyperviscosity.jl:181-200, apply :382-409."""
    kind = L.SRC_RESIDUAL

    def __init__(self, solver, equations, domain, c_rv=1.0, c_uw=1.0, polydeg=4):
        self.c_rv, self.c_uw, self.polydeg, self.dx_avg = c_rv, c_uw, polydeg, domain.pd.dx_avg


def cg_(*args, **kw):
    """IterativeSolvers.cg!: the linear solver SourceIGR runs on the device (the only one implemented)."""
    raise TypeError("cg_ is a tag for SourceIGR(linear_solver=cg_); the solve runs inside libmft_b200")


class SourceIGR(_Source):
    """IGR.jl:14-36: information-geometric regularisation.  Per rhs!: b = alpha (tr(Du)^2 + tr((Du)^2)) from the
    gradients of the primitive variables (:117-158), (rho^-1 - alpha (Dx rho^-1 Dx + Dy rho^-1 Dy)) sigma = b solved by
    `maxiter` (hard-wired 20, :190) conjugate-gradient iterations from sigma = 0 (:169-191), then du[2] -= Dx sigma,
    du[3] -= Dy sigma (:211-239).  `linear_solver`: the reference's default `cg` cannot take its three positional
    arguments; the in-place `cg!` (here `cg_`) is what runs."""
    kind = L.SRC_IGR

    def __init__(self, solver, equations, domain, alpha=1.0, linear_solver=cg_, maxiter=20):
        if linear_solver is not cg_:
            raise NotImplementedError("SourceIGR: only the conjugate-gradient solver (cg_) is implemented on the device")
        self.alpha, self.maxiter = float(alpha), int(maxiter)


class SourceTerms:
    """generic_sources.jl:7-18: iteration order = keyword order."""

    def __init__(self, **kwargs):
        self.sources = dict(kwargs)

    def values(self):
        return list(self.sources.values())

    def __len__(self):
        return len(self.sources)

    def __getattr__(self, name):
        try:
            return self.__dict__["sources"][name]
        except KeyError:
            raise AttributeError(name)


# ---- semidiscretization (Trixi.SemidiscretizationHyperbolic + create_cache, rbfsolver.jl:130-185) -----------------
@dataclass
class Cache:
    pd: PointData
    rbf_differentiation_matrices: list


class SemidiscretizationHyperbolic:
    def __init__(self, domain, equations, initial_condition, solver, boundary_conditions=None, source_terms=None,
                 operators=None):
        if not isinstance(solver.engine, RBFFDEngineCUDA):
            raise TypeError("this package implements the RBFFDEngineCUDA engine only (no CPU engine, no fallback)")
        self.domain, self.equations, self.initial_condition, self.solver = domain, equations, initial_condition, solver
        self.boundary_conditions = dict(boundary_conditions or {})
        self.source_terms = source_terms or SourceTerms()
        eng = solver.engine
        pd = domain.pd
        self.n, self.V = pd.num_points, equations.nvars
        p, N = solver.basis.approx_type.rbf_type.Nrbf, solver.basis.N
        part = getattr(domain, "partition", None)
        self.partition = part
        if part is not None:
            ops = part.ops
            for src in self.source_terms.values():
                if src.kind == L.SRC_HV_TOMINEC and not part.wide_halo:
                    raise ValueError("SourceHyperviscosityTominec on a partitioned cloud needs "
                                     "ParallelPointCloudDomain(..., wide_halo=True): L'L reaches two stencil rings")
        else:
            ops = operators or setup_ops.flux_operator_with(eng, pd.points, pd.neighbors, p, N, None, _hybrid_of(solver.basis))
        self.cache = Cache(pd, ops)
        lib = L.load()
        ctx = C.c_void_p()
        n_local = part.n_local if part is not None else self.n
        L.check(lib.mft_ctx_create(C.byref(ctx), eng.device, n_local, self.n - n_local, self.V, 2, pd.num_neighbors))
        self.ctx = ctx
        prm = np.asarray(equations.params(), dtype=np.float64)
        L.check(lib.mft_set_equation(ctx, equations.kind, L.ptr(prm), len(prm)))
        L.check(lib.mft_set_option(ctx, L.OPT_EXACT_ORDER, float(eng.exact_order)))
        L.check(lib.mft_set_option(ctx, L.OPT_DIAGNOSTICS, float(eng.diagnostics)))
        L.check(lib.mft_set_option(ctx, L.OPT_MEAN_DIVISOR_VN, float(eng.mean_divisor_vn)))
        L.check(lib.mft_set_option(ctx, L.OPT_MAX_LEXICOGRAPHIC, float(eng.max_lexicographic)))
        L.check(lib.mft_set_option(ctx, L.OPT_STAGE_WEIGHTS, float(eng.stage_weights)))
        L.check(lib.mft_set_option(ctx, L.OPT_REFINE_ORDER, float(eng.refine_order)))
        L.check(lib.mft_set_option(ctx, L.OPT_CUDA_GRAPH, float(eng.cuda_graph)))
        L.check(lib.mft_set_option(ctx, L.OPT_SINGLE_SWEEP_EXACT, float(eng.single_sweep_exact)))
        L.check(lib.mft_set_option(ctx, L.OPT_PAIR_ROWS, float(eng.pair_rows)))
        L.check(lib.mft_set_option(ctx, L.OPT_TILE, float(eng.tile)))
        L.check(lib.mft_set_option(ctx, L.OPT_TILE_ROWS, float(eng.tile_rows)))
        L.check(lib.mft_set_option(ctx, L.OPT_FUSED_STEP, 1.0 if eng.fused_step else 0.0))
        L.check(lib.mft_set_option(ctx, L.OPT_PDL, 1.0 if eng.pdl else 0.0))
        L.check(lib.mft_set_option(ctx, L.OPT_LAYOUT_DEVICE, 1.0 if eng.layout_device else 0.0))
        if eng.prefetch_distance is not None:
            L.check(lib.mft_set_option(ctx, L.OPT_PREFETCH_DISTANCE, float(eng.prefetch_distance)))
        if part is not None:
            # local numbering is already [owned along the curve ; halo]; sums must run in ascending GLOBAL column order
            self.perm = None
            keys = np.ascontiguousarray(part.local_gid, dtype=np.int64)
            L.check(lib.mft_set_order_keys(ctx, L.ptr(keys)))
        elif eng.reorder == "hilbert":
            self.perm = L.sfc_order(pd.points)
            perm1 = np.ascontiguousarray(self.perm + 1)
            L.check(lib.mft_set_permutation(ctx, L.ptr(perm1)))
        else:
            self.perm = None
        self._neighbors_set = False        # domain.pd.neighbors go to the device when a Zhang-Shu limiter first needs them
        for slot, A in ((L.OP_DX, ops[0]), (L.OP_DY, ops[1])):
            cp, rv, nz = setup_ops.julia_csc(A)
            L.check(lib.mft_set_operator_csc(ctx, slot, L.ptr(cp), L.ptr(rv), L.ptr(nz)))
        self._bc_groups = []
        for name, bc in self.boundary_conditions.items():
            tag = domain.boundary_tags[name]   # KeyError for unknown names, like rbfsolver.jl:300
            idx1 = np.ascontiguousarray(tag.idx + 1, dtype=np.int64)
            nrm = np.ascontiguousarray(tag.normals, dtype=np.float64)
            vals = None
            if bc.kind == L.BC_DIRICHLET:
                vals = np.ascontiguousarray(bc.boundary_value_function(pd.points[tag.idx], 0.0, equations), dtype=np.float64)
                assert vals.shape == (self.V, len(tag.idx))
            L.check(lib.mft_add_boundary(ctx, bc.kind, len(idx1), L.ptr(idx1), L.ptr(nrm), L.ptr(vals)))
            self._bc_groups.append((name, bc, tag))
        for i, src in enumerate(self.source_terms.values()):
            src.index, src.semi = i, self
            if src.kind in (L.SRC_HV_FLYER, L.SRC_HV_TOMINEC):
                prm = np.asarray([src.gamma], dtype=np.float64)
                cp, rv, nz = setup_ops.julia_csc(src.hv_differentiation_matrix)
                L.check(lib.mft_add_source(ctx, src.kind, L.ptr(prm), 1, L.ptr(cp), L.ptr(rv), L.ptr(nz)))
            elif src.kind == L.SRC_UPWIND:
                prm = np.asarray([src.c_uw, src.dx_avg], dtype=np.float64)
                L.check(lib.mft_add_source(ctx, src.kind, L.ptr(prm), 2, None, None, None))
            elif src.kind == L.SRC_IGR:
                if part is not None:
                    raise NotImplementedError("SourceIGR is single-GPU")
                prm = np.asarray([src.alpha, float(src.maxiter)], dtype=np.float64)
                L.check(lib.mft_add_source(ctx, src.kind, L.ptr(prm), 2, None, None, None))
            else:
                prm = np.asarray([src.c_rv, src.c_uw, src.dx_avg, float(src.polydeg)], dtype=np.float64)
                L.check(lib.mft_add_source(ctx, src.kind, L.ptr(prm), 4, None, None, None))
        L.check(lib.mft_finalize(ctx))
        if part is not None and part.nranks > 1:
            comm = domain.comm
            if eng.exchange == "nccl":
                uid = C.create_string_buffer(128)
                if comm.rank == 0:
                    L.check(lib.mft_nccl_unique_id(uid))
                raw = comm.broadcast_bytes(bytes(uid.raw), src=0)
                uid = C.create_string_buffer(raw, 128)
                L.check(lib.mft_comm_init(ctx, comm.nranks, comm.rank, uid))
            peers = (C.c_int * max(1, len(part.peers)))(*part.peers)
            send_off = np.zeros(len(part.peers) + 1, dtype=np.int64)
            for i, sidx in enumerate(part.send_idx):
                send_off[i + 1] = send_off[i] + len(sidx)
            send_idx1 = np.ascontiguousarray(np.concatenate(part.send_idx) + 1 if part.send_idx else np.zeros(0), dtype=np.int64)
            recv = np.ascontiguousarray(part.recv_count, dtype=np.int64)
            L.check(lib.mft_set_halo(ctx, len(part.peers), peers, L.ptr(send_off), L.ptr(send_idx1), L.ptr(recv)))
            if eng.exchange == "p2p":
                # map the peers' state arrays (CUDA IPC) and tell every sender where its block starts in my halo tail
                hbuf = C.create_string_buffer(192)
                L.check(lib.mft_p2p_handles(ctx, hbuf))
                all_h = comm.allgather(bytes(hbuf.raw))
                my_rows, off = {}, part.n_local
                for q, cnt in zip(part.peers, part.recv_count):
                    my_rows[int(q)] = int(off)
                    off += cnt
                all_rows = comm.allgather(my_rows)
                dst_row = np.ascontiguousarray([all_rows[q].get(comm.rank, 0) for q in part.peers], dtype=np.int64)
                blob = C.create_string_buffer(b"".join(all_h), 192 * comm.nranks)
                L.check(lib.mft_p2p_connect(ctx, comm.nranks, comm.rank, blob, L.ptr(dst_row), part.n_global))
                comm.barrier()
            elif eng.exchange != "nccl":
                raise ValueError("RBFFDEngineCUDA.exchange must be 'p2p' or 'nccl'")

    def count_nonfinite(self):
        """number of NaN / Inf entries in the owned rows of the device-resident state (mft_count_nonfinite)"""
        cnt = C.c_int64()
        L.check(L.load().mft_count_nonfinite(self.ctx, C.byref(cnt)))
        return int(cnt.value)

    def refresh_boundary_values(self, t):
        lib = L.load()
        for g, (name, bc, tag) in enumerate(self._bc_groups):
            if bc.kind == L.BC_DIRICHLET and bc.time_dependent:
                vals = np.ascontiguousarray(bc.boundary_value_function(self.domain.pd.points[tag.idx], t, self.equations))
                L.check(lib.mft_update_boundary_values(self.ctx, g, L.ptr(vals)))

    def has_time_dependent_bcs(self):
        return any(bc.kind == L.BC_DIRICHLET and bc.time_dependent for _, bc, _ in self._bc_groups)

    def set_stage_boundary_values(self, t, dt):
        """Dirichlet tables for the stage times of the next device-resident step: the reference's BC closures receive the
        stage time (rbfsolver.jl:311-316) and a step evaluates rhs! at t + dt and t + dt/2 (mft_set_stage_boundary_values)."""
        lib = L.load()
        for g, (name, bc, tag) in enumerate(self._bc_groups):
            if bc.kind == L.BC_DIRICHLET and bc.time_dependent:
                for slot, ts in ((0, t + dt), (1, t + dt / 2)):
                    vals = np.ascontiguousarray(bc.boundary_value_function(self.domain.pd.points[tag.idx], ts, self.equations),
                                                dtype=np.float64)
                    L.check(lib.mft_set_stage_boundary_values(self.ctx, g, slot, L.ptr(vals)))

    def close(self):
        if getattr(self, "ctx", None):
            L.load().mft_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


@dataclass
class ODEProblem:
    u0: np.ndarray
    tspan: tuple
    p: SemidiscretizationHyperbolic


def compute_coefficients(t, semi):
    """rbfsolver.jl:174-185: u[i] = initial_condition(points[i], t, equations)"""
    u = np.ascontiguousarray(semi.initial_condition(semi.domain.pd.points, t, semi.equations), dtype=np.float64)
    assert u.shape == (semi.V, semi.n)
    return u


def semidiscretize(semi, tspan):
    return ODEProblem(compute_coefficients(tspan[0], semi), tuple(tspan), semi)


def rhs_(du, u, semi, t):
    """Trixi.rhs!(du_ode, u_ode, semi, t) -> Trixi.rhs!(du,u,t,domain,...) rbfsolver.jl:397-428.  u is IN/OUT."""
    semi.refresh_boundary_values(t)
    L.check(L.load().mft_rhs(semi.ctx, float(t), L.soa_ptrs(u), L.soa_ptrs(du), L.MEM_HOST))


def calc_fluxes_(du, u, semi):
    """calc_fluxes!(du,u,domain,...,engine::RBFFDEngineCUDA,...)  (reference CPU method: rbfsolver.jl:247-265)"""
    L.check(L.load().mft_calc_fluxes(semi.ctx, L.soa_ptrs(u), L.soa_ptrs(du)))


def calc_boundary_flux_(du, u, semi, t):
    semi.refresh_boundary_values(t)
    L.check(L.load().mft_boundary_pass(semi.ctx, float(t), L.soa_ptrs(u), L.soa_ptrs(du)))


# ---- callbacks + time integration (history.jl:15-129; SSPRK: SURVEY.md appendix B.5) -------------------------------
@dataclass
class HistoryCallback:
    approx_order: int = 4


class SSPRK33:
    """SSPRK33(stage_limiter!) of OrdinaryDiffEq: the optional stage limiter runs on the device after every stage update"""
    scheme = L.SSPRK33
    stages = 3

    def __init__(self, stage_limiter=None):
        self.stage_limiter = stage_limiter


def density(u, equations=None):
    """Trixi.density (third party): first conserved variable"""
    return u[0]


def pressure(u, equations):
    """Trixi.pressure for CompressibleEulerEquations2D (third party)"""
    return (equations.gamma - 1.0) * (u[3] - 0.5 * (u[1] * u[1] + u[2] * u[2]) / u[0])


class PositivityPreservingLimiterZhangShu:
    """PositivityPreservingLimiterZhangShu(; thresholds, variables) (src/callbacks_stage/positivity_zhang_shu.jl:15-48):
    applied to the scalar `variables` (density, pressure) in their given order with the associated thresholds."""

    _KINDS = {density: L.VAR_DENSITY, pressure: L.VAR_PRESSURE}

    def __init__(self, thresholds, variables):
        if len(thresholds) != len(variables):
            raise ValueError("thresholds and variables must have the same length")
        self.thresholds = tuple(float(x) for x in thresholds)
        self.variables = tuple(variables)
        try:
            self.kinds = tuple(self._KINDS[v] for v in self.variables)
        except KeyError:
            raise NotImplementedError("the device limiter knows Trixi's `density` and `pressure`") from None

    def _arrays(self):
        return (np.asarray(self.thresholds, dtype=np.float64), np.asarray(self.kinds, dtype=np.int32))

    @staticmethod
    def _neighbors(semi):
        """hand domain.pd.neighbors (list order = kNN order) to the library once per semidiscretization"""
        if semi._neighbors_set:
            return
        part = semi.partition
        if part is None:
            nbr = semi.domain.pd.neighbors
        else:   # one rank of a partitioned cloud: stencils of the owned rows (global ids) in local [owned ; halo] numbering
            lut = np.full(part.n_global, -1, dtype=np.int64)
            lut[part.local_gid] = np.arange(part.n_local + part.n_halo)
            nbr = lut[part.neighbors_owned]
            assert (nbr >= 0).all()
        nbr1 = np.ascontiguousarray(nbr + 1, dtype=np.int64)
        L.check(L.load().mft_set_neighbors(semi.ctx, L.ptr(nbr1)))
        semi._neighbors_set = True

    def __call__(self, u, semi, t=None):
        """limiter!(u_ode, integrator, semi, t) on a host array (in place)"""
        thr, kinds = self._arrays()
        self._neighbors(semi)
        L.check(L.load().mft_limiter_zhang_shu(semi.ctx, len(kinds), L.ptr(thr), L.ptr(kinds), L.soa_ptrs(u), L.MEM_HOST))
        return u

    def install(self, semi):
        thr, kinds = self._arrays()
        self._neighbors(semi)
        L.check(L.load().mft_set_stage_limiter(semi.ctx, len(kinds), L.ptr(thr), L.ptr(kinds)))


class SSPRK43:
    """adaptive; the error estimate comes from the library, the step-size controller runs on the host"""
    stages = 4


class PIController:
    """OrdinaryDiffEq's default PI controller for an order-3 method (third party, unpinned): beta1 = 7/(10k),
    beta2 = 2/(5k), gamma = 0.9, qmin = 0.2, qmax = 10, qsteady in [1, 1.2], qoldinit = 1e-4.  In the Julia deployment
    OrdinaryDiffEq keeps its own controller; this one drives the Python host loop."""

    def __init__(self, order=3, gamma=0.9, qmin=0.2, qmax=10.0, qsteady_min=1.0, qsteady_max=1.2, qoldinit=1e-4):
        self.beta1, self.beta2 = 7.0 / (10.0 * order), 2.0 / (5.0 * order)
        self.gamma, self.qmin, self.qmax = gamma, qmin, qmax
        self.qsteady_min, self.qsteady_max, self.qoldinit = qsteady_min, qsteady_max, qoldinit
        self.qold = qoldinit

    def propose(self, dt, eest):
        if eest == 0.0:
            q, q11 = 1.0 / self.qmax, 0.0
        else:
            q11 = eest ** self.beta1
            q = max(1.0 / self.qmax, min(1.0 / self.qmin, q11 / self.qold ** self.beta2 / self.gamma))
        if eest <= 1.0:
            if self.qsteady_min <= q <= self.qsteady_max:
                q = 1.0
            self.qold = max(eest, self.qoldinit)
            return True, dt / q
        return False, dt / min(1.0 / self.qmin, q11 / self.gamma)


@dataclass
class Solution:
    u: np.ndarray
    t: float
    nsteps: int
    nrhs: int


def _run_savers(savers, semi, t, it, finished, template):
    """SolutionSavingCallback affect! (save_solution_vtk.jl:136-173): download the resident state only when a file is due"""
    due = [c for c in savers if c.due(t, it, finished)]
    if not due:
        return
    u = np.empty_like(template)
    L.check(L.load().mft_download_state(semi.ctx, L.soa_ptrs(u)))
    for c in due:
        c(u, semi, t, it, finished)


def solve_adaptive(ode, alg, dt, abstol=1e-8, reltol=1e-8, callback=None, max_steps=100000, allreduce=None):
    """solve(ode, SSPRK43(); abstol, reltol, callback=...) (rbfsolver_test.jl:104-107): adaptive steps, FSAL, callbacks
    after every ACCEPTED step.  `allreduce(sumsq, count) -> (sumsq, count)` combines ranks in multi-GPU runs."""
    semi = ode.p
    lib = L.load()
    t0, t1 = ode.tspan
    cbs = [] if callback is None else (list(callback) if isinstance(callback, (list, tuple)) else [callback])
    hist = [c for c in cbs if isinstance(c, HistoryCallback)]
    savers = [c for c in cbs if hasattr(c, "due")]
    u0 = np.ascontiguousarray(ode.u0, dtype=np.float64)
    L.check(lib.mft_upload_state(semi.ctx, L.soa_ptrs(u0)))
    t, dt = float(t0), float(dt)
    tdep = semi.has_time_dependent_bcs()
    semi.refresh_boundary_values(t)
    for h in hist:
        L.check(lib.mft_history_push(semi.ctx, t, 0, h.approx_order))
    _run_savers(savers, semi, t, 0, False, u0)
    ctrl = PIController()
    accepted, log, nrhs = 0, [], 1
    ss, cnt = C.c_double(), C.c_int64()
    for _ in range(max_steps):
        if t >= t1 - 1e-14 * max(1.0, abs(t1)):
            break
        dt = min(dt, t1 - t)
        if tdep:
            semi.set_stage_boundary_values(t, dt)
        L.check(lib.mft_ssprk43_step(semi.ctx, t, dt, abstol, reltol, C.byref(ss), C.byref(cnt)))
        sumsq, count = (ss.value, cnt.value) if allreduce is None else allreduce(ss.value, cnt.value)
        eest = float(np.sqrt(sumsq / count))
        accept, dt_next = ctrl.propose(dt, eest)
        L.check(lib.mft_step_commit(semi.ctx, int(accept)))
        log.append((t, dt, eest, accept))
        nrhs += alg.stages
        if accept:
            t += dt
            accepted += 1
            for h in hist:
                L.check(lib.mft_history_push(semi.ctx, t, accepted, h.approx_order))
            _run_savers(savers, semi, t, accepted, t >= t1 - 1e-14 * max(1.0, abs(t1)), u0)
        dt = dt_next
    u = np.empty_like(u0)
    L.check(lib.mft_download_state(semi.ctx, L.soa_ptrs(u)))
    sol = Solution(u, t, accepted, nrhs)
    sol.log = log
    return sol


def solve(ode, alg, dt, callback=None, nsteps=None, **kw):
    """Fixed-step SSP integration with the state resident on the device (FSAL structure, callbacks after
    every step); mirrors solve(ode, SSPRK..; dt, adaptive=false, callback=CallbackSet(...)).  SSPRK43 -> adaptive."""
    if isinstance(alg, SSPRK43):
        return solve_adaptive(ode, alg, dt, callback=callback, **kw)
    semi = ode.p
    lib = L.load()
    t0, t1 = ode.tspan
    if nsteps is None:
        nsteps = int(round((t1 - t0) / dt))
    cbs = [] if callback is None else (list(callback) if isinstance(callback, (list, tuple)) else [callback])
    hist = [c for c in cbs if isinstance(c, HistoryCallback)]
    savers = [c for c in cbs if hasattr(c, "due")]
    lim = getattr(alg, "stage_limiter", None)
    if lim is not None:
        lim.install(semi)
        semi._stage_limiter_installed = True
    elif getattr(semi, "_stage_limiter_installed", False):
        L.check(lib.mft_set_stage_limiter(semi.ctx, 0, None, None))
        semi._stage_limiter_installed = False
    u0 = np.ascontiguousarray(ode.u0, dtype=np.float64)
    L.check(lib.mft_upload_state(semi.ctx, L.soa_ptrs(u0)))
    t = float(t0)
    tdep = semi.has_time_dependent_bcs()
    semi.refresh_boundary_values(t)      # Dirichlet data at the start time (no-op for time-independent tables)
    for h in hist:   # initialize! (history.jl:51-54)
        L.check(lib.mft_history_push(semi.ctx, t, 0, h.approx_order))
    _run_savers(savers, semi, t, 0, False, u0)   # initialize_save_cb! (save_solution_vtk.jl:123-134)
    nrhs = 1 if nsteps > 0 else 0
    for it in range(nsteps):
        if tdep:
            semi.set_stage_boundary_values(t, float(dt))
        L.check(lib.mft_ssprk_step(semi.ctx, alg.scheme, t, float(dt)))
        t = t + dt
        nrhs += alg.stages
        for h in hist:
            L.check(lib.mft_history_push(semi.ctx, t, it + 1, h.approx_order))
        _run_savers(savers, semi, t, it + 1, it + 1 == nsteps, u0)
    u = np.empty_like(u0)
    L.check(lib.mft_download_state(semi.ctx, L.soa_ptrs(u)))
    return Solution(u, t, nsteps, nrhs)
