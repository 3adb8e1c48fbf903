"""Second, independent restatement of the reference's `rhs!` path in plain numpy/scipy (TEST INFRASTRUCTURE).

Why it exists: Julia cannot run in the build image, so `oracle/mft_oracle.c` *is* the reference for everything the
reference's own tests do not pin (both BC passes, `update_residual_visc!`, `update_visc!`, whole-`rhs!`, the history
callback inside a time loop).  This module was written from the Julia sources alone, following
their execution structure literally (component vectors of a StructArray = rows of a (V,N) array, SparseMatrixCSC =
scipy CSC, one mat-vec per field), without looking at the C restatement, so that `tests/test_oracle_vs_numpy.py` can
demand that two independent readings of the reference agree to rounding.  It is not bit-exact by construction (scipy's
CSC mat-vec adds `nzval*x` terms, SparseArrays adds `nzval*(x*alpha)`; no FMA control in numpy) and it is never imported
by the product or by bench.py.

Every function cites the reference lines it follows (paths relative to /root/reference/src).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

EPS = float(np.finfo(np.float64).eps)  # Base.eps()


# ---- Trixi.jl (third party, 0.6 - 0.7.5): CompressibleEulerEquations2D / LinearScalarAdvectionEquation2D -------------
class Euler2D:
    nvars = 4

    def __init__(self, gamma):
        self.gamma = float(gamma)

    def flux(self, u, orientation):
        """Trixi `flux(u, orientation, equations)`; call site solvers/pointcloudsolver/rbfsolver.jl:259"""
        rho, rho_v1, rho_v2, rho_e = u
        v1 = rho_v1 / rho
        v2 = rho_v2 / rho
        p = (self.gamma - 1.0) * (rho_e - 0.5 * (rho_v1 * v1 + rho_v2 * v2))
        if orientation == 1:
            return np.stack([rho_v1, rho_v1 * v1 + p, rho_v1 * v2, (rho_e + p) * v1])
        return np.stack([rho_v2, rho_v2 * v1, rho_v2 * v2 + p, (rho_e + p) * v2])

    def cons2prim(self, u):
        """Trixi `cons2prim`; call site sources/hyperviscosity.jl:254"""
        rho, rho_v1, rho_v2, rho_e = u
        v1 = rho_v1 / rho
        v2 = rho_v2 / rho
        p = (self.gamma - 1.0) * (rho_e - 0.5 * (rho_v1 * v1 + rho_v2 * v2))
        return rho, v1, v2, p


class Advection2D:
    nvars = 1

    def __init__(self, a):
        self.a = (float(a[0]), float(a[1]))

    def flux(self, u, orientation):
        return self.a[orientation - 1] * u


# ---- boundary conditions: equations/PointCloudBCs.jl ------------------------------------------------------------------
class Dirichlet:
    """BoundaryConditionDirichlet functor, PointCloudBCs.jl:49-63: returns (FluxZero() = zeros, u_boundary)"""

    def __init__(self, value_fn):
        self.value_fn = value_fn

    def __call__(self, du_inner, u_inner, normal, x, t):
        return np.zeros_like(du_inner), np.asarray(self.value_fn(x[None, :], t))[:, 0]


class SlipWall:
    """boundary_condition_slip_wall, PointCloudBCs.jl:87-106 with apply_slip_velocity :15-21"""

    def __call__(self, du_inner, u_inner, normal_direction, x, t):
        norm_ = np.sqrt(normal_direction[0] ** 2 + normal_direction[1] ** 2)
        normal = normal_direction / norm_
        v = u_inner[1:3]
        v_slip = v - (v[0] * normal[0] + v[1] * normal[1]) * normal
        u_local = np.array([u_inner[0], v_slip[0], v_slip[1], u_inner[3]])
        return np.array([du_inner[0], 0.0, 0.0, du_inner[3]]), u_local


class DoNothing:
    """BoundaryConditionDoNothing, PointCloudBCs.jl:108-115"""

    def __call__(self, du_inner, u_inner, normal, x, t):
        return du_inner, u_inner


# ---- sources: sources/hyperviscosity.jl ---------------------------------------------------------------------------------
class Hyperviscosity:
    """SourceHyperviscosityFlyer / SourceHyperviscosityTominec functors, hyperviscosity.jl:52-64 / :121-134:
    apply_to_each_field(mul_by_accum!(hv_differentiation_matrix, -gamma), du, u)"""

    def __init__(self, H, gamma):
        self.H, self.gamma = sp.csc_matrix(H), float(gamma)

    def __call__(self, du, u, t, prob):
        for f in range(u.shape[0]):
            du[f] += self.H @ (u[f] * -self.gamma)


class TominecViscosity:
    """SourceUpwindViscosityTominec (:351-380) and SourceResidualViscosityTominec (:382-409) with their cache (:202-244)"""

    def __init__(self, n, nvars, dx_avg, residual, c_rv=1.0, c_uw=1.0, polydeg=4):
        self.use_residual = residual
        self.c_rv, self.c_uw, self.dx_avg = c_rv, c_uw, dx_avg
        self.eps_uw, self.eps_rv, self.eps = np.zeros(n), np.zeros(n), np.zeros(n)
        self.eps_c = np.zeros(n, dtype=np.int64)
        self.residual = np.zeros((nvars, n))
        self.approx_du = np.zeros((nvars, n))
        self.time_history = np.zeros(polydeg + 1)
        self.time_weights = np.zeros(polydeg + 1)
        self.sol_history = np.zeros((polydeg + 1, nvars, n))   # [slot] = one (V,N) snapshot; slot 0 = most recent
        self.success_iter = 0

    def update_upwind_visc(self, u, eq):
        """update_upwind_visc!, :246-285"""
        gamma = eq.gamma
        rho, v1, v2, p = (a.copy() for a in eq.cons2prim(u))
        speed = np.sqrt(v1 ** 2 + v2 ** 2)
        bad = (p < 0.0) | (rho < 0.0)
        with np.errstate(invalid="ignore", divide="ignore"):
            sound_speed = np.where(bad, 0.0, np.sqrt(gamma * p / rho))
        h_loc = self.dx_avg
        self.eps_uw[:] = self.c_uw * 0.5 * h_loc * (speed + sound_speed)

    def update_residual_visc(self, du, u):
        """update_residual_visc!, :289-329; ode_mean auxiliary/mpi.jl:40-52; ode_maximum(::StructArray) :71-81"""
        self.residual[:] = np.abs(self.approx_du - du)
        V, n = u.shape
        mean_u = u.sum(axis=1) / (V * n)                       # recursive_length(u) = V*N
        local_u = np.abs(u - mean_u[:, None])
        # maximum(::StructArray{SVector}) compares SVectors with isless = lexicographic order
        best = 0
        for i in range(1, n):
            if tuple(local_u[:, i]) > tuple(local_u[:, best]):
                best = i
        n_inf_norms = np.array([EPS if x == 0.0 else x for x in local_u[:, best]])
        max_res = (self.residual / n_inf_norms[:, None]).max(axis=0)
        h_loc = self.dx_avg
        self.eps_rv[:] = 0.5 * self.c_rv * h_loc ** 2 * max_res
        self.n_inf_norms = n_inf_norms

    def update_visc(self):
        """update_visc!, :331-349"""
        for i in range(len(self.eps)):
            rv, uw = self.eps_rv[i], self.eps_uw[i]
            if np.isnan(rv) or np.isinf(rv) or self.success_iter == 0:
                if np.isnan(uw) or np.isinf(uw):
                    self.eps[i], self.eps_c[i] = EPS, 2
                else:
                    self.eps[i], self.eps_c[i] = uw, 1
            else:
                self.eps[i] = np.minimum(rv, uw)      # Julia's min propagates NaN
                self.eps_c[i] = 0 if rv < uw else 1

    def __call__(self, du, u, t, prob):
        self.update_upwind_visc(u, prob.eq)
        if self.use_residual:
            self.update_residual_visc(du, u)
            self.update_visc()
        else:
            self.eps[:] = self.eps_uw
            self.eps_c[:] = 1
        for D in prob.D:
            for f in range(u.shape[0]):
                local = D @ u[f]                  # mul_by!(D)
                local = self.eps * local          # mul_by!(eps)
                du[f] += D.T @ (local * -1.0)     # mul_by_accum!(D', -1)

    # HistoryCallback, callbacks_step/history.jl:91-152
    def modify_cache(self, u, t, success_iter, approx_order):
        self.success_iter = success_iter
        # shift_soln_history!, :105-111
        self.time_history[1:] = self.time_history[:-1].copy()
        self.time_history[0] = t
        self.sol_history[1:] = self.sol_history[:-1].copy()
        self.sol_history[0] = u
        # update_approx_du!, :113-129
        self.approx_du[:] = 0.0
        num_time_points = min(success_iter + 1, approx_order + 1)
        if success_iter > 0:
            self.time_weights[:num_time_points] = time_deriv_weights(self.time_history[:num_time_points])
            for i in range(num_time_points):
                self.approx_du += self.time_weights[i] * self.sol_history[i]


def time_deriv_weights(t):
    """time_deriv_weights!, history.jl:131-152"""
    t = np.asarray(t, dtype=np.float64)
    scale = 1.0 / np.abs(t).max()
    t_ = t * scale
    t_eval = t_[0]
    m = len(t_)
    A = np.zeros((m, m))
    b_t = np.zeros(m)
    for k in range(1, m + 1):
        A[:, k - 1] = t_ ** (k - 1)
        b_t[k - 1] = 0.0 if k == 1 else (k - 1) * t_eval ** (k - 2)
    return scale * np.linalg.solve(A.T, b_t)


# ---- rhs!: solvers/pointcloudsolver/rbfsolver.jl --------------------------------------------------------------------------
class RefProblem:
    def __init__(self, points, eq, Dx, Dy, bcs=(), sources=()):
        """bcs: ordered list of (functor, idx 0-based, normals (nb,2)) = the boundary_conditions NamedTuple;
        sources: ordered list of source functors = values(source_terms)"""
        self.points = np.asarray(points, dtype=np.float64)
        self.eq = eq
        self.D = [sp.csc_matrix(Dx), sp.csc_matrix(Dy)]
        self.bcs = list(bcs)
        self.sources = list(sources)

    def calc_boundary_flux(self, du, u, t):
        """calc_boundary_flux! :277-286 / calc_single_boundary_flux! :288-318"""
        for functor, idx, normals in self.bcs:
            for i in range(len(idx)):
                b = idx[i]
                du_b, u_b = functor(du[:, b].copy(), u[:, b].copy(), normals[i], self.points[b], t)
                du[:, b] = du_b
                u[:, b] = u_b

    def calc_fluxes(self, du, u):
        """calc_fluxes! :247-265: flux_values[e] = flux(u[e], i); du_f += -D_i flux_f"""
        for i in (1, 2):
            flux_values = self.eq.flux(u, i)
            for f in range(u.shape[0]):
                du[f] += self.D[i - 1] @ (flux_values[f] * -1.0)

    def rhs(self, u, t):
        """Trixi.rhs! :397-428 (u is mutated by the BC passes)"""
        du = np.zeros_like(u)                    # reset_du!
        self.calc_boundary_flux(du, u, t)
        self.calc_fluxes(du, u)
        for source in self.sources:              # calc_sources! :380-395
            source(du, u, t, self)
        self.calc_boundary_flux(du, u, t)
        return du

    def history_callback(self, u, t, success_iter, approx_order):
        """HistoryCallback affect! -> update_history! -> modify_cache! (history.jl:62-103)"""
        for s in self.sources:
            if isinstance(s, TominecViscosity) and s.use_residual:
                s.modify_cache(u, t, success_iter, approx_order)

    def solve_ssprk33(self, u0, t0, dt, nsteps, approx_order=None):
        """OrdinaryDiffEq SSPRK33 (third party), in-place FSAL form: k = f(u_n) carried over;
        u1 = uprev + dt k;  u2 = (3 uprev + u1 + dt f(u1, t+dt)) / 4;  u = (uprev + 2 u2 + 2 dt f(u2, t+dt/2)) / 3.
        Callbacks (HistoryCallback) run at initialisation and after every step."""
        u = np.array(u0, dtype=np.float64)
        t, success_iter = float(t0), 0
        if approx_order is not None:
            self.history_callback(u, t, success_iter, approx_order)
        k = self.rhs(u, t)
        for _ in range(nsteps):
            uprev = u.copy()
            u = uprev + dt * k
            k = self.rhs(u, t + dt)
            u = (3.0 * uprev + u + dt * k) / 4.0
            k = self.rhs(u, t + dt / 2)
            u = (uprev + 2.0 * u + 2.0 * dt * k) / 3.0
            k = self.rhs(u, t + dt)
            t += dt
            success_iter += 1
            if approx_order is not None:
                self.history_callback(u, t, success_iter, approx_order)
        return u, t
