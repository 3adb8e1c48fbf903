"""SourceIGR (reference src/sources/IGR.jl), CPU tier:
 * the oracle's C restatement against an independent numpy/scipy restatement of the same formulas and of
   IterativeSolvers.cg! (parity unpinned: no reference test uses the source, the solver is third party);
 * the product's kernel thread bodies + launch sequence + reduction tree, run on the host by tests/emu, against the oracle.
The GPU tier is tests/test_zz_i_igr_gpu.py."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
import emu
from cases import orc


def _setup():
    fx = cases.fixture_setup(p=3, N=3)
    ops = orc.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    return fx, ops


def _numpy_igr(ops, u, alpha, maxiter):
    Dx, Dy = ops
    rho, v1, v2 = u[0], u[1] / u[0], u[2] / u[0]
    fx1, fx2, fx3, fy2, fy3 = Dx @ rho, Dx @ v1, Dx @ v2, Dy @ v1, Dy @ v2
    b = alpha * ((fx1 + fy2) ** 2 + (fx2 ** 2 + 2 * fy2 * fx3 + fy3 ** 2))      # the reference's own indexing, IGR.jl:141-157
    ri = 1.0 / rho
    A = sp.diags(ri) - alpha * (Dx @ sp.diags(ri) @ Dx + Dy @ sp.diags(ri) @ Dy)
    x, r, p = np.zeros_like(b), b.copy(), np.zeros_like(b)
    res, prev, it = np.linalg.norm(r), 1.0, 0
    tol = np.sqrt(np.finfo(float).eps) * res
    while res > tol and it < maxiter:
        beta = res ** 2 / prev ** 2
        p = r + beta * p
        c = A @ p
        a = res ** 2 / (p @ c)
        x += a * p
        r -= a * c
        prev, res, it = res, np.linalg.norm(r), it + 1
    return b, A, x, it, res


@pytest.mark.parametrize("alpha_scale,maxiter", [(20.0, 20), (1.0, 20), (20.0, 3), (0.01, 20)])
def test_oracle_igr_matches_numpy_restatement(alpha_scale, maxiter):
    fx, ops = _setup()
    alpha = alpha_scale * fx["dx_avg"] ** 2
    src = orc.source_igr(alpha=alpha, maxiter=maxiter)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], [], [src])
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    du0 = 0.01 * np.cos(u)
    du = du0.copy()
    P.apply_source(0, u, du)
    b, A, x, it, res = _numpy_igr(ops, u, alpha, maxiter)
    sigma = src.arrays["sigma"]
    assert src.arrays["iters"] == it
    assert np.abs(sigma - x).max() <= 1e-9 * np.abs(x).max()
    assert abs(src.arrays["res"] - res) <= 1e-8 * res
    # du[2] -= Dx sigma, du[3] -= Dy sigma; density and energy untouched
    assert np.array_equal(du[0], du0[0]) and np.array_equal(du[3], du0[3])
    assert np.abs(du[1] - (du0[1] - ops[0] @ sigma)).max() <= 1e-12 * np.abs(du[1]).max()
    assert np.abs(du[2] - (du0[2] - ops[1] @ sigma)).max() <= 1e-12 * np.abs(du[2]).max()
    if alpha_scale == 0.01:
        # weak coupling: the operator is close to diag(1/rho) and CG converges below sqrt(eps) |b| within the cap
        assert it < maxiter and np.linalg.norm(A @ sigma - b) <= 1e-7 * np.linalg.norm(b)


def test_oracle_igr_inside_rhs():
    """calc_sources! order (rbfsolver.jl:388-395): flux divergence, then IGR on the same u"""
    fx, ops = _setup()
    alpha = 0.01 * fx["dx_avg"] ** 2
    src = orc.source_igr(alpha=alpha)
    ic = cases.ic_smooth_euler
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1],
                          cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, ic), [src])
    P0 = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1],
                           cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, ic), [])
    u = ic(fx["points"], 0.0)
    du, du0 = P.rhs(u.copy()), P0.rhs(u.copy())
    interior = np.ones(len(fx["points"]), bool)
    for g in fx["bidx"]:
        interior[g] = False
    sigma = src.arrays["sigma"]
    assert np.array_equal(du[0], du0[0]) and np.array_equal(du[3], du0[3])
    d = du[1] - du0[1] + ops[0] @ sigma
    assert np.abs(d[interior]).max() <= 1e-11 * max(1.0, np.abs(du[1]).max())
    assert np.abs(sigma).max() > 0


def _sorted_rows(nb, ops):
    """neighbour table + weights with each row in ascending column order (the order the device stores a row in)"""
    o = np.argsort(nb, axis=1, kind="stable")
    nbs = np.take_along_axis(nb, o, 1)
    n = nb.shape[0]
    rows = np.repeat(np.arange(n), nb.shape[1])
    wx = np.asarray(ops[0].tocsr()[rows, nbs.ravel()]).reshape(nbs.shape)
    wy = np.asarray(ops[1].tocsr()[rows, nbs.ravel()]).reshape(nbs.shape)
    return nbs, wx, wy


@pytest.mark.parametrize("alpha_scale,maxiter", [(20.0, 20), (0.01, 20), (5.0, 0), (5.0, 2)])
def test_emulated_device_kernels_match_oracle(alpha_scale, maxiter):
    """right-hand side b and the flux accumulation are bit-exact (same summation order); sigma to rounding (dot products:
    fixed device tree vs pairwise in the oracle vs BLAS in the reference)"""
    fx, ops = _setup()
    alpha = alpha_scale * fx["dx_avg"] ** 2
    nbs, wx, wy = _sorted_rows(fx["nb"], ops)
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    du0 = 0.01 * np.sin(3 * u)
    du = du0.copy()
    b = np.empty(len(fx["points"]))
    sigma, (it, res, res0) = emu.igr_apply(nbs, wx, wy, alpha, maxiter, u, du, b_out=b)
    src = orc.source_igr(alpha=alpha, maxiter=maxiter)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], [], [src])
    du_ref = du0.copy()
    P.apply_source(0, u, du_ref)
    sref = src.arrays["sigma"]
    n = len(b)
    assert np.array_equal(b, src.arrays["igr_work"][n:2 * n])      # update_igr_rhs!: same summation order -> bit-identical
    assert it == src.arrays["iters"]
    if maxiter == 0:
        assert np.array_equal(sigma, np.zeros_like(sigma)) and np.array_equal(du, du0)
        return
    assert np.abs(sigma - sref).max() <= 1e-9 * np.abs(sref).max()
    assert abs(res - src.arrays["res"]) <= 1e-8 * res0
    assert np.array_equal(du[0], du0[0]) and np.array_equal(du[3], du0[3])
    assert cases.relerr(du, du_ref) <= 1e-9
    # bit-exact pieces: feed the oracle's sigma through the same accumulation -> identical du rows
    b_ref = _numpy_igr(ops, u, alpha, 0)[0]
    assert abs(res0 - np.linalg.norm(b_ref)) <= 1e-12 * res0


def test_emulated_kernels_one_iteration_is_bit_exact_up_to_the_dot_products():
    """with maxiter = 1 the only order-dependent quantities are |r|^2 and p.c; everything else must agree to the last
    bit once those two scalars agree, so sigma = a * b elementwise with one common factor a"""
    fx, ops = _setup()
    alpha = 2.0 * fx["dx_avg"] ** 2
    nbs, wx, wy = _sorted_rows(fx["nb"], ops)
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    du = np.zeros_like(u)
    sigma, (it, res, res0) = emu.igr_apply(nbs, wx, wy, alpha, 1, u, du)
    src = orc.source_igr(alpha=alpha, maxiter=1)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1], [], [src])
    P.apply_source(0, u, np.zeros_like(u))
    sref = src.arrays["sigma"]
    nz = sref != 0
    ratio = sigma[nz] / sref[nz]
    assert it == 1 and np.ptp(ratio) <= 1e-15 and abs(ratio[0] - 1.0) <= 1e-13
