"""Committed golden vectors of the "next" rows (tests/golden/fixture_golden_f.npz + the kNN table / weights stored in
fixture_golden.npz).  CPU: the oracle reproduces them, and so do the product's kernel bodies under host emulation, without
the oracle in the loop.  GPU (tests/test_zz_h_golden_f_gpu.py): the CUDA path reproduces them."""
import os
import sys

import numpy as np

import cases
import emu

sys.path.insert(0, cases.GOLDEN)
import make_golden_f as mg  # noqa: E402

G = np.load(os.path.join(cases.GOLDEN, "fixture_golden.npz"))
GF = np.load(os.path.join(cases.GOLDEN, "fixture_golden_f.npz"))


def test_oracle_reproduces_golden_f():
    from cases import orc

    fx = cases.fixture_setup(p=3, N=3)
    nb, ops = mg.golden_ops()
    u = orc.limiter_zhang_shu(mg.limiter_state(fx["points"]), nb, mg.LIMITER["thresholds"], mg.LIMITER["variables"], cases.GAMMA)
    assert np.array_equal(u, GF["limiter_u"])
    src = orc.source_igr(alpha=float(GF["igr_alpha"]), maxiter=mg.IGR_MAXITER)
    P = orc.OracleProblem(fx["points"], 4, orc.EQ_EULER2D, [cases.GAMMA], ops[0], ops[1],
                          cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, cases.ic_smooth_euler), [src])
    du = P.rhs(cases.ic_smooth_euler(fx["points"], 0.0), 0.0)
    assert src.arrays["iters"] == int(GF["igr_iters"])
    assert cases.relerr(du, GF["igr_rhs_du"]) < 1e-12
    assert np.abs(src.arrays["sigma"] - GF["igr_sigma"]).max() < 1e-12 * np.abs(GF["igr_sigma"]).max()


def test_emulated_setup_kernels_reproduce_golden_tables():
    pts = cases.orc.read_medusa_file(cases.FIXTURE)[0]
    nb, d = emu.setup_knn(pts, 20)
    assert np.array_equal(nb, G["neighbors"]) and d[:, 1].min() == float(G["dx_min"]) and d[:, 1].mean() == float(G["dx_avg"])
    wx, wy = emu.setup_rbf_weights(pts, nb, 3, 3, 1)
    assert np.abs(wx - G["wx"]).max() <= 1e-8 * np.abs(G["wx"]).max() and np.abs(wy - G["wy"]).max() <= 1e-8 * np.abs(G["wy"]).max()


def test_emulated_limiter_and_igr_kernels_reproduce_golden():
    pts = cases.orc.read_medusa_file(cases.FIXTURE)[0]
    nb = G["neighbors"].astype(np.int64)
    u = emu.limiter_zhang_shu(mg.limiter_state(pts), nb, mg.LIMITER["thresholds"], mg.LIMITER["variables"], cases.GAMMA)
    assert np.array_equal(u, GF["limiter_u"])
    # IGR source alone on the golden operators: sigma (the rhs! golden also contains the flux divergence and the BCs,
    # which the emulation does not cover; sigma depends on u only)
    o = np.argsort(nb, axis=1, kind="stable")
    nbs = np.take_along_axis(nb, o, 1)
    wx, wy = np.take_along_axis(G["wx"], o, 1), np.take_along_axis(G["wy"], o, 1)
    fx = cases.fixture_setup(p=3, N=3)
    u0 = cases.ic_smooth_euler(pts, 0.0)
    # BC pass 1 of rhs! writes u at boundary points before the sources see it: take the golden's boundary-imposed state
    P = cases.orc.OracleProblem(pts, 4, cases.orc.EQ_EULER2D, [cases.GAMMA], *mg.golden_ops()[1],
                                cases.oracle_bcs(fx, cases.DIVERGENCE_TEST_BCS, cases.ic_smooth_euler), [])
    P.rhs(u0, 0.0)                                      # u0 now carries the strong BCs (rhs! mutates u)
    sigma, (it, _, _) = emu.igr_apply(nbs, wx, wy, float(GF["igr_alpha"]), mg.IGR_MAXITER, u0, np.zeros_like(u0))
    assert it == int(GF["igr_iters"])
    assert np.abs(sigma - GF["igr_sigma"]).max() <= 1e-9 * np.abs(GF["igr_sigma"]).max()


def test_oracle_reproduces_advection_config_golden():
    """BASELINE configs[0] (advection + both hyperviscosity sources on the fixture cloud, 100 SSPRK33 steps) from the stored
    r^5 weight tables"""
    fx5 = cases.fixture_setup(p=5, N=3)
    nb = G["neighbors"].astype(np.int64)
    P, gam = mg.advection_problem(fx5, nb, GF)
    assert gam == (float(GF["adv_gamma_flyer"]), float(GF["adv_gamma_tominec"]))
    u = cases.ic_bump_advection(fx5["points"], 0.0)
    du = P.rhs(u, 0.0)
    assert np.array_equal(u, GF["adv_rhs_u"]) and cases.relerr(du, GF["adv_rhs_du"]) < 1e-13
    u_end, _ = P.solve_ssprk33(cases.ic_bump_advection(fx5["points"], 0.0), 0.0, float(GF["adv_dt"]), 100)
    assert cases.relerr(u_end, GF["adv_steps100_u"]) < 1e-12


def test_emulated_weight_kernel_reproduces_the_stored_r5_tables():
    pts = cases.orc.read_medusa_file(cases.FIXTURE)[0]
    nb = G["neighbors"].astype(np.int64)
    wx, wy = emu.setup_rbf_weights(pts, nb, 5, 3, 1)
    assert np.abs(wx - GF["wx5"]).max() <= 1e-8 * np.abs(GF["wx5"]).max() and np.abs(wy - GF["wy5"]).max() <= 1e-8 * np.abs(GF["wy5"]).max()
    lx, ly = emu.setup_rbf_weights(pts, nb, 5, 3, 2)
    assert np.abs(lx + ly - GF["lap"]).max() <= 1e-8 * np.abs(GF["lap"]).max()
    hx, hy = emu.setup_rbf_weights(pts, nb, 5, 3, 4)
    assert np.abs(hx + hy - GF["h4"]).max() <= 1e-5 * np.abs(GF["h4"]).max()     # 4th derivatives of r^5 at (eps,eps): SURVEY appendix A.8
