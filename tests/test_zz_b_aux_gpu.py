"""Auxiliary entry points: failure detection on the resident state.  CPU: the kernel's thread body under
host emulation.  GPU: mft_count_nonfinite through the C ABI (first hardware run: round-end pass)."""
import numpy as np
import pytest

import cases
import emu


def _poisoned(u):
    u = u.copy()
    u[0, 3] = np.nan
    u[2, 77] = np.inf
    u[3, 77] = -np.inf
    u[1, 2000] = np.nan
    return u


def test_emulated_nonfinite_count():
    fx = cases.fixture_setup()
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    assert emu.count_nonfinite(u) == 0
    assert emu.count_nonfinite(_poisoned(u)) == 4
    assert emu.count_nonfinite(np.full((1, 5), -np.inf)) == 5


@pytest.mark.gpu
def test_device_nonfinite_count():
    import mft_b200 as m

    fx = cases.fixture_setup(p=3, N=3)
    ops = m.setup_ops.compute_flux_operator(fx["points"], fx["nb"], 3, 3)
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=fx["nv"])
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA())
    domain = m.PointCloudDomain(solver, cases.FIXTURE, cases.BOUNDARY_NAMES)
    semi = m.SemidiscretizationHyperbolic(domain, m.CompressibleEulerEquations2D(cases.GAMMA), cases.ic_smooth_euler, solver,
                                          boundary_conditions=dict(inlet=m.BoundaryConditionDoNothing()), operators=ops)
    L = m._lib
    u = cases.ic_smooth_euler(fx["points"], 0.0)
    L.check(L.load().mft_upload_state(semi.ctx, L.soa_ptrs(u)))
    assert semi.count_nonfinite() == 0
    bad = _poisoned(u)
    L.check(L.load().mft_upload_state(semi.ctx, L.soa_ptrs(bad)))
    assert semi.count_nonfinite() == 4
    semi.close()
