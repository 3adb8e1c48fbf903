"""Union-tile layouts built on the GPU (MFT_OPT_LAYOUT_DEVICE, csrc/mft_layout_device.inl + csrc/mft_tile_build.cuh) against the host
builder: the layouts must be the same BYTES (FNV checksum over every array, forward and transposed operator), and rhs! / a few
SSPRK33 steps through them the same bits.  (The per-tile code is shared with the CPU tier: tests/test_abi_cpu.py compares it with the
host builder array by array; here it runs as the device kernels.)"""
import ctypes as C

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

NAMES = dict(left=1, right=2, bottom=3, top=4)


def _semi(m, cl, ic, source, **eng):
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(**eng))
    domain = m.PointCloudDomain(solver, cl, NAMES)
    eq = m.CompressibleEulerEquations2D(cases.GAMMA)
    if source == "residual":
        srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, c_rv=1.0, c_uw=1.0, polydeg=3))
    else:
        srcs = m.SourceTerms(uw=m.SourceUpwindViscosityTominec(solver, eq, domain, c_uw=1.0))
    bcs = {k: m.BoundaryConditionDirichlet(ic) for k in NAMES}
    return m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bcs, source_terms=srcs), domain


def _checksums(m, semi):
    L = m._lib
    lib = m.load()
    out = []
    for which in (0, 1):
        fnv, nbytes = C.c_ulonglong(0), C.c_longlong(0)
        L.check(lib.mft_debug_tiler_checksum(semi.ctx, which, C.byref(fnv), C.byref(nbytes)))
        out.append((fnv.value, nbytes.value))
    return out


def _steps(m, semi, u0, dt, nsteps):
    L = m._lib
    lib = m.load()
    ctx = semi.ctx
    L.check(lib.mft_upload_state(ctx, L.soa_ptrs(u0)))
    L.check(lib.mft_history_push(ctx, 0.0, 0, 3))
    t = 0.0
    for i in range(nsteps):
        L.check(lib.mft_ssprk_step(ctx, L.SSPRK33, t, dt))
        t += dt
        L.check(lib.mft_history_push(ctx, t, i + 1, 3))
    u = np.empty_like(u0)
    L.check(lib.mft_download_state(ctx, L.soa_ptrs(u)))
    return u


@pytest.mark.parametrize("tile", [31, 15, 7, 3])
def test_device_layout_equals_host_layout(tile):
    """every layout variant of the default one-row-per-thread tiles: plain slots (3), coloured (7), two copies (15), tuned (31)"""
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(160, 120, 10.0, 7.5, seed=3)   # 19200 points + boundary ring: 150 tiles, a partial last tile
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, cases.GAMMA, center=(5.0, 4.0))   # noqa: E731
    u0 = np.ascontiguousarray(ic(cl.points, 0.0))
    res = {}
    for dev in (True, False):
        semi, domain = _semi(m, cl, ic, "residual", tile=tile, layout_device=dev)
        res[dev] = (_checksums(m, semi), _steps(m, semi, u0, 0.1 * domain.pd.dx_min / 8.0, 3))
        semi.close()
    assert res[True][0] == res[False][0], f"device-built layout differs from the host-built one: {res[True][0]} vs {res[False][0]}"
    assert res[True][0][0][1] > 0 and res[True][0][1][1] > 0
    assert np.array_equal(res[True][1], res[False][1])


def test_device_layout_on_the_fixture_cloud_and_wider_stencils():
    """the reference's test cloud (its own boundary groups) and a wide stencil (nv = 42): long rows, large unions"""
    import mft_b200 as m

    ic = cases.ic_smooth_euler
    for cloud, names, nv in ((cases.FIXTURE, cases.BOUNDARY_NAMES, None), (m.cloud.jittered_lattice(48, 40, 6.0, 5.0, seed=5), NAMES, 42)):
        sums = {}
        for dev in (True, False):
            basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)), nv=nv)
            solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(layout_device=dev))
            domain = m.PointCloudDomain(solver, cloud, names)
            eq = m.CompressibleEulerEquations2D(cases.GAMMA)
            srcs = m.SourceTerms(uw=m.SourceUpwindViscosityTominec(solver, eq, domain, c_uw=1.0))
            bcs = {k: m.BoundaryConditionDirichlet(ic) for k in names}
            semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bcs, source_terms=srcs)
            u = np.ascontiguousarray(ic(domain.pd.points, 0.0))
            du = np.zeros_like(u)
            m.rhs_(du, u, semi, 0.0)
            sums[dev] = (_checksums(m, semi), du)
            semi.close()
        assert sums[True][0] == sums[False][0]
        assert np.array_equal(sums[True][1], sums[False][1])


def test_device_layout_build_time_is_reported(capfd):
    """MFT_TRACE prints the build time of either builder (262k points: 2 x 2056 tiles); same bytes again at that size"""
    import os

    import mft_b200 as m

    cl = m.cloud.jittered_lattice(512, 512, 10.0, 10.0, seed=4)
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, cases.GAMMA, center=(5.0, 5.0))   # noqa: E731
    sums = {}
    old = os.environ.get("MFT_TRACE")
    os.environ["MFT_TRACE"] = "1"
    try:
        for dev in (True, False):
            semi, _ = _semi(m, cl, ic, "residual", layout_device=dev, setup="device")
            sums[dev] = _checksums(m, semi)
            semi.close()
    finally:
        if old is None:
            del os.environ["MFT_TRACE"]
        else:
            os.environ["MFT_TRACE"] = old
    err = capfd.readouterr().err
    assert "union-tile layout on the device" in err and "host threads" in err, err
    with capfd.disabled():
        print("\n" + "\n".join(l for l in err.splitlines() if "union-tile layout" in l))
    assert sums[True] == sums[False]
