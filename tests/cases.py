"""Shared problem builders for the tests: the same inputs (cloud, operators, BC tables, sources, state) are handed
to the CPU oracle (oracle/, the checker) and to the product (libmft_b200.so through the host mirror)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import mft_oracle as orc  # noqa: E402  (test infrastructure)

FIXTURE = os.path.join(ROOT, "tests", "golden", "cyl_0_05", "cyl_0_05")
GOLDEN = os.path.join(ROOT, "tests", "golden")
GAMMA = 1.4
BOUNDARY_NAMES = dict(inlet=1, outlet=2, bottom=3, top=4, cyl=5)


# --- initial conditions of the reference tests ------------------------------------------------------------
def ic_gradient(x, t, equations=None):
    """test/divergence_test.jl:28-38"""
    s = 0.1 * x[:, 0] + 0.1 * x[:, 1]
    return np.stack([1.4 + s, 4.1 + s, 0.0 + s, 8.8 + s])


def ic_oscillatory(x, t, equations=None):
    """test/hyperviscosity_test.jl:17-37"""
    s = 0.1 * x[:, 0] + 0.1 * x[:, 1] + 1.0 * np.sin(2 * np.pi * 100.0 * x[:, 0]) * np.sin(2 * np.pi * 100.0 * x[:, 1])
    return np.stack([1.4 + s, 4.1 + s, 0.0 + s, 8.8 + s])


def ic_smooth_euler(x, t, equations=None):
    """smooth, strictly positive density/pressure state used for rhs!/time-integration parity"""
    X, Y = x[:, 0], x[:, 1]
    rho = 1.0 + 0.2 * np.sin(1.3 * X) * np.cos(0.7 * Y)
    v1 = 0.5 + 0.1 * np.cos(0.9 * X + 0.3 * Y)
    v2 = -0.2 + 0.1 * np.sin(0.5 * X - 1.1 * Y)
    p = 1.0 + 0.1 * np.cos(0.4 * X) * np.sin(0.8 * Y)
    return np.stack([rho, rho * v1, rho * v2, p / (GAMMA - 1.0) + 0.5 * rho * (v1 * v1 + v2 * v2)])


def ic_bump_advection(x, t, equations=None):
    return np.exp(-8.0 * ((x[:, 0] - 1.0) ** 2 + (x[:, 1] - 1.0) ** 2))[None, :].copy()


# --- oracle-side builders -----------------------------------------------------------------------------------
def fixture_setup(p=3, N=3, nv=None):
    pts, interior, bidx, bnrm = orc.read_medusa_file(FIXTURE)
    nv = nv or orc.num_neighbors(N)
    nb, dx_min, dx_avg = orc.point_data(pts, nv)
    return dict(points=pts, bidx=bidx, bnrm=bnrm, nb=nb, dx_min=dx_min, dx_avg=dx_avg, p=p, N=N, nv=nv)


def oracle_bcs(setup, spec, ic):
    """spec: ordered dict name -> 'dirichlet' | 'slip' | 'nothing' (names of BOUNDARY_NAMES)"""
    out = []
    for name, kind in spec.items():
        g = BOUNDARY_NAMES[name] - 1
        idx, nrm = setup["bidx"][g], setup["bnrm"][g]
        if kind == "dirichlet":
            out.append(orc.OracleBC(orc.BC_DIRICHLET, idx, nrm, value_fn=lambda x, t, _ic=ic: _ic(x, t)))
        elif kind == "slip":
            out.append(orc.OracleBC(orc.BC_SLIP_WALL, idx, nrm))
        else:
            out.append(orc.OracleBC(orc.BC_DO_NOTHING, idx, nrm))
    return out


DIVERGENCE_TEST_BCS = dict(inlet="dirichlet", outlet="nothing", top="slip", bottom="slip", cyl="slip")


def relerr(a, b):
    """normwise per-variable relative error  max_v ||a_v - b_v||_inf / ||b_v||_inf"""
    a = np.atleast_2d(a)
    b = np.atleast_2d(b)
    worst = 0.0
    for v in range(a.shape[0]):
        den = np.abs(b[v]).max()
        num = np.abs(a[v] - b[v]).max()
        worst = max(worst, num / den if den > 0 else num)
    return worst
