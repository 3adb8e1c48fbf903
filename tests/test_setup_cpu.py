"""CPU tests of the setup-time host code (kNN, batched RBF-FD weights, Medusa I/O, SFC order) against the oracle."""
import os

import numpy as np
import pytest

import cases
from cases import orc


def _mft():
    import mft_b200

    return mft_b200


@pytest.mark.parametrize("p,N,k", [(3, 3, None), (5, 3, None), (5, 3, 2), (5, 3, 4), (3, 2, None)])
def test_batched_weights_match_oracle(p, N, k):
    """product setup (batched LU) vs the oracle's per-point Bunch-Kaufman solve: same sparsity (bit-exact),
    weights to 1e-8 of the row scale (local systems have condition numbers 1e3-1e5)"""
    m = _mft()
    s = cases.fixture_setup(p=p, N=N)
    ours = m.setup_ops.compute_flux_operator(s["points"], s["nb"], p, N, k)
    ref = orc.compute_flux_operator(s["points"], s["nb"], p, N, k)
    for A, B in zip(ours, ref):
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
        scale = np.abs(B.data).max()
        tol = 1e-8 if (k or 1) <= 2 else 1e-5   # 4th derivatives of r^5 at (eps,eps) are ~1/eps
        assert np.abs(A.data - B.data).max() <= tol * scale


def test_knn_and_spacing_bit_exact():
    m = _mft()
    s = cases.fixture_setup()
    nb, dx_min, dx_avg = m.setup_ops.knn(s["points"], s["nv"])
    assert np.array_equal(nb, s["nb"]) and dx_min == s["dx_min"] and dx_avg == s["dx_avg"]
    assert m.setup_ops.num_neighbors(3) == 20 and m.setup_ops.num_neighbors(4) == 30


def test_medusa_roundtrip(tmp_path):
    m = _mft()
    cl = m.cloud.read_medusa_file(cases.FIXTURE)
    pts, _, bidx, bnrm = orc.read_medusa_file(cases.FIXTURE)
    assert np.array_equal(cl.points, pts)
    for g in range(5):
        assert np.array_equal(cl.boundary_idxs[g], bidx[g]) and np.array_equal(cl.boundary_normals[g], bnrm[g])
    case = os.path.join(tmp_path, "rt")
    m.cloud.write_medusa_file(case, cl)
    cl2 = m.cloud.read_medusa_file(case)
    assert np.array_equal(cl2.points, cl.points)
    for g in range(5):
        assert np.array_equal(np.sort(cl2.boundary_idxs[g]), np.sort(cl.boundary_idxs[g]))
    # the oracle's reader (restating read_medusa_file.jl) ingests what we wrote
    pts3, _, bidx3, _ = orc.read_medusa_file(case)
    assert np.array_equal(pts3, cl.points) and [len(b) for b in bidx3] == [40, 40, 60, 60, 26]


def test_synthetic_cloud_generator():
    m = _mft()
    cl = m.cloud.jittered_lattice(32, 16, 10.0, 5.0, seed=0)
    assert cl.points.shape == (32 * 16 + 2 * 16 + 2 * 32, 2)
    assert [len(b) for b in cl.boundary_idxs] == [16, 16, 32, 32]
    assert len(np.unique(cl.points, axis=0)) == len(cl.points)
    # same seed -> same cloud (PCG64), different seed -> different
    assert np.array_equal(cl.points, m.cloud.jittered_lattice(32, 16, 10.0, 5.0, seed=0).points)
    assert not np.array_equal(cl.points, m.cloud.jittered_lattice(32, 16, 10.0, 5.0, seed=1).points)
    u = m.cloud.isentropic_vortex(cl.points)
    assert u.shape == (4, len(cl.points)) and (u[0] > 0).all()


def test_sfc_order_is_a_locality_preserving_permutation():
    m = _mft()
    s = cases.fixture_setup()
    perm = m._lib.sfc_order(s["points"])
    assert np.array_equal(np.sort(perm), np.arange(len(perm)))
    # consecutive points along the curve are close: mean hop << domain size
    hop = np.linalg.norm(np.diff(s["points"][perm], axis=0), axis=1)
    assert hop.mean() < 3 * s["dx_avg"]
