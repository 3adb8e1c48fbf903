"""N>1 path.  CPU: world_size-2 gloo run of the production partition planner + the multi-rank rhs! algorithm
(tests/dist_worker.py --mode cpu).  GPU (needs >= 2 devices): the real library with NCCL halo exchange against the
serial oracle."""
import json
import os
import socket
import subprocess
import sys

import pytest

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _launch(nproc, mode, source, tmp_path, timeout=600, exchange="p2p", setup="host", fused=1, same_device=0):
    out = os.path.join(tmp_path, f"res_{mode}_{source}_{exchange}_{setup}_{fused}_{same_device}.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(cases.ROOT, "tests", "dist_worker.py"), "--mode", mode, "--source", source, "--out", out,
           "--exchange", exchange, "--setup", setup, "--fused", str(fused), "--same-device", str(same_device)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "MULTI_RANK_OK" in res.stdout
    return json.load(open(out))


def test_partition_and_algorithm_world2_gloo(tmp_path):
    res = _launch(2, "cpu", "upwind", str(tmp_path))
    assert len(res) == 2 and sum(r["n_local"] for r in res) == 72 * 60 + 2 * (72 + 60)
    assert all(r["n_halo"] > 0 and r["err"] < 1e-12 for r in res)


def test_flyer_hyperviscosity_on_a_partition_world2_gloo(tmp_path):
    """Flyer hyperviscosity needs only the u halo; the partition-aware source constructor gives every
    rank the global rows of H bit for bit, and the partitioned rhs! equals the serial oracle's"""
    res = _launch(2, "cpu", "flyer", str(tmp_path))
    assert len(res) == 2 and all(r["n_halo"] > 0 and r["err"] < 1e-12 for r in res)


def test_tominec_hyperviscosity_on_a_partition_world2_gloo(tmp_path):
    """SourceHyperviscosityTominec (H = L'L, two stencil rings) on a partitioned cloud: ParallelPointCloudDomain(wide_halo=True)
    keeps the stencil columns of the foreign rows R_r as column-only halo points; the owned rows of the local L'L equal the
    global ones and the partitioned rhs! equals the serial oracle's.  The narrow halo is refused."""
    res = _launch(2, "cpu", "tominec", str(tmp_path))
    assert len(res) == 2 and all(r["n_halo"] > 0 and r["err"] < 1e-12 for r in res)


def test_limiter_on_a_partition_world2_gloo(tmp_path):
    """Zhang-Shu limiter on a partitioned cloud: one u halo refresh per (threshold, variable) pass, owned rows limited from
    their global stencils -- equals the serial limiter on the global cloud"""
    res = _launch(2, "cpu", "limiter", str(tmp_path))
    assert len(res) == 2 and all(r["err"] < 1e-14 for r in res)


def test_partition_world3_gloo(tmp_path):
    res = _launch(3, "cpu", "upwind", str(tmp_path))
    assert len(res) == 3 and all(r["err"] < 1e-12 for r in res)


def _ngpu():
    try:
        import mft_b200

        return mft_b200._lib.load().mft_device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("nproc,source,fused", [(2, "residual", 1), (2, "upwind", 1), (3, "residual", 1), (2, "residual", 0)])
def test_ranks_sharing_one_gpu_match_serial_oracle(tmp_path, nproc, source, fused):
    """The multi-rank path on a SINGLE-GPU box: every rank is its own process with its own ctx on cuda:0, the ranks map each
    other's u / g / flag windows with CUDA IPC exactly as on several GPUs, and the device time-slices their kernels (a kernel
    that waits for a peer's flag spins until the peer's process gets the GPU).  Same checks as on several GPUs: one rhs! against
    the serial oracle to 1e-12, 10 SSPRK33 steps with the history callback to 1e-9, halo copies equal the owners' values, no
    norm misses.  fused = 1: fused stage kernel, band-tile waits and g puts inside pass A / pass B; 0: the separate kernels."""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    res = _launch(nproc, "gpu", source, str(tmp_path), timeout=420, exchange="p2p", fused=fused, same_device=1)
    assert len(res) == nproc and all(r["err"] < 1e-12 and r["err_steps"] < 1e-9 for r in res)


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
@pytest.mark.parametrize("source", ["upwind", "residual", "flyer", "tominec"])
def test_two_gpus_match_serial_oracle(tmp_path, source, exchange):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    res = _launch(2, "gpu", source, str(tmp_path), timeout=240, exchange=exchange)
    assert all(r["err"] < 1e-12 and r["err_steps"] < 1e-9 for r in res)


@pytest.mark.gpu
@pytest.mark.parametrize("source", ["upwind", "residual"])
def test_two_gpus_separate_kernels_match_serial_oracle(tmp_path, source):
    """MFT_OPT_FUSED_STEP = 0: the round-1 sequence (put / wait kernels, two norm round trips) stays covered"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    res = _launch(2, "gpu", source, str(tmp_path), timeout=240, exchange="p2p", fused=0)
    assert all(r["err"] < 1e-12 and r["err_steps"] < 1e-9 for r in res)


@pytest.mark.gpu
@pytest.mark.parametrize("nproc", [4, 8])
def test_many_gpus_p2p_match_serial_oracle(tmp_path, nproc):
    if _ngpu() < nproc:
        pytest.skip(f"needs {nproc} GPUs (run under gpurun --gpus {nproc})")
    res = _launch(nproc, "gpu", "residual", str(tmp_path), timeout=300, exchange="p2p")
    assert len(res) == nproc and all(r["err"] < 1e-12 and r["err_steps"] < 1e-9 for r in res)


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpus_limiter_matches_serial_oracle(tmp_path, exchange):
    """collective mft_limiter_zhang_shu (bit-identical owned rows) and the stage limiter inside the multi-GPU SSPRK33 step"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    res = _launch(2, "gpu", "limiter", str(tmp_path), timeout=240, exchange=exchange)
    assert all(r["err"] == 0.0 and r["err_steps"] < 1e-9 for r in res)


@pytest.mark.gpu
def test_two_gpus_device_setup_match_serial_oracle(tmp_path):
    """RBFFDEngineCUDA(setup="device") on every rank: kNN tables / weights from the GPU pipeline (row f1), then the same
    multi-GPU rhs! and time loop against the serial oracle (weights agree to rounding, so 1e-9 / 1e-8)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    res = _launch(2, "gpu", "residual", str(tmp_path), timeout=240, exchange="p2p", setup="device")
    assert all(r["err"] < 1e-9 and r["err_steps"] < 1e-8 for r in res)


@pytest.mark.parametrize("nranks,shape,wide", [(1, (40, 30), False), (4, (64, 48), False), (5, (50, 70), False), (8, (96, 64), False),
                                               (4, (64, 48), True), (7, (60, 50), True)])
def test_partition_planner_invariants(nranks, shape, wide):
    """host planner on 1/4/5/8 ranks (threads stand in for the ranks; the planner only needs an allgather): ownership is a
    partition of the cloud, every stencil column of an owned row is local, halo blocks are grouped by owner, and the send
    lists of one rank are exactly what its peers expect to receive, in the receivers' halo order"""
    import threading

    import numpy as np

    import mft_b200 as m
    from mft_b200 import partition

    cl = m.cloud.jittered_lattice(shape[0], shape[1], 10.0, 10.0 * shape[1] / shape[0], seed=3)
    nb_g, dx_min, dx_avg = cases.orc.point_data(cl.points, 20)
    barrier, slots, parts, errs = threading.Barrier(nranks), [None] * nranks, [None] * nranks, []

    def work(r):
        def allgather(obj):
            slots[r] = obj
            barrier.wait()
            out = list(slots)
            barrier.wait()
            return out
        try:
            parts[r] = partition.build_rank_partition(cl.points, [np.asarray(b) for b in cl.boundary_idxs], cl.boundary_normals,
                                                      r, nranks, 3, 3, 20, allgather, wide_halo=wide)
        except Exception as e:   # noqa: BLE001
            errs.append(e)
            barrier.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    owned = np.concatenate([p.owned_gid for p in parts])
    assert len(owned) == len(cl.points) == len(np.unique(owned))
    sizes = [p.n_local for p in parts]
    assert max(sizes) - min(sizes) <= 1                                   # contiguous equal ranges of the curve
    owner_of = np.empty(len(cl.points), dtype=np.int64)
    for p in parts:
        owner_of[p.owned_gid] = p.rank
    for p in parts:
        assert p.dx_min == dx_min and abs(p.dx_avg - dx_avg) <= 1e-15
        assert np.array_equal(p.neighbors_owned, nb_g[p.owned_gid])      # global stencils, bit-identical tables
        local = set(p.local_gid.tolist())
        assert set(np.unique(p.neighbors_owned).tolist()) <= local       # every column of an owned row is local
        assert np.array_equal(owner_of[p.halo_gid], p.halo_owner) and (np.diff(p.halo_owner) >= 0).all()
        assert (p.halo_owner != p.rank).all()
        off = 0
        for q, cnt in zip(p.peers, p.recv_count):
            want = p.halo_gid[off:off + cnt]
            assert (p.halo_owner[off:off + cnt] == q).all()
            sender = parts[q]
            i = sender.peers.index(p.rank)
            assert np.array_equal(sender.owned_gid[sender.send_idx[i]], want)   # what q sends is what p expects, in order
            off += cnt
        assert off == p.n_halo
        if wide:
            # wide halo: the foreign rows whose stencils contain an owned point (R_r) are halo rows, every one of their stencil
            # columns is local, the column-only points carry no operator row, and the narrow plan is a subset
            touches = np.isin(nb_g, p.owned_gid).any(axis=1)
            R = np.setdiff1d(np.nonzero(touches)[0], p.owned_gid)
            assert set(R.tolist()) <= set(p.halo_gid.tolist())
            assert set(np.unique(nb_g[R]).tolist()) <= local
            col_only = p.neighbors_halo[:, 0] < 0
            assert col_only.any() and not np.isin(p.halo_gid[col_only], R).any()
            assert np.array_equal(p.neighbors_halo[~col_only], nb_g[p.halo_gid[~col_only]])
            for A in p.ops:
                assert A.tocsr()[p.n_local + np.nonzero(col_only)[0]].nnz == 0
        # boundary points: each global boundary point belongs to exactly one rank's list
    for g in range(4):
        got = np.concatenate([p.owned_gid[p.boundary_idxs[g]] for p in parts])
        assert np.array_equal(np.sort(got), np.sort(cl.boundary_idxs[g]))


def _plan_all(points, boundary_idxs, boundary_normals, nranks, nv=20, wide=False):
    """the production planner on `nranks` ranks, threads standing in for the ranks"""
    import threading

    from mft_b200 import partition

    barrier, slots, parts, errs = threading.Barrier(nranks), [None] * nranks, [None] * nranks, []

    def work(r):
        def allgather(obj):
            slots[r] = obj
            barrier.wait()
            out = list(slots)
            barrier.wait()
            return out
        try:
            parts[r] = partition.build_rank_partition(points, boundary_idxs, boundary_normals, r, nranks, 3, 3, nv, allgather, wide_halo=wide)
        except Exception as e:   # noqa: BLE001
            errs.append(e)
            barrier.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    return parts


def test_reverse_halo_is_exact_on_a_graded_cloud():
    """kNN is not symmetric: on a graded cloud a foreign row can hold an owned point in its stencil without being anywhere near
    the owned rows' own stencils (3 such rows were missed on this kind of cloud).  R_r comes from the other ranks'
    reports now: every foreign row that references an owned point must be a halo row, on every rank; and the padded kNN box
    must have grown where the coarse region's stencil radius exceeds the default padding (tables equal the global ones)."""
    import numpy as np

    rng = np.random.default_rng(11)
    pts = np.concatenate([rng.random((2000, 2)) * 10.0, 5.0 + 0.05 * rng.standard_normal((18000, 2))])
    pts = np.unique(pts, axis=0)
    nb_g, dx_min, dx_avg = cases.orc.point_data(pts, 20)
    parts = _plan_all(pts, [np.zeros(0, dtype=np.int64)], [np.zeros((0, 2))], 8)
    missed = 0
    for p in parts:
        assert np.array_equal(p.neighbors_owned, nb_g[p.owned_gid])          # exact kNN despite the partition box
        touches = np.isin(nb_g, p.owned_gid).any(axis=1)
        R = np.setdiff1d(np.nonzero(touches)[0], p.owned_gid)
        missed += len(np.setdiff1d(R, p.halo_gid))
        F = np.setdiff1d(np.unique(nb_g[p.owned_gid]), p.owned_gid)
        assert len(np.setdiff1d(F, p.halo_gid)) == 0
        has_row = p.neighbors_halo[:, 0] >= 0
        assert np.array_equal(p.neighbors_halo[has_row], nb_g[p.halo_gid[has_row]])
    assert missed == 0
