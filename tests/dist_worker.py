"""Worker launched by torch.distributed.run from tests/test_multi_rank.py.

--mode cpu : gloo, no GPU.  Builds the rank's partition with the production planner (partition.py) and runs the
             multi-rank rhs! ALGORITHM (halo exchange of u, forward pass on owned rows, halo exchange of g,
             transposed pass) in numpy/scipy on the local operators; compares the owned rows with the serial oracle on the
             global cloud.  This pins the partition / halo plan / algorithm that the CUDA path implements.
--mode gpu : nccl, one GPU per rank.  Runs the real library (NCCL halo exchange inside libmft_b200.so) and compares rhs!
             and a few SSPRK33 steps with the serial oracle.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GAMMA = 1.4


def euler_flux(u):
    rho, m1, m2, E = u
    v1, v2 = m1 / rho, m2 / rho
    p = (GAMMA - 1) * (E - 0.5 * (m1 * v1 + m2 * v2))
    return np.stack([m1, m1 * v1 + p, m1 * v2, (E + p) * v1]), np.stack([m2, m2 * v1, m2 * v2 + p, (E + p) * v2]), (v1, v2, p)


def exchange(dist, part, field):
    """field: (W, n_tot); fills the halo tail from the owners (the plan of partition.py, object collectives on gloo)"""
    n_local = part.n_local
    out = {int(q): np.ascontiguousarray(field[:, idx]) for q, idx in zip(part.peers, part.send_idx)}
    boxes = [None] * part.nranks
    dist.all_gather_object(boxes, out)
    off = n_local
    for q, cnt in zip(part.peers, part.recv_count):
        if cnt:
            field[:, off:off + cnt] = boxes[q][part.rank]
        off += cnt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="cpu")
    ap.add_argument("--out", default="")
    ap.add_argument("--source", default="upwind")
    ap.add_argument("--exchange", default="p2p")
    ap.add_argument("--setup", default="host", help="device: every rank builds its kNN tables / weights with the GPU pipeline")
    ap.add_argument("--fused", type=int, default=1, help="0: the separate stage / boundary / norm / put / wait kernels (MFT_OPT_FUSED_STEP = 0)")
    ap.add_argument("--same-device", type=int, default=0,
                    help="1: every rank uses cuda:0 (CUDA IPC between processes on ONE GPU; the ranks' kernels are time-sliced): the "
                         "multi-rank code path on a single-GPU box.  Process group: gloo (NCCL refuses two ranks on one device)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import mft_b200 as m
    import mft_oracle as orc
    from mft_b200 import partition

    rank = int(os.environ["RANK"])
    if args.mode == "gpu" and args.same_device:
        dist.init_process_group("gloo")
    elif args.mode == "gpu":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist.init_process_group("nccl")
    else:
        dist.init_process_group("gloo")
    comm = partition.TorchComm()

    cl = m.cloud.jittered_lattice(72, 60, 10.0, 10.0 * 60 / 72, seed=4)
    names = dict(left=1, right=2, bottom=3, top=4)
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, GAMMA, center=(5.0, 4.0))
    PHS = 5 if args.source == "flyer" else 3      # 4th derivatives (Flyer hyperviscosity) want r^5
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(PHS)))

    # ---- serial oracle on the global cloud (every rank computes it; it is small) -------------------------------
    nb, dx_min, dx_avg = orc.point_data(cl.points, basis.nv)
    ops = m.setup_ops.compute_flux_operator(cl.points, nb, PHS, 3)
    obc = [orc.OracleBC(orc.BC_DIRICHLET, cl.boundary_idxs[g], cl.boundary_normals[g], value_fn=lambda x, t: ic(x, t))
           for g in range(4)]
    if args.source == "flyer":
        # hyperviscosity.jl:34-50 with the product's (batched-LU) weights, so that serial and partitioned H agree bit for bit
        hops = m.setup_ops.compute_flux_operator(cl.points, nb, PHS, 3, 4)
        src_o = orc.OracleSource(kind=orc.SRC_HV_FLYER, hv=orc.JuliaCSC((hops[0] + hops[1]).tocsc()), gamma=1.0 * dx_min ** 4)
    elif args.source == "tominec":
        # hyperviscosity.jl:101-119: H = lap' * lap with the product's (batched-LU) weights, gamma = dx_min^4.5
        lops = m.setup_ops.compute_flux_operator(cl.points, nb, PHS, 3, 2)
        lap = (lops[0] + lops[1]).tocsc()
        src_o = orc.OracleSource(kind=orc.SRC_HV_TOMINEC, hv=orc.JuliaCSC((lap.T @ lap).tocsc()), gamma=1.0 * dx_min ** 4.5)
    else:
        src_o = orc.source_residual(dx_avg, polydeg=3) if args.source == "residual" else orc.source_upwind(dx_avg)
    P = orc.OracleProblem(cl.points, 4, orc.EQ_EULER2D, [GAMMA], ops[0], ops[1], obc, [src_o])
    u0 = ic(cl.points, 0.0) * (1.0 + 0.01 * np.sin(cl.points[:, 0]))
    u_ser = u0.copy()
    du_ser = P.rhs(u_ser, 0.0)

    # Zhang-Shu limiter case: a state with low-density / low-pressure pockets, limited serially on the global cloud
    LIM_THR, LIM_VAR = (0.6, 5.0), (orc.VAR_DENSITY, orc.VAR_PRESSURE)
    u_lim0 = u0.copy()
    u_lim0[0, 100:160] *= 0.3
    u_lim0[3, 900:960] *= 0.1
    u_lim_ser = orc.limiter_zhang_shu(u_lim0.copy(), nb, LIM_THR, LIM_VAR, GAMMA)

    results = {}
    if args.mode == "cpu":
        part = partition.build_rank_partition(cl.points, cl.boundary_idxs, cl.boundary_normals, comm.rank, comm.nranks,
                                              PHS, 3, basis.nv, comm.allgather, wide_halo=args.source == "tominec")
        gid = part.local_gid
        nl = part.n_local
        # the local operator rows of owned points are the global rows, bit for bit
        Dxg = ops[0].tocsr()
        Dxl = part.ops[0].tocsr()
        for i in range(0, nl, 37):
            gl, ll = Dxg[gid[i]], Dxl[i]
            assert np.array_equal(np.sort(gid[ll.indices]), np.sort(gl.indices))
            o1, o2 = np.argsort(gid[ll.indices]), np.argsort(gl.indices)
            assert np.array_equal(ll.data[o1], gl.data[o2])
        assert abs(part.dx_avg - dx_avg) < 1e-15 and part.dx_min == dx_min
        # ownership is a partition of the cloud
        owned_all = comm.allgather(part.owned_gid)
        allg = np.concatenate(owned_all)
        assert len(allg) == len(cl.points) and len(np.unique(allg)) == len(allg)
        # ---- the multi-rank algorithm on the local data ------------------------------------------------------------
        u = np.ascontiguousarray(u0[:, gid])
        u[:, nl:] = np.nan                                  # halo values must come from the exchange
        for g in range(4):                                  # BC pass 1 on owned boundary points
            bi = part.boundary_idxs[g]
            u[:, bi] = ic(part.points[bi], 0.0)
        exchange(dist, part, u)
        assert np.array_equal(u, u_ser[:, gid])             # halo copies carry the owner's BC-imposed values
        if args.source == "limiter":
            # multi-rank limiter ALGORITHM: per (threshold, variable) pass one u halo refresh, then the owned rows are limited
            # from their global stencils in local numbering (what mft_limiter_zhang_shu does on every rank)
            lut = np.full(len(cl.points), -1, dtype=np.int64)
            lut[gid] = np.arange(len(gid))
            nbl = lut[part.neighbors_owned]
            assert (nbl >= 0).all()
            w = np.ascontiguousarray(u_lim0[:, gid])
            w[:, nl:] = np.nan
            for thr, var in zip(LIM_THR, LIM_VAR):
                exchange(dist, part, w)
                val = (lambda q: q[0]) if var == orc.VAR_DENSITY else (lambda q: (GAMMA - 1) * (q[3] - 0.5 * (q[1] * q[1] + q[2] * q[2]) / q[0]))
                vmin = np.min(np.stack([val(w[:, nbl[:, q]]) for q in range(nbl.shape[1])]), axis=0)
                mean = np.zeros((4, nl))
                for q in range(nbl.shape[1]):
                    mean = mean + w[:, nbl[:, q]]
                mean = mean / nbl.shape[1]
                lim = vmin < thr
                vm = val(mean)
                with np.errstate(all="ignore"):
                    theta = (vm - thr) / (vm - vmin)
                    new = theta * w[:, :nl] + (1 - theta) * mean
                w[:, :nl] = np.where(lim & (new != 0).any(axis=0), new, w[:, :nl])
            ref = u_lim_ser[:, part.owned_gid]
            err = float(np.abs(w[:, :nl] - ref).max() / np.abs(ref).max())
            changed = int((ref != u_lim0[:, part.owned_gid]).any(axis=0).sum())
            assert err < 1e-14, err                      # fma vs separate rounding in the blend only
            results = dict(rank=rank, n_local=nl, n_halo=part.n_halo, err=err, changed=changed)
            allres = comm.allgather(results)
            assert sum(r["changed"] for r in allres) > 50
            if rank == 0:
                print("MULTI_RANK_OK", allres)
                if args.out:
                    import json

                    json.dump(allres, open(args.out, "w"))
            dist.barrier()
            dist.destroy_process_group()
            return
        F, G, (v1, v2, p) = euler_flux(u)
        Dx, Dy = part.ops[0].tocsr(), part.ops[1].tocsr()
        du = np.stack([-(Dx[:nl] @ F[v]) - (Dy[:nl] @ G[v]) for v in range(4)])
        if args.source in ("flyer", "tominec"):
            # SourceHyperviscosityFlyer on a partition: du += -gamma H u on owned rows, only the u halo is needed.
            # SourceHyperviscosityTominec: H = L'L reaches a second stencil ring -> wide halo (column-only halo points)
            import types

            dom = types.SimpleNamespace(partition=part, cloud=cl, pd=types.SimpleNamespace(dx_min=part.dx_min))
            solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA())
            if args.source == "flyer":
                hv = m.SourceHyperviscosityFlyer(solver, None, dom, k=2, c=1.0)
            else:
                hv = m.SourceHyperviscosityTominec(solver, None, dom, c=1.0)
                assert (part.neighbors_halo[:, 0] < 0).any(), "the wide halo added no column-only points"
                narrow = partition.build_rank_partition(cl.points, cl.boundary_idxs, cl.boundary_normals, comm.rank, comm.nranks,
                                                        PHS, 3, basis.nv, comm.allgather)
                assert narrow.n_halo < part.n_halo and np.array_equal(narrow.owned_gid, part.owned_gid)
                try:
                    m.SourceHyperviscosityTominec(solver, None, types.SimpleNamespace(partition=narrow, cloud=cl, pd=dom.pd))
                    raise AssertionError("the narrow halo must be refused")
                except ValueError:
                    pass
            H = hv.hv_differentiation_matrix.tocsr()
            assert H.shape == (len(gid), len(gid)) and H[nl:].nnz == 0
            Hg = src_o.hv.scipy.tocsr()
            for i in range(0, nl, 41):        # owned rows of the local H are the global rows (Flyer: bit for bit)
                gl, ll = Hg[gid[i]], H[i]
                if args.source == "flyer":
                    o1, o2 = np.argsort(gid[ll.indices]), np.argsort(gl.indices)
                    assert np.array_equal(gid[ll.indices][o1], gl.indices[o2]) and np.array_equal(ll.data[o1], gl.data[o2])
                else:                         # L'L: same entries up to the summation order inside the sparse product
                    dl = np.zeros(len(cl.points))
                    dl[gid[ll.indices]] = ll.data
                    dg = np.zeros(len(cl.points))
                    dg[gl.indices] = gl.data
                    assert np.abs(dl - dg).max() <= 1e-13 * np.abs(dg).max()
            for v in range(4):
                du[v] -= hv.gamma * (H[:nl] @ u[v])
            for g in range(4):
                du[:, part.boundary_idxs[g]] = 0.0
            ref = du_ser[:, part.owned_gid]
            err = max(np.abs(du[v] - ref[v]).max() / np.abs(du_ser[v]).max() for v in range(4))
            assert err < 1e-12, err
            results = dict(rank=rank, n_local=nl, n_halo=part.n_halo, err=float(err))
            allres = comm.allgather(results)
            if rank == 0:
                print("MULTI_RANK_OK", allres)
                if args.out:
                    import json

                    json.dump(allres, open(args.out, "w"))
            dist.barrier()
            dist.destroy_process_group()
            return
        eps = 0.5 * part.dx_avg * (np.hypot(v1, v2) + np.sqrt(GAMMA * p / u[0]))[:nl]
        g8 = np.zeros((8, len(gid)))
        for v in range(4):
            g8[v, :nl] = eps * (Dx[:nl] @ u[v])
            g8[4 + v, :nl] = eps * (Dy[:nl] @ u[v])
        g8[:, nl:] = np.nan
        exchange(dist, part, g8)
        for v in range(4):
            du[v] -= (Dx.T @ g8[v])[:nl] + (Dy.T @ g8[4 + v])[:nl]
        for g in range(4):
            du[:, part.boundary_idxs[g]] = 0.0
        ref = du_ser[:, part.owned_gid]
        err = max(np.abs(du[v] - ref[v]).max() / np.abs(du_ser[v]).max() for v in range(4))
        results = dict(rank=rank, n_local=nl, n_halo=part.n_halo, err=float(err))
        assert args.source == "upwind"
        assert err < 1e-12, err
    else:
        solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(device=0 if args.same_device else int(os.environ.get("LOCAL_RANK", rank)),
                                                                    diagnostics=True, exchange=args.exchange, setup=args.setup,
                                                                    fused_step=bool(args.fused)))
        domain = m.ParallelPointCloudDomain(solver, cl, names, comm, wide_halo=args.source == "tominec")
        part = domain.partition
        eq = m.CompressibleEulerEquations2D(GAMMA)
        bc = {k: m.BoundaryConditionDirichlet(ic) for k in names}
        if args.source in ("upwind", "limiter"):
            srcs = m.SourceTerms(rv=m.SourceUpwindViscosityTominec(solver, eq, domain))
        elif args.source == "flyer":
            srcs = m.SourceTerms(hv=m.SourceHyperviscosityFlyer(solver, eq, domain, k=2, c=1.0))
        elif args.source == "tominec":
            srcs = m.SourceTerms(hv=m.SourceHyperviscosityTominec(solver, eq, domain, c=1.0))
        else:
            srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3))
        semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs)
        gid = part.local_gid
        nl = part.n_local
        if args.source == "limiter":
            # collective limiter call on every rank, owned rows bit-identical to the serial oracle (same fma blend)
            lim = m.PositivityPreservingLimiterZhangShu(thresholds=LIM_THR, variables=(m.density, m.pressure))
            w = np.ascontiguousarray(u_lim0[:, gid])
            w[:, nl:] = 0.0
            lim(w, semi)
            assert np.array_equal(w[:, :nl], u_lim_ser[:, part.owned_gid])
            # and as SSPRK33 stage limiter inside the (graph-replayed) multi-GPU step, against the serial loop
            import ctypes as C

            thr1, var1 = (0.9,), (orc.VAR_DENSITY,)
            lim1 = m.PositivityPreservingLimiterZhangShu(thresholds=thr1, variables=(m.density,))
            dt = 0.1 * dx_min / 8.0
            ode = m.ODEProblem(np.ascontiguousarray(u0[:, gid]), (0.0, 4 * dt), semi)
            sol = m.solve(ode, m.SSPRK33(stage_limiter=lim1), dt=dt, nsteps=4)
            lib = orc.lib()
            us = u0.copy()
            k = P.rhs(us, 0.0)
            for _ in range(4):
                uprev = us.copy()
                for st in (1, 2, 3):
                    lib.orc_ssprk33_stage(C.c_int64(us.size), st, C.c_double(dt), C.c_void_p(uprev.ctypes.data),
                                          C.c_void_p(k.ctypes.data), C.c_void_p(us.ctypes.data))
                    orc.limiter_zhang_shu(us, nb, thr1, var1, GAMMA)
                    k = P.rhs(us, 0.0)
            err2 = max(np.abs(sol.u[v, :nl] - us[v, part.owned_gid]).max() / np.abs(us[v]).max() for v in range(4))
            assert err2 < 1e-9, err2
            results = dict(rank=rank, n_local=nl, n_halo=part.n_halo, err=0.0, err_steps=float(err2))
            semi.close()
            allres = comm.allgather(results)
            if rank == 0:
                print("MULTI_RANK_OK", allres)
                if args.out:
                    import json

                    json.dump(allres, open(args.out, "w"))
            dist.barrier()
            dist.destroy_process_group()
            return
        u = np.ascontiguousarray(u0[:, gid])
        u[:, nl:] = 0.0
        du = np.zeros_like(u)
        m.rhs_(du, u, semi, 0.0)
        ref = du_ser[:, part.owned_gid]
        err = max(np.abs(du[v, :nl] - ref[v]).max() / np.abs(du_ser[v]).max() for v in range(4))
        assert np.array_equal(u[:, :nl], u_ser[:, part.owned_gid])
        assert np.array_equal(u[:, nl:], u_ser[:, part.halo_gid]), "halo state after rhs! must be the owners' values"
        assert (du[:, nl:] == 0).all()                       # reset_halos! parallel_rbfsolver.jl:74-91
        # time integration: 10 SSPRK33 steps with the history callback, against the serial oracle
        dt = 0.1 * dx_min / 8.0
        hist = None if args.source in ("flyer", "tominec") else 3
        u_ref, _ = P.solve_ssprk33(u0, 0.0, dt, 10, approx_order=hist)
        ode = m.ODEProblem(np.ascontiguousarray(u0[:, gid]), (0.0, 10 * dt), semi)
        sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=None if hist is None else m.HistoryCallback(3), nsteps=10)
        err2 = max(np.abs(sol.u[v, :nl] - u_ref[v, part.owned_gid]).max() / np.abs(u_ref[v]).max() for v in range(4))
        miss = np.zeros(1)
        m._lib.check(m.load().mft_get_field(semi.ctx, m._lib.FIELD_NORM_MISSES, m._lib.ptr(miss)))
        results = dict(rank=rank, n_local=nl, n_halo=part.n_halo, err=float(err), err_steps=float(err2), norm_misses=int(miss[0]))
        assert miss[0] == 0, miss
        # device-made weights differ from the serial oracle's by rounding (different elimination order): 1e-9 instead of 1e-12
        assert err < (1e-12 if args.setup == "host" else 1e-9), err
        assert err2 < (1e-9 if args.setup == "host" else 1e-8), err2
        semi.close()
    allres = comm.allgather(results)
    if rank == 0:
        print("MULTI_RANK_OK", allres)
        if args.out:
            import json

            json.dump(allres, open(args.out, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
