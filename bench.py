#!/usr/bin/env python
"""bench.py -- point-stage updates/sec of the Euler RBF-FD rhs! hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (libmft_b200.so)
    python bench.py --impl reference ...                      # the reference's CPU path (C port in oracle/), host cores

Workload at N=1: BASELINE.json configs[1] -- 2-D compressible Euler (isentropic vortex) with residual viscosity on a
1,048,576-point jittered-lattice cloud (+ boundary ring), PHS r^3 + degree-3 polynomials, k = 20, HistoryCallback(3),
SSPRK33 with fixed dt.  A "step" is one SSPRK33 time step: 3 rhs! evaluations + 3 stage updates + the history
callback, i.e. 3 point-stage updates per point.  Synthetic data; setup (cloud, kNN, weights, layouts) is untimed.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "point-stage updates/sec of Euler RBF-FD rhs!"
UNIT = "point-stage updates/s"
GAMMA = 1.4
K_STENCIL = 20
STAGES = 3
CLOUD_ORDER = "hilbert"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """samples before this point are warm-up; keep the last one (GPU already under load) and everything after"""
        self.first = max(0, len(self.rows) - 1)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def parse(rows):
            sm, mx, reasons = [], [], set()
            for r in rows:
                try:
                    sm.append(float(r[1]))
                    mx.append(float(r[2]))
                    for nm, val in zip(names, r[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
                except Exception:
                    continue
            return sm, mx, reasons
        sm, mx, reasons = parse(self.rows[getattr(self, 'first', 0):])
        window = "timed region"
        if not sm:   # a timed region shorter than the sampling period: the samples of the warm-up steps right before it
            sm, mx, reasons = parse(self.rows[-3:])
            window = "warm-up steps just before the timed region (the region was shorter than one sampling period)"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def build_workload(nx, ny, seed, m, need_oracle_ops=False):
    """config 2 of SURVEY.md 8d: jittered lattice on [0,10]^2 + boundary ring, Dirichlet(vortex at t=0) on all sides"""
    t0 = time.time()
    cl = m.cloud.jittered_lattice(nx, ny, 10.0, 10.0 * ny / nx, seed=seed)
    if CLOUD_ORDER == "hilbert":   # emit the synthetic cloud along a Hilbert curve (point numbering is arbitrary)
        cl = m.cloud.reorder(cl, m._lib.sfc_order(cl.points))
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    return cl, basis, time.time() - t0


def vortex_ic(m, center):
    return lambda x, t, eq=None: m.cloud.isentropic_vortex(x, GAMMA, center=center)


GRID_FOR_GPUS = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}   # lattice multiples: fixed points per GPU (weak scaling)


def multi_gpu_parity(m, comm, dist, args, local_rank):
    """N > 1 only, OUTSIDE the timed region: a small jittered cloud partitioned over the same ranks, through the same engine /
    exchange settings as the timed run (fused step, peer-memory puts), against the SERIAL oracle on the global cloud (the
    checker; every rank evaluates it, the cloud is small): one rhs! (tolerance 1e-12) and 10 SSPRK33 steps with the history
    callback (1e-9).  Returns the max over ranks; the bench refuses to print a line above the tolerances."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mft_oracle as orc

    nx, ny = 40 * max(1, GRID_FOR_GPUS.get(comm.nranks, (comm.nranks, 1))[0]), 36 * max(1, GRID_FOR_GPUS.get(comm.nranks, (comm.nranks, 1))[1])
    cl = m.cloud.jittered_lattice(nx, ny, 10.0, 10.0 * ny / nx, seed=4)
    names = dict(left=1, right=2, bottom=3, top=4)
    ic = lambda x, t, e=None: m.cloud.isentropic_vortex(x, GAMMA, center=(5.0, 5.0 * ny / nx))   # noqa: E731
    basis = m.PointCloudBasis(m.Point2D(), 3, approximation_type=m.RBF(m.PolyharmonicSpline(3)))
    nb, dx_min, dx_avg = orc.point_data(cl.points, basis.nv)
    ops = m.setup_ops.compute_flux_operator(cl.points, nb, 3, 3)
    obc = [orc.OracleBC(orc.BC_DIRICHLET, cl.boundary_idxs[g], cl.boundary_normals[g], value_fn=lambda x, t: ic(x, t)) for g in range(4)]
    residual = args.source == "residual"
    src_o = orc.source_residual(dx_avg, polydeg=3) if residual else orc.source_upwind(dx_avg)
    P = orc.OracleProblem(cl.points, 4, orc.EQ_EULER2D, [GAMMA], ops[0], ops[1], obc, [src_o])
    u0 = ic(cl.points, 0.0) * (1.0 + 0.01 * np.sin(cl.points[:, 0]))
    u_ser = u0.copy()
    du_ser = P.rhs(u_ser, 0.0)
    dt = 0.1 * dx_min / 8.0
    u_ref, _ = P.solve_ssprk33(u0, 0.0, dt, 10, approx_order=3 if residual else None)

    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(device=local_rank, exact_order=not args.fma, cuda_graph=int(args.graph),
                                                                exchange=args.exchange, tile=int(args.tile), tile_rows=int(args.tile_rows),
                                                                fused_step=bool(args.fused_step)))
    domain = m.ParallelPointCloudDomain(solver, cl, names, comm)
    part = domain.partition
    eq = m.CompressibleEulerEquations2D(GAMMA)
    bc = {k: m.BoundaryConditionDirichlet(ic) for k in names}
    srcs = (m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, polydeg=3)) if residual
            else m.SourceTerms(uw=m.SourceUpwindViscosityTominec(solver, eq, domain)))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs)
    gid, nl = part.local_gid, part.n_local
    u = np.ascontiguousarray(u0[:, gid])
    u[:, nl:] = 0.0
    du = np.zeros_like(u)
    m.rhs_(du, u, semi, 0.0)
    ref = du_ser[:, part.owned_gid]
    e_rhs = max(np.abs(du[v, :nl] - ref[v]).max() / np.abs(du_ser[v]).max() for v in range(4))
    ode = m.ODEProblem(np.ascontiguousarray(u0[:, gid]), (0.0, 10 * dt), semi)
    sol = m.solve(ode, m.SSPRK33(), dt=dt, callback=m.HistoryCallback(3) if residual else None, nsteps=10)
    e_steps = max(np.abs(sol.u[v, :nl] - u_ref[v, part.owned_gid]).max() / np.abs(u_ref[v]).max() for v in range(4))
    miss = np.zeros(1)
    m._lib.check(m.load().mft_get_field(semi.ctx, m._lib.FIELD_NORM_MISSES, m._lib.ptr(miss)))
    semi.close()
    t = torch.tensor([e_rhs, e_steps, miss[0]], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e_rhs, e_steps, misses = (float(x) for x in t.tolist())
    out = {"ranks": comm.nranks, "points": int(len(cl.points)), "rhs_relerr": e_rhs, "steps_relerr": e_steps, "steps": 10,
           "norm_misses": int(misses), "against": "serial oracle (oracle/mft_oracle.c) on the global cloud",
           "tolerances": {"rhs": 1e-12, "steps": 1e-9}, "ok": bool(e_rhs <= 1e-12 and e_steps <= 1e-9 and misses == 0)}
    if not out["ok"]:
        raise SystemExit(f"bench: multi-GPU parity check failed: {out}")
    return out


def run_ours(args):
    import mft_b200 as m

    lib = m.load()
    L = m._lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1:
        raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    multi = world > 1
    dist = None
    if multi:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl")
        from mft_b200 import partition
        comm = partition.TorchComm()
    parity = multi_gpu_parity(m, comm, dist, args, local_rank) if multi and not args.no_parity else None
    topo = None
    if multi and rank == 0:   # how the GPUs of this box reach each other (NV# = NVLink lanes; PIX/PXB/NODE/SYS = PCIe only)
        try:
            rows = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()
            topo = [" ".join(r.split()[:world + 1]) for r in rows if r.startswith("GPU")][:world]
        except Exception:
            topo = None
    gx, gy = GRID_FOR_GPUS.get(world, (world, 1))
    nx, ny = args.n_side * gx, args.n_side * gy
    cl, basis, _ = build_workload(nx, ny, 0, m)
    solver = m.PointCloudSolver(basis, engine=m.RBFFDEngineCUDA(device=local_rank, exact_order=not args.fma,
                                                                stage_weights=int(args.stage_weights),
                                                                cuda_graph=int(args.graph),
                                                                single_sweep_exact=bool(args.single_sweep),
                                                                exchange=args.exchange, pair_rows=int(args.pair_rows),
                                                                tile=int(args.tile), tile_rows=int(args.tile_rows),
                                                                prefetch_distance=args.pf_dist, setup=args.setup,
                                                                refine_order=bool(args.refine_order),
                                                                fused_step=bool(args.fused_step), pdl=bool(args.pdl),
                                                                layout_device=bool(args.layout_device)))
    names = dict(left=1, right=2, bottom=3, top=4)
    t_setup = time.time()
    domain = m.ParallelPointCloudDomain(solver, cl, names, comm) if multi else m.PointCloudDomain(solver, cl, names)
    eq = m.CompressibleEulerEquations2D(GAMMA)
    residual = args.source == "residual"
    if args.workload == "sod":     # configs[3]: Sod shock tube, slip walls top/bottom, Dirichlet left/right
        ic = lambda x, t, e=None: m.cloud.sod(x, GAMMA, x_mid=5.0)   # noqa: E731
        bc = dict(left=m.BoundaryConditionDirichlet(ic), right=m.BoundaryConditionDirichlet(ic),
                  bottom=m.boundary_condition_slip_wall, top=m.boundary_condition_slip_wall)
    else:                          # configs[1]: isentropic vortex, Dirichlet data on all sides
        ic = vortex_ic(m, (5.0, 5.0 * ny / nx))
        bc = {k: m.BoundaryConditionDirichlet(ic) for k in names}
    if residual:
        srcs = m.SourceTerms(rv=m.SourceResidualViscosityTominec(solver, eq, domain, c_rv=1.0, c_uw=1.0, polydeg=3))
    else:                          # configs[2]: upwind (first-order) viscosity, no history callback
        srcs = m.SourceTerms(uw=m.SourceUpwindViscosityTominec(solver, eq, domain, c_uw=1.0))
    semi = m.SemidiscretizationHyperbolic(domain, eq, ic, solver, boundary_conditions=bc, source_terms=srcs)
    t_setup = time.time() - t_setup
    N = cl.points.shape[0]                                    # global points
    n_own = domain.partition.n_local if multi else N          # rows this rank computes
    pd = domain.pd
    # CFL 0.1*dx_min/(|v|+c): |v|+c ~ 7.4 for the vortex base state, ~ 2.2 behind the Sod shock
    dt = 0.1 * pd.dx_min / (8.0 if args.workload == "vortex" else 3.0)
    ode = m.semidiscretize(semi, (0.0, 1.0))
    u0 = ode.u0
    ctx = semi.ctx

    def barrier():
        if multi:
            dist.barrier()

    def max_over_ranks(x):
        if not multi:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def steps(n, t, it0):
        for i in range(n):
            L.check(lib.mft_ssprk_step(ctx, L.SSPRK33, t, dt))
            t += dt
            if residual:
                L.check(lib.mft_history_push(ctx, t, it0 + i + 1, 3))
        return t

    # ---- device-resident timed region: barrier + sync on both sides, CUDA events on the ctx stream, max over ranks ----
    # The fused step's one-pass ode_maximum statistic is verified row by row inside pass A; rows it missed (a rounding tie decided
    # the reference's lexicographic order: DESIGN.md 3d) make mft_synchronize fail with MFT_ENORMS.  The bench reports them and
    # NEVER times a region that had any: misses during warm-up are noted (norm_misses.warmup), misses inside the timed region send
    # every rank to the two-pass kernels (MFT_OPT_FUSED_STEP = 0) and the region is run again.
    MFT_ENORMS = -6

    def sync_soft():
        rc = lib.mft_synchronize(ctx)
        if rc not in (0, MFT_ENORMS):
            L.check(rc)

    def norm_misses():
        x = np.zeros(1)
        L.check(lib.mft_get_field(ctx, L.FIELD_NORM_MISSES, L.ptr(x)))
        return int(x[0])

    sampler = ClockSampler(local_rank)

    def timed_region():
        L.check(lib.mft_upload_state(ctx, L.soa_ptrs(u0)))
        if residual:
            L.check(lib.mft_history_push(ctx, 0.0, 0, 3))
        if rank == 0:
            sampler.start()   # nvidia-smi needs ~100 ms to deliver its first sample: start before the warm-up steps
        m0 = norm_misses()
        t_ = steps(args.warmup, 0.0, 0)
        sync_soft()
        m1 = norm_misses()
        barrier()
        if rank == 0:
            sampler.mark()
        l0 = lib.mft_launch_count(ctx)
        L.check(lib.mft_timer_start(ctx))
        t_ = steps(args.steps, t_, args.warmup)
        ms_ = C.c_double()
        L.check(lib.mft_timer_stop(ctx, C.byref(ms_)))
        barrier()
        sync_soft()
        m2 = norm_misses()
        return t_, ms_.value, lib.mft_launch_count(ctx) - l0, (sampler.stop() if rank == 0 else None), m1 - m0, m2 - m1

    t, ms_value, launches, clocks, miss_warm, miss_timed = timed_region()
    fallback = None
    if max_over_ranks(float(miss_timed)) > 0:
        if rank == 0:
            sys.stderr.write("bench: the one-pass norms missed rows inside the timed region; re-running with the two-pass kernels\n")
        L.check(lib.mft_set_option(ctx, L.OPT_FUSED_STEP, 0.0))
        sampler = ClockSampler(local_rank)
        t, ms_value, launches, clocks, _, miss_again = timed_region()
        fallback = "MFT_OPT_FUSED_STEP = 0 (two-pass norm kernels): the one-pass statistic missed rows inside the first timed region"
        args.fused_step = 0
        if max_over_ranks(float(miss_again)) > 0:
            raise SystemExit("bench: norm misses with the two-pass kernels (cannot happen: they do not use the statistic)")
    total_ms = max_over_ranks(ms_value)
    ms_per_step = total_ms / args.steps
    value = N * STAGES * args.steps / (total_ms * 1e-3)
    miss_detail = {"warmup": int(max_over_ranks(float(miss_warm))), "timed": 0, "fallback": fallback,
                   "note": "rows per rank (max over ranks) that exceeded the fused step's one-pass ode_maximum statistic; the timed region never has any"}

    # sanity: the state is still finite after the timed steps
    u_end = np.empty_like(u0)
    L.check(lib.mft_download_state(ctx, L.soa_ptrs(u_end)))
    if not np.isfinite(u_end[:, :n_own]).all():
        raise SystemExit("bench: non-finite state after the timed region")
    miss = np.zeros(1)
    L.check(lib.mft_get_field(ctx, L.FIELD_NORM_MISSES, L.ptr(miss)))

    # ---- per-kernel CUDA-event pass (same steps, events around every launch on the ctx stream) -----------
    L.check(lib.mft_set_kernel_timing(ctx, 1))
    t = steps(args.steps, t, args.warmup + args.steps)
    ktime = {}
    for name, cls in (("pass_a", L.K_PASS_A), ("pass_b", L.K_PASS_B), ("reduce", L.K_REDUCE), ("stage", L.K_STAGE),
                      ("bc", L.K_BC), ("other", L.K_OTHER)):
        kms, kn = C.c_double(), C.c_int64()
        L.check(lib.mft_kernel_time_ms(ctx, cls, C.byref(kms), C.byref(kn)))
        ktime[name] = (kms.value, kn.value)
    L.check(lib.mft_set_kernel_timing(ctx, 0))
    peak, peak_src = measured_peak()
    V, k = 4, K_STENCIL
    src_bytes = 288 if residual else 256                         # 40k + 288 (residual) / 40k + 256 (upwind)
    bytes_a = n_own * (20 * k + 8 * V + (8 * V if residual else 0) + 8 * V + 16 * V)  # idx+wx+wy, u gather, approx_du, du, g
    bytes_b = n_own * (20 * k + 16 * V + 16 * V)                 # idxT+wxT+wyT, g gather, du read+write
    a_ms, a_n = ktime["pass_a"]
    b_ms, b_n = ktime["pass_b"]
    dom, dom_bytes, dom_ms, dom_n = ("k_pass_a", bytes_a, a_ms, a_n) if a_ms >= b_ms else ("k_pass_b", bytes_b, b_ms, b_n)
    achieved = dom_bytes / (dom_ms / max(dom_n, 1) * 1e-3) / 1e9
    step_bytes = N * STAGES * (40 * k + src_bytes + 128)
    agg_peak = peak * world
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": TRAFFIC.get(dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes,
                "avg_launch_ms": round(dom_ms / max(dom_n, 1), 4),
                "whole_step": {"algorithmic_GBps": round(step_bytes / (ms_per_step * 1e-3) / 1e9, 1),
                               "frac_of_n_gpu_peak": round(step_bytes / (ms_per_step * 1e-3) / 1e9 / agg_peak, 4),
                               "bytes_per_point_stage": 40 * k + src_bytes + 128},
                "kernel_ms_per_step": {kk: round(v[0] / args.steps, 4) for kk, v in ktime.items()},
                "note": "per-kernel times: rank 0, CUDA events around every launch (eager replay of the same steps)"
                        + ("; at N > 1 the pass kernels' times include their in-kernel waits for the peers' halo rows and norm records, and "
                           "eagerly launched ranks drift apart -- the N = 1 line carries the roofline of the kernels themselves" if multi else "")}

    # ---- end-to-end through the reference-facing call: mft_rhs with HOST buffers (pinned) -------------------
    n_arr = u0.shape[1]
    u_h = L.pinned_empty((4, n_arr))
    du_h = L.pinned_empty((4, n_arr))
    u_h[:] = u_end
    up, dup = L.soa_ptrs(u_h), L.soa_ptrs(du_h)
    for _ in range(3):
        L.check(lib.mft_rhs(ctx, t, up, dup, L.MEM_HOST))
    e2e_calls = max(3, min(3 * args.steps, 30))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_calls):
        L.check(lib.mft_rhs(ctx, t, up, dup, L.MEM_HOST))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = N * e2e_calls / e2e_s
    n_touched = sum(len(domain.boundary_tags[k].idx) for k in names) + (u0.shape[1] - n_own if multi else 0)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": STAGES * 32 * N,
           "d2h_bytes_per_step": STAGES * 32 * (N + n_touched * world),
           "call": "mft_rhs(MFT_MEM_HOST): u H2D, rhs!, du D2H + the rows of u that rhs! changes (boundary points, halo), "
                   "pinned host arrays", "calls_timed": e2e_calls}
    # informational (single GPU): the device-resident stepper driven with host-visible state EVERY STEP instead of every
    # stage -- u H2D, one SSPRK33 step (3 stages + the f(u_n) evaluation an uploaded state needs), u D2H.  The strict
    # per-rhs! number above stays the e2e value.
    if not multi:
        try:
            tt = t
            for _ in range(3):
                L.check(lib.mft_upload_state(ctx, up))
                L.check(lib.mft_ssprk_step(ctx, L.SSPRK33, tt, dt))
                L.check(lib.mft_download_state(ctx, up))
            calls = max(3, min(args.steps, 20))
            t0 = time.perf_counter()
            for _ in range(calls):
                L.check(lib.mft_upload_state(ctx, up))
                L.check(lib.mft_ssprk_step(ctx, L.SSPRK33, tt, dt))
                L.check(lib.mft_download_state(ctx, up))
                tt += dt
            el = time.perf_counter() - t0
            e2e["per_step_host_state"] = {"value": N * STAGES * calls / el, "unit": UNIT, "h2d_bytes_per_step": 32 * N,
                                          "d2h_bytes_per_step": 32 * N, "calls_timed": calls,
                                          "call": "mft_upload_state + mft_ssprk_step + mft_download_state per step (pinned)"}
        except Exception as exc:   # informational only: never lose the bench line over it
            e2e["per_step_host_state"] = {"unavailable": repr(exc)}

    # ---- CPU baseline: the oracle's C port of the reference structure, bounded sample, rank 0, N=1 only -----------
    cpu = None
    if not args.no_cpu_baseline and not multi:
        cpu = cpu_baseline(semi, domain, u_end, m, budget_s=args.cpu_seconds, residual=residual)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": (f"BASELINE configs[1]: 2D Euler isentropic vortex + residual viscosity, {N}-point "
                                       f"jittered cloud ({nx}x{ny} + ring), PHS3 deg3 k=20, SSPRK33 + HistoryCallback(3)")
                          if (args.workload, args.source) == ("vortex", "residual") else
                          (f"2D Euler {'Sod shock tube (slip walls top/bottom, Dirichlet left/right)' if args.workload == 'sod' else 'isentropic vortex'}"
                           f" + {'residual viscosity + HistoryCallback(3)' if residual else 'upwind (first-order) viscosity'}, {N}-point "
                           f"jittered cloud ({nx}x{ny} + ring), PHS3 deg3 k=20, SSPRK33"),
                          "points": N, "points_per_gpu": N // world, "k": k, "stages_per_step": STAGES,
                          "summation": "fma single-sweep" if args.fma else "reference order (bit-exact sums)",
                          "partition": "single GPU" if not multi else f"Hilbert-curve ranges over {world} ranks, halo exchange (u, g) + global norms per stage via {'NVLink peer-memory puts (CUDA IPC), graph-replayed' if args.exchange == 'p2p' else 'NCCL send/recv + all-gather'}",
                          "l2": "inputs larger than L2 (operators 2 x %.0f MB per GPU streamed every stage)" % (n_own * 20 * k / 1e6),
                          "setup_s": round(t_setup, 1), "setup": args.setup,
                          "layout": {"tile": int(args.tile), "tile_rows": int(args.tile_rows), "refine_order": int(args.refine_order),
                                     "exchange": args.exchange if multi else None, "fused_step": int(args.fused_step),
                                     "built_on": "device" if args.layout_device else "host threads"},
                          "gpu_topology": topo},
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
               "parity": parity, "norm_misses": int(miss[0]), "norm_misses_detail": miss_detail,
               # the reference's own (printed, never recorded) metric: PerformanceCallback's performance index
               # PID = runtime * nranks / (ndofsglobal * ncalls_rhs), src/callbacks_step/performance.jl:229-235 (one DOF = one point)
               "reference_native_metric": {"name": "PID [s per DOF per rhs! per rank]", "value": world / value}}
        print(json.dumps(out))
    semi.close()
    if multi:
        dist.barrier()
        dist.destroy_process_group()


# DRAM bytes per launch from the committed ncu --set full capture (profiles/); None until captured
TRAFFIC = {}
try:
    TRAFFIC = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
except Exception:
    pass


def cpu_baseline(semi, domain, u, m, budget_s=15.0, residual=True):
    """Times the reference's CPU execution structure (oracle/mft_oracle.c: CSC/Int64 operators, one SpMV per variable
    per direction, serial loops -- the reference's hot loops are single-threaded) on this box's host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mft_oracle as orc

    pd = domain.pd
    ops = semi.cache.rbf_differentiation_matrices
    obc = []
    all_dirichlet = True
    for name, bc, tag in semi._bc_groups:
        if hasattr(bc, "boundary_value_function"):
            vals = np.ascontiguousarray(bc.boundary_value_function(pd.points[tag.idx], 0.0, None))
            obc.append(orc.OracleBC(orc.BC_DIRICHLET, tag.idx, tag.normals, values=vals))
        else:
            all_dirichlet = False
            obc.append(orc.OracleBC(orc.BC_SLIP_WALL, tag.idx, tag.normals))
    src = orc.source_residual(pd.dx_avg, polydeg=3) if residual else orc.source_upwind(pd.dx_avg)
    src.success_iter = 5
    P = orc.OracleProblem(pd.points, 4, orc.EQ_EULER2D, [GAMMA], ops[0], ops[1], obc, [src])
    uu = np.ascontiguousarray(u.copy())
    t0 = time.perf_counter()
    P.rhs_repeat(uu, 1)
    one = time.perf_counter() - t0
    reps = int(max(2, min(200, budget_s / max(one, 1e-3))))
    t0 = time.perf_counter()
    P.rhs_repeat(uu, reps)
    dt = time.perf_counter() - t0
    out = {"value": pd.num_points * reps / dt, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"{reps} rhs! evaluations (Euler + {'residual' if residual else 'upwind'} viscosity) on the full {pd.num_points}-point cloud, "
                     "C port of the reference's serial CSC-SpMV structure (oracle/mft_oracle.c), 1 thread"}
    # variant (ii) of SURVEY.md 8(d): best-effort CPU (fused row-parallel kernel on all host cores, oracle/mft_cpu_fast.c)
    try:
        if not (all_dirichlet and residual):
            raise NotImplementedError("the fused CPU kernel covers the configs[1] workload only (Dirichlet data, residual viscosity)")
        bidx = np.concatenate([tag.idx for _, _, tag in semi._bc_groups])
        bvals = np.concatenate([np.asarray(bc.boundary_value_function(pd.points[tag.idx], 0.0, None))
                                for _, bc, tag in semi._bc_groups], axis=1)
        F = orc.FastCpuProblem(ops[0], ops[1], GAMMA, pd.dx_avg, bidx, bvals, success_iter=5)
        uu = np.ascontiguousarray(u.copy())
        t0 = time.perf_counter()
        F.rhs(uu, reps=1)
        one = time.perf_counter() - t0
        reps2 = int(max(3, min(500, 0.5 * budget_s / max(one, 1e-4))))
        t0 = time.perf_counter()
        F.rhs(uu, reps=reps2)
        dt2 = time.perf_counter() - t0
        out["best_effort"] = {"value": pd.num_points * reps2 / dt2, "unit": UNIT, "cores": int(orc.fast_lib().fast_max_threads()),
                              "kind": "port",
                              "sample": f"{reps2} rhs! evaluations on the same cloud, fused row-parallel CPU kernel "
                                        "(oracle/mft_cpu_fast.c: row-major operators, AoS state, one pass A + one pass B, "
                                        "pthreads over rows) -- not the reference's structure, the best a CPU port would do"}
    except Exception as exc:   # the baseline is a report, never a gate
        out["best_effort"] = {"unavailable": repr(exc)}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU path on this box's host cores.  The reference is pure Julia and Julia is not
    in this image, so the oracle's C port of its execution structure is timed (kind = "port"),
    single-threaded like the reference's hot loops (serial CSC SpMVs).  Same config as the GPU arm at N = 1: the full
    1,052,672-point cloud in the same numbering, Euler + residual viscosity, SSPRK33 with the HistoryCallback.  Nothing of the
    product's shared library is used: cloud numbering, kNN and weights come from oracle/ (bench_setup.py, mft_oracle.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import bench_setup as bs
    import mft_oracle as orc

    cm = bs.cloud_module()
    nx = ny = args.ref_n_side
    t_setup = time.time()
    cl = cm.jittered_lattice(nx, ny, 10.0, 10.0 * ny / nx, seed=0)
    if CLOUD_ORDER == "hilbert":
        cl = cm.reorder(cl, bs.hilbert_order(cl.points))
    nb, dx_min, dx_avg = orc.point_data(cl.points, K_STENCIL)
    ops = bs.flux_operator_batched(cl.points, nb, 3, 3)
    residual = args.source == "residual"
    if args.workload == "sod":
        ic = lambda x, t: cm.sod(x, GAMMA, x_mid=5.0)   # noqa: E731
        kinds = [orc.BC_DIRICHLET, orc.BC_DIRICHLET, orc.BC_SLIP_WALL, orc.BC_SLIP_WALL]
    else:
        ic = lambda x, t: cm.isentropic_vortex(x, GAMMA, center=(5.0, 5.0 * ny / nx))   # noqa: E731
        kinds = [orc.BC_DIRICHLET] * 4
    obc = [orc.OracleBC(kinds[g], cl.boundary_idxs[g], cl.boundary_normals[g],
                        values=np.ascontiguousarray(ic(cl.points[cl.boundary_idxs[g]], 0.0)) if kinds[g] == orc.BC_DIRICHLET else None)
           for g in range(4)]
    src = orc.source_residual(dx_avg, polydeg=3) if residual else orc.source_upwind(dx_avg)
    P = orc.OracleProblem(cl.points, 4, orc.EQ_EULER2D, [GAMMA], ops[0], ops[1], obc, [src])
    N = cl.points.shape[0]
    u = np.ascontiguousarray(ic(cl.points, 0.0))
    lib = orc.lib()
    dt = 0.1 * dx_min / (8.0 if args.workload == "vortex" else 3.0)
    n_el = u.size
    t_setup = time.time() - t_setup
    state = {"t": 0.0, "it": 0, "k": None}
    if residual:
        P.history_callback(u, 0.0, 0, 3)          # initialize!  history.jl:51-54
    state["k"] = P.rhs(u, 0.0)                    # FSAL: k = f(u_0)

    def step():
        # one SSPRK33 step of the CPU path, OrdinaryDiffEq's FSAL structure: 3 rhs! + 3 stage updates + the history callback
        t, k = state["t"], state["k"]
        uprev = u.copy()
        for s_, ts in ((1, t + dt), (2, t + dt / 2), (3, t + dt)):
            lib.orc_ssprk33_stage(C.c_int64(n_el), s_, C.c_double(dt), C.c_void_p(uprev.ctypes.data),
                                  C.c_void_p(k.ctypes.data), C.c_void_p(u.ctypes.data))
            k = P.rhs(u, ts)
        state["t"], state["it"], state["k"] = t + dt, state["it"] + 1, k
        if residual:
            P.history_callback(u, state["t"], state["it"], 3)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    if not np.isfinite(u).all():
        raise SystemExit("bench --impl reference: non-finite state")
    value = N * STAGES * args.steps / el
    same = (nx == args.n_side) and args.gpus == 1
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": (f"BASELINE configs[1]: 2D Euler isentropic vortex + residual viscosity, {N}-point "
                                   f"jittered cloud ({nx}x{ny} + ring), PHS3 deg3 k=20, SSPRK33 + HistoryCallback(3)")
                      if (args.workload, args.source) == ("vortex", "residual") else
                      f"2D Euler {args.workload} + {args.source} viscosity, {N}-point jittered cloud ({nx}x{ny} + ring), PHS3 deg3 k=20, SSPRK33",
                      "points": N, "k": K_STENCIL, "stages_per_step": STAGES, "setup_s": round(t_setup, 1),
                      "same_config_as_gpu_arm": bool(same),
                      "note": None if same else "CPU arm runs the N = 1 cloud (throughput metric; the N-GPU arm's cloud is N times larger)"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                            "sample": f"{args.steps} SSPRK33 steps (3 rhs! + history callback each) on the full {N}-point cloud; C port of "
                                      "the reference's serial structure (oracle/mft_oracle.c; Julia is not installed; the reference's hot "
                                      "loops are single-threaded)"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    try:   # the repo's own shared objects this process has mapped: oracle/ only (the product library is not on this path)
        libs = sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ROOT in ln and ".so" in ln})
        out["repo_native_libs"] = [os.path.relpath(p_, ROOT) for p_ in libs]
    except Exception:
        pass
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-side", type=int, default=1024, help="lattice side; 1024 -> the 1M-point cloud of configs[1]")
    ap.add_argument("--ref-n-side", type=int, default=1024, help="lattice side of the cloud the CPU arm runs (1024: the full configs[1] cloud)")
    ap.add_argument("--fma", action="store_true", help="single-sweep FMA summation instead of the reference order")
    ap.add_argument("--stage-weights", type=int, default=5, help="bit0: pass A slices staged in smem, bit1: pass B, bit2: one weight buffer refilled between the x and y sweeps")
    ap.add_argument("--graph", type=int, default=1, help="0 eager, 1 CUDA-graph replay on one GPU, 2 also multi-rank")
    ap.add_argument("--single-sweep", type=int, default=0, help="k=20 single-sweep exact pass A (register-parked y-products)")
    ap.add_argument("--tile", type=int, default=31, help="union-tile kernels (bit0: pass A, bit1: pass B, bit2: bank-coloured slots, bit3: two record copies, bit4 (31): tuned second copy)")
    ap.add_argument("--pf-dist", type=int, default=None, help="slices ahead for the L2 prefetch of operator data (0: off; default: 16 per SM)")
    ap.add_argument("--tile-rows", type=int, default=11, help="rows per thread of the tile kernels: units digit pass A, tens digit pass B")
    ap.add_argument("--pair-rows", type=int, default=1, help="row-pair (union stencil) operator layout (bit0: pass B, bit1: pass A)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="multi-GPU halo exchange mechanism")
    ap.add_argument("--setup", default="device", choices=["host", "device"],
                    help="kNN + RBF-FD weight generation (untimed setup): host numpy/LAPACK, or the GPU pipeline (mft_setup_*)")
    ap.add_argument("--workload", default="vortex", choices=["vortex", "sod"],
                    help="vortex: BASELINE configs[1] (default, the bench line); sod: configs[3] (shock tube, slip/Dirichlet mix)")
    ap.add_argument("--source", default="residual", choices=["residual", "upwind"],
                    help="stabilisation source: residual viscosity + history (configs[1], [3]) or upwind viscosity (configs[2])")
    ap.add_argument("--refine-order", type=int, default=0, help="1: order the rows inside a tile by D' row length (fewer padding steps in pass B)")
    ap.add_argument("--fused-step", type=int, default=1, help="0: separate stage / boundary / norm / halo put / wait kernels (round-1 sequence)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="label of the JSON line: weak (default: --n-side is the per-GPU lattice side) or strong (the caller chose --n-side so that the TOTAL cloud is fixed)")
    ap.add_argument("--layout-device", type=int, default=1, help="0: union-tile layouts on host threads instead of the device builder (same bytes)")
    ap.add_argument("--pdl", type=int, default=1, help="0: no programmatic dependent launch between the kernels of a fused stage")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the untimed parity check against the serial oracle")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cloud-order", default="hilbert", choices=["hilbert", "lattice"],
                    help="numbering of the synthetic cloud's points")
    args = ap.parse_args()
    global CLOUD_ORDER
    CLOUD_ORDER = args.cloud_order
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args)


if __name__ == "__main__":
    main()
