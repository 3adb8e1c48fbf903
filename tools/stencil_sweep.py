#!/usr/bin/env python
"""BASELINE.json configs[4]: stencil-width sweep (k = 13..50) of the fused flux + stencil kernel and of the full
Euler + residual-viscosity stage, reporting achieved algorithmic GB/s against the measured HBM peak.

N = 2^20 jittered-lattice points; neighbour tables are real kNN, weights are synthetic (`default_rng(3)`, row sums zero)
because only bandwidth is measured.  Operators enter through mft_set_operator_ell.
Output: one JSON line per k (also appended to profiles/ when --out is given).
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ks", default="13,15,20,25,30,36,42,50")
    ap.add_argument("--n-side", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--out", default="")
    ap.add_argument("--tiles", default="31,0", help="MFT_OPT_TILE values to compare per width: 31 = union-tile kernels (default), 0 = thread-per-row sliced ELL")
    args = ap.parse_args()
    import mft_b200 as m

    L = m._lib
    lib = m.load()
    peak = 6464.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    cl = m.cloud.jittered_lattice(args.n_side, args.n_side, 10.0, 10.0, seed=3, ring=False)
    pts = cl.points
    n = pts.shape[0]
    perm1 = np.ascontiguousarray(L.sfc_order(pts) + 1)
    u0 = np.ascontiguousarray(m.cloud.isentropic_vortex(pts, 1.4))
    rng = np.random.default_rng(3)
    rows = []
    for k in [int(x) for x in args.ks.split(",")]:
        nb, dx_min, dx_avg = m.setup_ops.knn(pts, k)
        w = rng.standard_normal((2, n, k)) / dx_avg
        w -= w.mean(axis=2, keepdims=True)               # row sums zero: constants are annihilated
        nbr1 = np.ascontiguousarray(nb + 1)
        res = {"k": k, "points": n}
        for tile, mode in [(int(t), md) for t in args.tiles.split(",") for md in ("flux_only", "residual_viscosity")]:
            ctx = C.c_void_p()
            L.check(lib.mft_ctx_create(C.byref(ctx), 0, n, 0, 4, 2, k))
            g = np.array([1.4])
            L.check(lib.mft_set_equation(ctx, L.EQ_EULER2D, L.ptr(g), 1))
            L.check(lib.mft_set_permutation(ctx, L.ptr(perm1)))
            L.check(lib.mft_set_option(ctx, L.OPT_TILE, float(tile)))
            wx, wy = np.ascontiguousarray(w[0]), np.ascontiguousarray(w[1])
            L.check(lib.mft_set_operator_ell(ctx, L.ptr(nbr1), L.ptr(wx), L.ptr(wy)))
            if mode == "residual_viscosity":
                prm = np.array([1.0, 1.0, dx_avg, 3.0])
                L.check(lib.mft_add_source(ctx, L.SRC_RESIDUAL, L.ptr(prm), 4, None, None, None))
            L.check(lib.mft_finalize(ctx))
            L.check(lib.mft_upload_state(ctx, L.soa_ptrs(u0)))
            if mode == "residual_viscosity":
                L.check(lib.mft_history_push(ctx, 0.0, 0, 3))
                L.check(lib.mft_history_push(ctx, 1e-3, 1, 3))
            for _ in range(5):
                L.check(lib.mft_rhs(ctx, 0.0, None, None, L.MEM_DEVICE))
            L.check(lib.mft_timer_start(ctx))
            for _ in range(args.reps):
                L.check(lib.mft_rhs(ctx, 0.0, None, None, L.MEM_DEVICE))
            ms = C.c_double()
            L.check(lib.mft_timer_stop(ctx, C.byref(ms)))
            per = ms.value / args.reps
            bytes_pt = (20 * k + 64) if mode == "flux_only" else (40 * k + 288)
            gbs = n * bytes_pt / (per * 1e-3) / 1e9
            res[f"{mode}_tile{tile}"] = {"ms_per_rhs": round(per, 4), "alg_bytes_per_point": bytes_pt, "GBps": round(gbs, 1),
                         "frac_of_measured_peak": round(gbs / peak, 4), "Gpoint_rhs_per_s": round(n / (per * 1e-3) / 1e9, 3)}
            L.check(lib.mft_ctx_destroy(ctx))
        for md in ("flux_only", "residual_viscosity"):   # which kernel family wins at this width (the selection rule's evidence)
            best = max((int(t) for t in args.tiles.split(",")), key=lambda t: res[f"{md}_tile{t}"]["GBps"])
            res[f"{md}_best_tile"] = best
        print(json.dumps(res), flush=True)
        rows.append(res)
    if args.out:
        json.dump({"peak_GBps": peak, "rows": rows}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
