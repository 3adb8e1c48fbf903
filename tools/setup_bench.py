"""Times the setup pipeline on a GPU box: mft_setup_knn / mft_setup_rbf_weights (host arrays in,
host arrays out, so the numbers are end-to-end: cell sort + H2D + kernels + D2H) beside the host mirror (cKDTree + batched
LAPACK LU on all cores).  One JSON line.   python tools/setup_bench.py --n-side 1024 [--host-sample 200000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-side", type=int, default=1024)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--host-sample", type=int, default=200000, help="points the host weight solve is timed on (scaled up)")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import mft_b200 as m

    cl = m.cloud.jittered_lattice(args.n_side, args.n_side, 10.0, 10.0, seed=0).points
    n = len(cl)
    out = {"points": n, "k": args.k, "host_cores": os.cpu_count()}

    def best(f):
        ts = []
        for _ in range(args.reps):
            t = time.perf_counter()
            r = f()
            ts.append(time.perf_counter() - t)
        return min(ts), r

    m.setup_ops.knn_device(cl[:4096], args.k)      # context creation, module load
    t, (nb, dmin, davg) = best(lambda: m.setup_ops.knn_device(cl, args.k))
    out["knn_device_s"] = t
    t, (nbh, hmin, havg) = best(lambda: m.setup_ops.knn(cl, args.k))
    out["knn_host_s"] = t
    out["knn_equal"] = bool(np.array_equal(nb, nbh) and dmin == hmin and davg == havg)
    t, (wx, wy) = best(lambda: m.setup_ops.rbf_fd_weights_device(cl, nb, 3, 3))
    out["weights_device_s"] = t
    ns = min(n, args.host_sample)
    t, (hx, hy) = best(lambda: m.setup_ops.rbf_fd_weights(cl, nb[:ns], 3, 3))
    out["weights_host_s_scaled"] = t * n / ns
    out["weights_max_rel_diff"] = float(max(np.abs(wx[:ns] - hx).max() / np.abs(hx).max(), np.abs(wy[:ns] - hy).max() / np.abs(hy).max()))
    out["setup_speedup"] = (out["knn_host_s"] + out["weights_host_s_scaled"]) / (out["knn_device_s"] + out["weights_device_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
