"""Design-time model of the union-tile kernels (mft_tile_kernels.cuh): for the bench cloud (jittered lattice, Hilbert
order, k = 20) it measures, per tile of 128*R rows,
  * the size of the stencil union (points per row, 128-byte lines touched by the phase-1 load),
  * the union steps per row when a thread owns R consecutive rows,
  * the LDS.128 bank-conflict degree (8 lanes per phase, 8 bank groups of 16 bytes) for natural slots, greedily coloured
    slots, and two record copies with a per-phase optimal choice,
for the forward operator and its transpose.  The measured shared-memory wavefront counts (profiles/README.md) match this
model to ~4 %.  Usage: python tools/tile_sim.py [n_side=256] [tiles=12]"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mft_b200 as m  # noqa: E402
from mft_b200 import setup_ops  # noqa: E402


def build(n_side):
    cl = m.cloud.jittered_lattice(n_side, n_side, 10.0, 10.0, seed=0)
    cl = m.cloud.reorder(cl, m._lib.sfc_order(cl.points))
    nbr, _, _ = setup_ops.knn(cl.points, 20)
    N = nbr.shape[0]
    fwd = [r for r in np.sort(nbr, axis=1)]
    A = sp.csr_matrix((np.ones(N * 20), (np.repeat(np.arange(N), 20), nbr.reshape(-1))), shape=(N, N))
    AT = A.T.tocsr()
    AT.sort_indices()
    tra = [AT.indices[AT.indptr[i]:AT.indptr[i + 1]] for i in range(N)]
    return N, fwd, tra


def greedy(reqs, nU, nb=8):
    adj = [dict() for _ in range(nU)]
    for s in reqs:
        for a in s:
            for b in s:
                if a != b:
                    adj[a][b] = adj[a].get(b, 0) + 1
    order = np.argsort([-sum(d.values()) for d in adj], kind="stable")
    col = -np.ones(nU, dtype=int)
    fill = np.zeros(nb, dtype=int)

    def choose(p):
        cost = np.zeros(nb)
        for q, w in adj[p].items():
            if col[q] >= 0:
                cost[col[q]] += w
        return int(np.argmin(cost + 1e-3 * fill))

    for p in order:
        col[p] = choose(p)
        fill[col[p]] += 1
    for _ in range(2):
        for p in order:
            fill[col[p]] -= 1
            col[p] = -1
            col[p] = choose(p)
            fill[col[p]] += 1
    return col


def best_of_two(s, c0, c1):
    best = 99
    for bits in range(1 << len(s)):
        cnt = [0] * 8
        for i, x in enumerate(s):
            cnt[c1[x] if (bits >> i) & 1 else c0[x]] += 1
        best = min(best, max(cnt))
        if best == 1:
            break
    return best


def analyse(N, rows, R, name, ntiles, rng):
    blk = 128 * R
    out = []
    for b in rng.choice(N // blk, size=min(ntiles, N // blk), replace=False):
        rr = rows[b * blk:(b + 1) * blk]
        U = np.unique(np.concatenate(rr))
        loc = {g: i for i, g in enumerate(U)}
        reqs, steps = [], 0
        for w in range(4):
            lanes = [np.unique(np.concatenate(rr[(w * 32 + l) * R:(w * 32 + l + 1) * R])) for l in range(32)]
            width = max(len(x) for x in lanes)
            steps += width
            for c in range(width):
                for ph in range(4):
                    s = sorted(set(loc[lanes[l][c]] for l in range(ph * 8, ph * 8 + 8) if c < len(lanes[l])))
                    if s:
                        reqs.append(s)
        c0 = greedy(reqs, len(U))
        c1 = rng.permutation(len(U)) % 8
        nat = np.mean([np.bincount(np.array(s) % 8, minlength=8).max() for s in reqs])
        col = np.mean([np.bincount(c0[s], minlength=8).max() for s in reqs])
        two = np.mean([best_of_two(s, c0, c1) for s in reqs]) if R <= 2 else float("nan")
        out.append((len(U) / blk, len(np.unique(U // 4)), steps * 32 / blk, nat, col, two))
    o = np.mean(out, axis=0)
    print(f"{name} R={R}: union/row {o[0]:.2f}, 128-B lines/tile {o[1]:.0f}, union steps/row {o[2]:.1f}, "
          f"LDS.128 conflict degree natural {o[3]:.2f} coloured {o[4]:.2f} two copies {o[5]:.2f}")


if __name__ == "__main__":
    n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ntiles = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    N, fwd, tra = build(n_side)
    rng = np.random.default_rng(1)
    for R in (1, 2, 4):
        analyse(N, fwd, R, "forward   ", ntiles, rng)
        analyse(N, tra, R, "transposed", ntiles, rng)
