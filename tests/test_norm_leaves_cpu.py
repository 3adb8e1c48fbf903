"""The one-pass statistic behind the fused step's ode_maximum (csrc/mft_fused_kernels.cuh), checked as
MATHEMATICS on the CPU tier: for states u_i in R^4 and any mean m, the lexicographic maximum of the deviation vectors
|u_i - m| (maximum(::StructArray{SVector}) with isless, src/auxiliary/mpi.jl:71-81) equals the lexicographic maximum over the 16
"leaves" leaf(s) = lexmax_i (s0 rho_i, s1 m1_i, s2 m2_i, s3 E_i), s in {+,-}^4 -- whatever the mean is, including heavy exact ties.
Also: leaves merge associatively / commutatively, so any split of the points into blocks / ranks gives the same answer.
(The device code itself is compared with the two-pass kernels bit for bit in tests/test_zz_j_fused_step_gpu.py.)"""
import itertools

import numpy as np
import pytest


def lexmax_rows(a):
    """lexicographic maximum of the rows of a (n, 4)"""
    order = np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))
    return a[order[-1]]


def leaves(u):
    """16 x 4: for every sign pattern the state that maximises (s0 rho, s1 m1, s2 m2, s3 E) lexicographically"""
    out = []
    for s in itertools.product((1.0, -1.0), repeat=4):
        key = u * np.asarray(s)
        order = np.lexsort((key[:, 3], key[:, 2], key[:, 1], key[:, 0]))
        out.append(u[order[-1]])
    return np.asarray(out)


def merge(la, lb):
    """leaf-wise merge of two leaf sets (what blocks, groups and ranks do)"""
    out = []
    for k, s in enumerate(itertools.product((1.0, -1.0), repeat=4)):
        pair = np.stack([la[k], lb[k]]) * np.asarray(s)
        out.append((la[k], lb[k])[int(np.lexsort((pair[:, 3], pair[:, 2], pair[:, 1], pair[:, 0]))[-1])])
    return np.asarray(out)


def norms_from_leaves(lv, mean):
    return lexmax_rows(np.abs(lv - mean))


def brute(u, mean):
    return lexmax_rows(np.abs(u - mean))


def _tie_heavy(rng, n, levels):
    """states on a coarse dyadic grid: every key of the lexicographic order is decided among exact ties, and |x - m| is exact
    for means on the same grid (no rounding ties by construction)"""
    return rng.integers(-levels, levels + 1, size=(n, 4)).astype(np.float64) / 8.0


@pytest.mark.parametrize("levels", [1, 2, 5, 40])
def test_leaves_are_sufficient_on_tie_heavy_states(levels):
    rng = np.random.default_rng(levels)
    for trial in range(60):
        u = _tie_heavy(rng, int(rng.integers(1, 400)), levels)
        lv = leaves(u)
        for _ in range(12):
            mean = rng.integers(-levels - 2, levels + 3, size=4).astype(np.float64) / 8.0 + rng.choice([0.0, 1.0 / 16.0])
            assert np.array_equal(norms_from_leaves(lv, mean), brute(u, mean)), (trial, mean)
        # the reference's own mean (sum / (V N), recursive_length) and the plain mean
        for div in (4.0 * len(u), float(len(u))):
            mean = u.sum(axis=0) / div
            assert np.array_equal(norms_from_leaves(lv, mean), brute(u, mean))


def test_leaves_on_generic_and_plateau_states():
    rng = np.random.default_rng(7)
    for trial in range(60):
        n = int(rng.integers(2, 600))
        u = rng.standard_normal((n, 4))
        u[:, 0] = np.abs(u[:, 0]) + 0.5
        if trial % 3 == 1:   # a plateau: rho AND m1 tie exactly on many points (an undisturbed region), the rest varies
            idx = rng.choice(n, n // 2, replace=False)
            u[idx, 0] = u[:, 0].max() + 0.25
            u[idx, 1] = 0.0
        if trial % 3 == 2:   # a noisy plateau whose noise is resolved by the deviation: m1 is noise EVERYWHERE, so is its mean
            idx = rng.choice(n, n // 2, replace=False)
            u[idx, 0] = u[:, 0].max() + 0.25
            u[:, 1] = 1e-17 * rng.standard_normal(n)
        lv = leaves(u)
        mean = u.sum(axis=0) / (4.0 * n)
        if trial % 3 == 2:   # (make sure this really is a case without rounding ties on the plateau)
            dev = np.abs(u[u[:, 0] == u[:, 0].max(), 1] - mean[1])
            if len(np.unique(dev)) < len(dev):
                continue
        assert np.array_equal(norms_from_leaves(lv, mean), brute(u, mean)), trial


def test_leaves_merge_like_the_blocks_and_ranks_do():
    rng = np.random.default_rng(11)
    for trial in range(30):
        u = _tie_heavy(rng, 500, 3)
        cuts = np.sort(rng.choice(np.arange(1, 500), size=int(rng.integers(1, 9)), replace=False))
        parts = np.split(u, cuts)
        order = rng.permutation(len(parts))           # any combination order
        acc = leaves(parts[order[0]])
        for k in order[1:]:
            acc = merge(acc, leaves(parts[k]))
        whole = leaves(u)
        mean = u.sum(axis=0) / (4.0 * len(u))
        assert np.array_equal(np.abs(acc - mean), np.abs(whole - mean)) or np.array_equal(norms_from_leaves(acc, mean), norms_from_leaves(whole, mean))
        assert np.array_equal(norms_from_leaves(acc, mean), brute(u, mean))


def test_a_rounding_tie_is_what_leaves_cannot_see():
    """The documented limit: a plateau that ties exactly on rho, carries rounding noise in m1 while the MEAN of m1 is
    O(1) -- so |m1 - mean| rounds to one double for the whole plateau -- and varies materially in m2.  The reference's
    lexicographic maximum is then decided on m2 among ALL plateau points; the leaves only hold the points with extreme m1.  The
    first two components still agree (same rounded deviations); the device counts such rows in pass A and reports MFT_ENORMS
    (tests/test_zz_j_fused_step_gpu.py::test_a_rounding_tie_is_reported_loudly)."""
    rng = np.random.default_rng(3)
    n = 400
    u = np.stack([np.full(n, 1.0), np.full(n, 1.0), np.zeros(n), np.full(n, 30.0)], axis=1)
    u[: n // 2, 0] = 2.0
    u[: n // 2, 1] = 1e-20 * rng.standard_normal(n // 2)
    u[: n // 2, 2] = np.sin(np.arange(n // 2))
    mean = u.sum(axis=0) / (4.0 * n)
    assert len(np.unique(np.abs(u[: n // 2, 1] - mean[1]))) == 1          # the rounding tie
    got, want = norms_from_leaves(leaves(u), mean), brute(u, mean)
    assert np.array_equal(got[:2], want[:2]) and got[2] < want[2]


# ---- the second density level (round 2): rounding ties on the FIRST key ---------------------------------------------------------

def _steps_inside(x, k, direction):
    for _ in range(k):
        x = np.nextafter(x, direction)
    return x


def leaves2(u, near_ulps=4):
    """the record of the device statistic: per side the leaves of the extreme density and of the next distinct density, the latter
    only if it lies within `near_ulps` representable steps of the extreme (kNearUlps in csrc/mft_fused_kernels.cuh)"""
    out = []
    vals = np.unique(u[:, 0])
    for level_vals, inside in ((vals[::-1][:2], -np.inf), (vals[:2], np.inf)):   # max side: two largest distinct; min side: two smallest
        for i, v in enumerate(level_vals):
            if i == 1 and abs(v - level_vals[0]) > abs(_steps_inside(level_vals[0], near_ulps, inside) - level_vals[0]):
                continue
            out.append(leaves(u[u[:, 0] == v]))
    return np.concatenate(out)


def test_adjacent_densities_can_tie_after_rounding_and_the_second_level_sees_it():
    """17 far-field points share the largest density T (the 2-GPU bench cloud, profiles/r2y_norm_miss_diagnostic_2ranks.log); a stage
    update moves some of them to T + ulp.  With the mean on a finer grid (ode_mean divides by V*N) |T - m| and |T + ulp - m| round
    to the same double for half of the means: the lexicographic maximum is then decided on |m1 - mean| among BOTH groups.  One
    level of leaves (the exact maximum only) misses that; two levels do not -- for every mean."""
    rng = np.random.default_rng(5)
    T = 1.0 - 5 * 2.0 ** -53
    n = 4000
    u = np.stack([0.4 + 0.5 * rng.random(n), 1.0 + 1e-3 * rng.standard_normal(n), 1e-3 * rng.standard_normal(n), 30 + rng.random(n)], axis=1)
    tie = np.arange(17)
    u[tie, 0] = T
    u[tie, 1] = 1.0 + 5e-7 * np.linspace(-1.0, 1.0, 17)
    moved = tie[[2, 9]]                                    # two rows, not the ones with the extreme momenta
    one_level_fails = two_level_fails = collisions = 0
    for up in (np.nextafter(T, 2.0), np.nextafter(T, 0.0)):
        w = u.copy()
        w[moved, 0] = up
        base = w.sum(axis=0) / (4.0 * n)
        for k in range(-16, 17):                           # the last bits of the mean decide whether the two densities collide
            mean = base.copy()
            mean[0] = base[0] + k * 2.0 ** -56
            want = brute(w, mean)
            d0 = np.abs(w[:, 0] - mean[0])
            if len(np.unique(w[d0 == d0.max(), 0])) > 1:
                collisions += 1
            one_level_fails += not np.array_equal(norms_from_leaves(leaves(w), mean), want)
            two_level_fails += not np.array_equal(norms_from_leaves(leaves2(w), mean), want)
    assert collisions > 0 and one_level_fails > 0          # the failure mode is real ...
    assert two_level_fails == 0                            # ... and the second level removes it


def test_two_levels_on_tie_heavy_and_generic_states():
    rng = np.random.default_rng(13)
    for trial in range(80):
        u = _tie_heavy(rng, int(rng.integers(1, 300)), int(rng.integers(1, 6))) if trial % 2 else rng.standard_normal((int(rng.integers(1, 300)), 4))
        for div in (4.0 * len(u), float(len(u))):
            mean = u.sum(axis=0) / div
            assert np.array_equal(norms_from_leaves(leaves2(u), mean), brute(u, mean))


# ---- a model of the DEVICE bookkeeping (csrc/mft_fused_kernels.cuh: lex_side_update, slot_insert, norms_from_records) ------------
# The functions below restate, value for value, what a warp does with a chunk of 32 rows and how records are folded into each other
# (same branches, same window rule), so that the sufficiency and the order independence of the windowed two-level records can be
# checked on adversarial data without a GPU: densities on a few-ulp grid around the extremes, exact ties at every level, random
# splits into chunks / warps / ranks, random merge orders.  (The CUDA code itself is compared with the two-pass kernels on the
# device: tests/test_zz_j_fused_step_gpu.py.)

K_NEAR = 4
SIGNS = [tuple(1.0 if not (k >> c) & 1 else -1.0 for c in range(3)) for k in range(8)]   # leaf bit c set: component c+1 minimised


def _key(x):
    b = np.float64(x).view(np.int64)
    return int(b) ^ (0x7FFFFFFFFFFFFFFF if b < 0 else 0) if b < 0 else int(b)   # monotone in x (signed integer order)


def _near(side, a, b):
    """b lies at most K_NEAR representable steps inside of a (side 0: below, side 1: above)"""
    if not (np.isfinite(a) and np.isfinite(b)):
        return False
    ka, kb = _key(a), _key(b)
    return 0 <= (ka - kb if side == 0 else kb - ka) <= K_NEAR


def _leaf_better(a, b, k):
    for c in range(3):
        if SIGNS[k][c] * a[c] > SIGNS[k][c] * b[c]:
            return True
        if SIGNS[k][c] * a[c] < SIGNS[k][c] * b[c]:
            return False
    return False


def _tie_leaves(rows):
    """8 leaves of a tie set: rows (n, 3) = (m1, m2, E)"""
    out = []
    for k in range(8):
        best = rows[0]
        for r in rows[1:]:
            if _leaf_better(r, best, k):
                best = r
        out.append(tuple(best))
    return out


class SideRec:
    def __init__(self, side):
        self.side, self.e, self.e2, self.L, self.L2 = side, (-np.inf if side == 0 else np.inf), (-np.inf if side == 0 else np.inf), None, None

    def beyond(self, x, y):
        return x > y if self.side == 0 else x < y

    def insert(self, ext, leaves):
        """slot_insert for the 8 leaf slots of this side at once (their (e, e2) evolve identically)"""
        if leaves is None or not np.isfinite(ext):
            return
        if self.beyond(ext, self.e):
            if _near(self.side, ext, self.e):
                self.e2, self.L2 = self.e, self.L
            else:
                self.e2, self.L2 = (-np.inf if self.side == 0 else np.inf), None
            self.e, self.L = ext, list(leaves)
        elif ext == self.e:
            self.L = [leaves[k] if _leaf_better(leaves[k], self.L[k], k) else self.L[k] for k in range(8)]
        elif _near(self.side, self.e, ext):
            if self.beyond(ext, self.e2):
                self.e2, self.L2 = ext, list(leaves)
            elif ext == self.e2:
                self.L2 = [leaves[k] if _leaf_better(leaves[k], self.L2[k], k) else self.L2[k] for k in range(8)]

    def merge(self, other):
        self.insert(other.e, other.L)
        self.insert(other.e2, other.L2)

    def chunk(self, u):
        """lex_side_update: one chunk of <= 32 rows (n, 4) against the running record"""
        side = self.side
        has2 = self.L2 is not None

        def window_bound(a):
            if not np.isfinite(a):
                return a
            x = np.float64(a)
            for _ in range(K_NEAR):
                x = np.nextafter(x, -np.inf if side == 0 else np.inf)
            return float(x)

        thr = self.e2 if has2 else window_bound(self.e)
        reaches = (lambda x, y: x >= y) if side == 0 else (lambda x, y: x <= y)
        hot = [i for i in range(len(u)) if reaches(u[i, 0], thr)]
        for _level in range(2):
            if not hot:
                return
            cm = max(u[i, 0] for i in hot) if side == 0 else min(u[i, 0] for i in hot)
            tied = [i for i in hot if u[i, 0] == cm]
            hot = [i for i in hot if u[i, 0] != cm]
            leaves = _tie_leaves(u[tied][:, 1:])
            if self.beyond(cm, self.e):
                keep = _near(side, cm, self.e)
                self.e2, self.L2 = (self.e, self.L) if keep else ((-np.inf if side == 0 else np.inf), None)
                self.e, self.L = cm, leaves
                has2 = keep
                thr = self.e2 if keep else window_bound(cm)
            elif cm == self.e:
                self.L = [leaves[k] if _leaf_better(leaves[k], self.L[k], k) else self.L[k] for k in range(8)]
            elif has2 and cm == thr:
                self.L2 = [leaves[k] if _leaf_better(leaves[k], self.L2[k], k) else self.L2[k] for k in range(8)]
            else:
                self.e2, self.L2, has2, thr = cm, leaves, True, cm
            hot = [i for i in hot if reaches(u[i, 0], thr)]

    def state(self):
        return (self.e, tuple(self.L) if self.L else None, self.e2 if self.L2 else None, tuple(self.L2) if self.L2 else None)


def device_record(u, rng, chunk=32):
    """rows -> chunks -> 'warps' (random runs of chunks) -> random merge tree, as blocks / groups / ranks would"""
    chunks = [u[i:i + chunk] for i in range(0, len(u), chunk)]
    warps = []
    i = 0
    while i < len(chunks):
        n = int(rng.integers(1, 6))
        recs = [SideRec(0), SideRec(1)]
        for ch in chunks[i:i + n]:
            for r in recs:
                r.chunk(ch)
        warps.append(recs)
        i += n
    order = rng.permutation(len(warps))
    acc = [SideRec(0), SideRec(1)]
    pending = [warps[k] for k in order]
    while len(pending) > 1:                       # random pairwise merges (any tree, any order)
        a = pending.pop(int(rng.integers(len(pending))))
        b = pending.pop(int(rng.integers(len(pending))))
        for s in range(2):
            a[s].merge(b[s])
        pending.append(a)
    for s in range(2):
        acc[s].merge(pending[0][s])
    return acc


def norms_from_device_record(rec, mean):
    cands = []
    for r in rec:
        for ext, L in ((r.e, r.L), (r.e2, r.L2)):
            if L is not None and np.isfinite(ext):
                cands += [(ext,) + tuple(l) for l in L]
    return lexmax_rows(np.abs(np.asarray(cands) - mean))


def _adversarial_states(rng, n):
    """densities on a one-ulp grid around both extremes (exact ties AND neighbours), momenta / energy on a few exact levels"""
    hi, lo = 1.0 - 5 * 2.0 ** -53, 0.5 + 3 * 2.0 ** -53
    u = np.stack([0.6 + 0.3 * rng.random(n), rng.integers(-3, 4, n) / 4.0, rng.integers(-2, 3, n) / 4.0, 30.0 + rng.integers(0, 3, n) / 2.0], axis=1)
    for base, sgn in ((hi, -1.0), (lo, 1.0)):
        idx = rng.choice(n, int(rng.integers(2, 40)), replace=False)
        steps = rng.integers(0, 7, len(idx))       # 0 .. 6 ulps inside the extreme: inside and outside the window
        vals = np.full(len(idx), base)
        for k in range(6):
            vals = np.where(steps > k, np.nextafter(vals, -np.inf if sgn < 0 else np.inf), vals)
        u[idx, 0] = vals
        u[idx, 1] = 1.0 + 1e-6 * rng.standard_normal(len(idx)) * (rng.random(len(idx)) < 0.7)   # some exact ties on m1 too
    return u


def test_device_bookkeeping_is_sufficient_and_order_independent():
    rng = np.random.default_rng(21)
    decided_by_second_level = 0
    for trial in range(120):
        n = int(rng.integers(40, 700))
        u = _adversarial_states(rng, n)
        rec = device_record(u, rng)
        again = device_record(u[rng.permutation(n)], rng)      # other chunks, other warps, another merge tree
        assert [r.state() for r in rec] == [r.state() for r in again], trial
        base = u.sum(axis=0) / (4.0 * n)
        for k in range(-6, 7):                                 # sweep the last bits of the mean: collisions come and go
            mean = base.copy()
            mean[0] = base[0] + k * 2.0 ** -56
            want = brute(u, mean)
            got = norms_from_device_record(rec, mean)
            assert np.array_equal(got, want), (trial, k, got, want)
            one = norms_from_leaves(leaves(u), mean)
            decided_by_second_level += not np.array_equal(one, want)
    assert decided_by_second_level > 0      # the sweep really contains cases that one density level gets wrong
