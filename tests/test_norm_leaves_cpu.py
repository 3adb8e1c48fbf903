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
